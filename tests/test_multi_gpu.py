"""Two GPUs, one process (rank) per GPU, individuals sharded across ranks, the library's own NCCL exchanges
(MIN all-reduce of first-allele keys, SUM all-reduce of the per-SNP counters, all-gather of thinned windows):
the merged result must equal the single-shard oracle bit for bit.  Skipped with fewer than 2 GPUs
(run with `gpurun --gpus 2 -- python -m pytest tests/test_multi_gpu.py -m gpu`)."""
import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

from tests.common import arg, load_case, oracle_roh_idx

pytestmark = pytest.mark.gpu


def _worker(rank, world, comm_id, q, name):
    from garlic_b200 import shard
    from garlic_b200.api import GarlicGPU
    ds, args = load_case(name)
    W = arg(args, "--winsize", cast=int)
    err = arg(args, "--error", cast=float)
    cutoff = arg(args, "--lod-cutoff", cast=float)
    ov = arg(args, "--overlap-frac", 0.25, float)
    lo, hi = shard.shard_range(ds.n_ind, world, rank)
    g = GarlicGPU(rank)
    g.comm_init(comm_id, rank, world)
    g.set_shape(hi - lo, ds.n_loci, ds.chr_offsets, ds.pos, ind_offset=lo)
    g.put_alleles(np.ascontiguousarray(ds.alleles[:, lo:hi]), 0)
    g.code_alleles()                                   # MIN all-reduce of the first-allele keys inside
    one = g.get_one_allele().copy()
    freq, keep, L = g.filter()                         # SUM all-reduce of the counters inside
    freq = freq.copy()
    c = [x.copy() for x in g.get_counts()]
    cens = [ds.centromeres.get("chr" + n, (0, 0)) for n in ds.chr_names]
    g.set_tables(err, 200000, np.array(cens, np.int32))
    kde = np.unique(np.minimum(np.array([1, 5, 6, 20, ds.n_ind - 1]), ds.n_ind - 1))
    parts = shard.split_individuals(kde, ds.n_ind, world)
    rows = max(len(p) for p in parts)
    thin = g.windows_gather(W, W, parts[rank], rows, world, exact=True).copy()
    roh = g.call_roh(W, cutoff, ov)
    q.put((rank, dict(one=one, freq=freq, counts=c, L=L, thin=thin, roh=roh)))
    g.close()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
@pytest.mark.parametrize("name", ["lod_0", "lod_small"])
def test_two_gpu_shards_equal_single_shard(name):
    from garlic_b200 import shard
    from garlic_b200.api import GarlicGPU
    from oracle import oracle as orc
    world = 2
    comm_id = GarlicGPU.comm_id()
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    procs = [ctx.Process(target=_worker, args=(r, world, comm_id, q, name)) for r in range(world)]
    for p in procs:
        p.start()
    got = dict(q.get() for _ in range(world))
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    ds, args = load_case(name)
    W = arg(args, "--winsize", cast=int)
    err = arg(args, "--error", cast=float)
    cutoff = arg(args, "--lod-cutoff", cast=float)
    ov = arg(args, "--overlap-frac", 0.25, float)
    geno, na, tot, one, freq = orc.code_tped(ds.alleles)
    res = orc.run_pipeline(ds, W, err, cutoff, ov)
    for r in range(world):
        assert np.array_equal(got[r]["one"], one)
        assert np.array_equal(got[r]["counts"][0], na) and np.array_equal(got[r]["counts"][1], tot)
        assert np.array_equal(got[r]["freq"], freq)
        assert got[r]["L"] == res["n_used"]
    # thinned windows: identical on both ranks, equal to the oracle's windows of those individuals
    assert np.array_equal(got[0]["thin"], got[1]["thin"], equal_nan=True)
    kde = np.unique(np.minimum(np.array([1, 5, 6, 20, ds.n_ind - 1]), ds.n_ind - 1))
    win = [c["win"][:, ::W] for c in res["chroms"]]
    want = np.concatenate(win, axis=1)[kde]
    rows = got[0]["thin"]
    rows = rows[~np.all(rows == -9999.0, axis=1)]
    ok = want != orc.MISSING
    assert np.array_equal(rows == orc.MISSING, ~ok)
    assert np.max(np.abs(rows[ok] - want[ok]) / np.maximum(np.abs(want[ok]), 1e-3)) <= 1e-9
    merged = shard.merge_roh([got[r]["roh"] for r in range(world)], ds.n_ind, world)
    assert [tuple(int(v) for v in r) for r in merged] == oracle_roh_idx(res)


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
@pytest.mark.parametrize("name", ["lod_0", "lod_2", "lod_small", "gl_pl", "auto_overlap_hg19", "freq_file", "lod_cm"])
def test_cli_two_gpus_equals_reference_binary(name):
    """garlic_b200 --gpus 2 (individuals sharded over two GPUs, NCCL exchanges inside the library): every output
    file equals the single-process reference binary's."""
    import gzip
    import os
    import tempfile
    from tests.common import GOLDEN, golden_text
    from tests.test_cli_gpu import run_cli
    with tempfile.TemporaryDirectory() as tmp:
        ds, args, r = run_cli(name, tmp, extra=["--gpus", "2"])
        assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
        out = os.path.join(tmp, "out")
        assert open(out + ".roh.bed").read() == golden_text(name, "out.roh.bed")
        if os.path.exists(os.path.join(GOLDEN, name, "out.freq")):
            assert gzip.open(out + ".freq.gz", "rt").read() == golden_text(name, "out.freq")
