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


def _run_ranks(target, world, extra, timeout=150):
    """Spawn one process per rank and collect one result each; a rank that dies or stalls fails the test instead of
    blocking it (workers also arm faulthandler so a stall leaves a traceback)."""
    import queue as _queue
    import time
    from garlic_b200.api import GarlicGPU
    comm_id = GarlicGPU.comm_id()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=target, args=(r, world, comm_id, q) + tuple(extra)) for r in range(world)]
    for p in procs:
        p.start()
    got, t0 = {}, time.time()
    try:
        while len(got) < world:
            try:
                r, val = q.get(timeout=2.0)
                got[r] = val
            except _queue.Empty:
                dead = [p.exitcode for p in procs if p.exitcode not in (None, 0)]
                assert not dead, "a rank exited with %s" % dead
                assert time.time() - t0 < timeout, "ranks stalled"
        for p in procs:
            p.join(60)
            assert p.exitcode == 0
    finally:
        for p in procs:
            if p.is_alive():
                p.kill()
    return got


def _worker(rank, world, comm_id, q, name):
    import faulthandler
    faulthandler.dump_traceback_later(100, exit=True)
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
    got = _run_ranks(_worker, world, (name,))
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


def _weighted_worker(rank, world, comm_id, q, ld_list):
    import faulthandler
    faulthandler.dump_traceback_later(100, exit=True)
    from garlic_b200 import shard
    from garlic_b200.api import GarlicGPU
    from garlic_b200.pipeline import interpolate_map
    ds, args = load_case("wlod_cm")
    lo, hi = shard.shard_range(ds.n_ind, world, rank)
    g = GarlicGPU(rank)
    g.comm_init(comm_id, rank, world)
    g.set_shape(hi - lo, ds.n_loci, ds.chr_offsets, ds.pos, ind_offset=lo)
    g.put_alleles(np.ascontiguousarray(ds.alleles[:, lo:hi]), 0)
    g.code_alleles()
    cens = [ds.centromeres.get("chr" + n, (0, 0)) for n in ds.chr_names]
    C = len(ds.chr_names)
    chr_param = np.array([[ds.map_pos[c][0], ds.map_pos[c][-1], cens[c][0], cens[c][1]] for c in range(C)], np.int32)
    freq, keep, L = g.filter(True, chr_param)
    kept = g.get_kept_index()
    pos = np.asarray(ds.pos)[kept]
    off = np.searchsorted(kept, np.asarray(ds.chr_offsets))
    gpos = np.empty(L)
    for c in range(C):
        gpos[off[c]:off[c + 1]], _ = interpolate_map(pos[off[c]:off[c + 1]], ds.map_pos[c], ds.map_cm[c])
    g.set_tables(0.001, 200000, np.array(cens, np.int32), gpos)
    g.set_wlod(1e-9, 7)
    ld = g.ld_band(25, None if ld_list is None else np.asarray(ld_list, np.int32), want_ld=True).copy()
    roh = g.call_roh(25, 0.5, 0.25, weighted=True)
    q.put((rank, dict(ld=ld, roh=roh, L=L)))
    g.close()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
@pytest.mark.parametrize("ld_list", [None, [0, 3, 4, 9, 10, 17, 20, 23]])
def test_two_gpu_weighted_ld_band_and_roh(ld_list):
    """--weighted over two shards: the LD individuals' bit-planes are all-reduced inside ld_band, every rank builds the
    same band (bit-identical to the single-shard oracle) and the merged wLOD ROH equal the oracle's."""
    from garlic_b200 import shard
    from garlic_b200.api import GarlicGPU
    from oracle import oracle as orc
    world = 2
    got = _run_ranks(_weighted_worker, world, (ld_list,))
    ds, args = load_case("wlod_cm")
    res = orc.run_pipeline(ds, 25, 0.001, 0.5, 0.25, weighted=True, cm=True,
                           ld_individuals=None if ld_list is None else np.asarray(ld_list, np.int32))
    want_ld = np.concatenate([c["LD"] for c in res["chroms"]], axis=0)
    for r in range(world):
        assert got[r]["L"] == res["n_used"]
        assert np.array_equal(got[r]["ld"], want_ld, equal_nan=True)
    merged = shard.merge_roh([got[r]["roh"] for r in range(world)], ds.n_ind, world)
    assert [tuple(int(v) for v in r) for r in merged] == oracle_roh_idx(res)


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
@pytest.mark.parametrize("name", ["lod_0", "lod_2", "lod_small", "gl_pl", "auto_overlap_hg19", "freq_file", "lod_cm", "wlod_cm", "wlod_phased"])
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


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
@pytest.mark.parametrize("name", ["lod_small", "wlod_cm"])
def test_cli_two_gpus_raw_lod(name):
    """--raw-lod --gpus 2: every rank dumps the windows of its own individuals, lines in individual order."""
    from tests.test_cli_gpu import check_raw_lod
    check_raw_lod(name, ["--gpus", "2"])


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
@pytest.mark.parametrize("name", ["winsize_multi", "auto_winsize", "auto_winsize_weighted"])
def test_cli_two_gpus_window_size_search(name):
    """The window-size drivers (--winsize-multi … --auto-winsize, --auto-winsize alone, --auto-winsize --weighted)
    over two GPUs: with the reproducible KDE (--kde-direct) every output file equals the one-GPU run's, so the
    thinned-window all-gather, the per-size passes and the sharded ROH calling follow the same path."""
    import os
    import tempfile
    from tests.test_cli_gpu import run_cli
    outs = []
    for extra in (["--kde-direct"], ["--kde-direct", "--gpus", "2"]):
        with tempfile.TemporaryDirectory() as tmp:
            ds, args, r = run_cli(name, tmp, extra=extra)
            assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
            files = {}
            for fn in sorted(os.listdir(tmp)):
                if fn.startswith("out.") and (fn.endswith(".roh.bed") or fn.endswith(".kde")):
                    files[fn] = open(os.path.join(tmp, fn)).read()
            log = [l for l in open(os.path.join(tmp, "out.log")).read().splitlines()[1:] if tmp not in l and "GPU" not in l]
            outs.append((files, log))
    assert outs[0][0].keys() == outs[1][0].keys() and any(k.endswith(".roh.bed") for k in outs[0][0])
    for k in outs[0][0]:
        assert outs[0][0][k] == outs[1][0][k], k
    assert outs[0][1] == outs[1][1]


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_cli_two_gpus_kde_on_device():
    """--kde-gpu over two GPUs: rank 0 runs computeKDE on the all-gathered thinned windows in its HBM; the .kde and the
    cutoff equal the one-GPU run's to rounding (the gathered matrix holds the same values, padded differently, so the
    fixed-order sums group them differently), the ROH are identical."""
    import os
    import tempfile
    from tests.test_cli_gpu import run_cli
    outs = []
    for extra in (["--kde-gpu"], ["--kde-gpu", "--gpus", "2"]):
        with tempfile.TemporaryDirectory() as tmp:
            ds, args, r = run_cli("auto_cutoff", tmp, extra=extra)
            assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
            cut = float(r.stdout.split("(17 digits): ")[1].split()[0])
            kde = np.loadtxt(os.path.join(tmp, [f for f in os.listdir(tmp) if f.endswith(".kde")][0]))
            outs.append((cut, kde, open(os.path.join(tmp, "out.roh.bed")).read()))
    assert abs(outs[0][0] - outs[1][0]) <= 1e-9 * max(1.0, abs(outs[0][0]))
    assert np.allclose(outs[0][1], outs[1][1], rtol=1e-6, atol=1e-12)
    assert outs[0][2] == outs[1][2]
