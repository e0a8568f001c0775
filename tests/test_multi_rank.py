"""world_size-2 gloo test of the individual-sharding logic (garlic_b200/shard.py): per-rank counters summed by
all-reduce equal the single-rank counters, the MIN all-reduce of first-allele keys picks the globally first
allele, thinned windows are gathered in rank order, and merged ROH equal the single-shard result.  The per-rank
"kernel outputs" are produced with the CPU oracle here (test infrastructure); on the GPU the same exchanges run
over NCCL on the library's device buffers (bench.py --gpus N)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from garlic_b200 import shard
from oracle import oracle as orc
from tests.common import load_case


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _first_keys(alleles, ind_offset, missing=ord("0")):
    """what first_allele_kernel computes per rank: min over local calls of (global call index << 8 | char)."""
    L0, N, _ = alleles.shape
    flat = alleles.reshape(L0, 2 * N)
    ok = flat != missing
    first = np.where(ok.any(1), ok.argmax(1), 0)
    ch = flat[np.arange(L0), first].astype(np.int64)
    key = ((2 * ind_offset + first).astype(np.int64) << 8) | ch
    return np.where(ok.any(1), key, np.iinfo(np.int64).max)


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    ds, args = load_case("lod_0")
    N = ds.n_ind
    lo, hi = shard.shard_range(N, world, rank)
    sub = ds.alleles[:, lo:hi]
    # phase a: first-allele keys, MIN all-reduce
    keys = torch.from_numpy(_first_keys(sub, lo))
    shard.allreduce_first_allele_keys(dist, keys)
    one = np.where(keys.numpy() == np.iinfo(np.int64).max, ord("0"), keys.numpy() & 0xFF).astype(np.uint8)
    # phase b: local coding against the global "1" allele, local counters, SUM all-reduce
    a1, a2 = sub[:, :, 0], sub[:, :, 1]
    miss = ord("0")
    na = ((a1 == one[:, None]) & (a1 != miss)).sum(1) + ((a2 == one[:, None]) & (a2 != miss)).sum(1)
    tot = (a1 != miss).sum(1) + (a2 != miss).sum(1)
    code = np.where((a1 == miss) | (a2 == miss), 3, (a1 == one[:, None]).astype(int) + (a2 == one[:, None]).astype(int))
    hom = ((code == 0) | (code == 2)).sum(1)
    nm = (code != 3).sum(1)
    counts = torch.from_numpy(np.stack([na, tot, hom, nm]).astype(np.int32))
    shard.allreduce_counts(dist, counts)
    # thinned windows of the KDE individuals this rank owns, gathered in rank order
    kde = np.array([1, 5, 6, 20, 29])
    mine_idx = shard.split_individuals(kde, N, world)[rank]
    res = orc.run_pipeline(ds, 25, 0.002, 1.5, 0.3)
    win = np.concatenate([c["win"] for c in res["chroms"]], axis=1)
    rows = max(len(x) for x in shard.split_individuals(kde, N, world))
    allv = shard.allgather_thinned(torch, dist, torch.from_numpy(win[lo:hi][mine_idx][:, ::25].copy()), rows)
    # ROH of this rank's individuals with LOCAL indices, as garlic_gpu_call_roh returns them
    roh = np.array([(r[0] - lo, r[1], r[5], r[6]) for r in res["roh"] if lo <= r[0] < hi], np.int32).reshape(-1, 4)
    gathered = [None] * world
    dist.all_gather_object(gathered, roh)
    # --weighted: bit-planes of the LD individuals (a list over the whole sample), SUM (= OR) all-reduce
    ld = np.array([0, 3, 4, 9, 10, 17, 20, 29])
    planes = torch.from_numpy(shard.ld_planes_local(code[:, :].T.astype(np.uint8).copy(), ld, lo))
    shard.allreduce_ld_planes(dist, planes)
    if rank == 0:
        q.put(dict(one=one, counts=counts.numpy(), thin=allv.numpy(), roh=shard.merge_roh(gathered, N, world),
                   planes=planes.numpy()))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_exchanges_equal_single_rank():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    out = q.get()
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    ds, args = load_case("lod_0")
    geno, na, tot, one, freq = orc.code_tped(ds.alleles)
    codes = geno.astype(np.uint8)
    assert np.array_equal(out["one"], one)
    assert np.array_equal(out["counts"][0], na) and np.array_equal(out["counts"][1], tot)
    assert np.array_equal(out["counts"][2], ((codes == 0) | (codes == 2)).sum(1))
    assert np.array_equal(out["counts"][3], (codes != 3).sum(1))
    res = orc.run_pipeline(ds, 25, 0.002, 1.5, 0.3)
    win = np.concatenate([c["win"] for c in res["chroms"]], axis=1)
    kde = np.array([1, 5, 6, 20, 29])
    got = out["thin"][~np.all(out["thin"] == -9999.0, axis=1)]
    assert np.array_equal(got, win[kde][:, ::25])
    want = np.array([(r[0], r[1], r[5], r[6]) for r in res["roh"]], np.int32).reshape(-1, 4)
    assert np.array_equal(out["roh"], want)
    # the reduced planes are the single-rank planes, and give the oracle's pair counts (hr2's total / HAB popcounts)
    ld = np.array([0, 3, 4, 9, 10, 17, 20, 29])
    whole = shard.ld_planes_local(codes.T.copy(), ld, 0)
    assert np.array_equal(out["planes"], whole)
    pl = out["planes"].view(np.uint64)
    i, j = 100, 117
    both = sum(bin(int(pl[i, 0, w] & pl[j, 0, w])).count("1") for w in range(pl.shape[2]))
    assert both == int(((codes[i, ld] != 3) & (codes[j, ld] != 3)).sum())


def test_shard_ranges_cover_everything():
    for n, w in [(2000, 8), (45, 8), (7, 8), (500, 3), (1, 2)]:
        seen = []
        for r in range(w):
            lo, hi = shard.shard_range(n, w, r)
            seen += list(range(lo, hi))
        assert seen == list(range(n))
    parts = shard.split_individuals([0, 3, 9, 10, 44], 45, 8)
    assert sum(len(p) for p in parts) == 5 and list(parts[0]) == [0, 3] and list(parts[1]) == [3, 4]


def _xchg_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    ds, args = load_case("lod_0")
    geno, na, tot, one, freq = orc.code_tped(ds.alleles)
    codes = geno.astype(np.uint8)                              # [snp][ind]
    lo, hi = shard.shard_range(ds.n_ind, world, rank)
    c = codes[:, lo:hi]
    sub = ds.alleles[:, lo:hi]                                 # allele counts per character, as K1 makes them (half-missing calls count)
    a1, a2, miss = sub[:, :, 0], sub[:, :, 1], ord("0")
    n_a = ((a1 == one[:, None]) & (a1 != miss)).sum(1) + ((a2 == one[:, None]) & (a2 != miss)).sum(1)
    n_t = (a1 != miss).sum(1) + (a2 != miss).sum(1)
    local = np.stack([n_a, n_t, ((c == 0) | (c == 2)).sum(1), (c != 3).sum(1)]).astype(np.int32)
    counts = torch.from_numpy(local.copy())
    f, k = shard.exchange_counts_freq_keep(torch, dist, counts)
    gathered = [None] * world
    dist.all_gather_object(gathered, (counts.numpy().copy(), f.numpy().copy(), k.numpy().copy()))
    if rank == 0:
        q.put(gathered)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_counter_exchange_by_owned_slices_equals_single_rank(world):
    """The algorithm of csrc/xchg.cu (the NVLink counter exchange fused with freq + keep): every rank owns a slice of the SNP
    axis, sums all ranks' counters there, evaluates freq / keep and hands sums, freq and keep to everybody.  Run over gloo:
    every rank ends with the single-shard counters (rows 0, 1 summed in place, rows 2, 3 untouched), the oracle's freq[] bit
    for bit and its keep mask; the slices cover the axis exactly once also when it does not divide (world = 3)."""
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    port = _free_port()
    procs = [ctx.Process(target=_xchg_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    out = q.get()
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    ds, args = load_case("lod_0")
    geno, na, tot, one, freq = orc.code_tped(ds.alleles)
    L0 = len(freq)
    seen = []
    for r in range(world):
        a, b = shard.xchg_slice(L0, world, r)
        seen += list(range(a, b))
    assert seen == list(range(L0))
    keep = orc.keep_mask(freq, np.asarray(ds.pos))
    for r, (counts, f, k) in enumerate(out):
        assert np.array_equal(counts[0], na) and np.array_equal(counts[1], tot), r
        assert np.array_equal(f, freq) and np.array_equal(k, keep.astype(bool)), r
    # rows 2, 3 stay local (they are summed by NCCL only when the LD band asks)
    codes = geno.astype(np.uint8)
    lo, hi = shard.shard_range(ds.n_ind, world, 1)
    assert np.array_equal(out[1][0][3], (codes[:, lo:hi] != 3).sum(1))
