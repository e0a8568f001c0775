"""Sharding by individual across the GPUs of one box.

One process per GPU; rank r owns a contiguous block of individuals.  The only data-path exchanges are
  * a SUM all-reduce of the per-SNP counters int32[4][L0] (nalleles, total, hom, nonmiss) and, for text ingest,
    a MIN all-reduce of the first-allele keys,
  * a small all-gather of the thinned KDE windows,
and ROH lists are concatenated in rank order (= individual order, the order of the reference's BED tracks,
garlic-roh.cpp:426-431).  The functions take torch tensors so the same code runs over NCCL (device buffers of the
library, see `dev_tensor`) and over gloo on CPU (tests/test_multi_rank.py)."""
from __future__ import annotations

import numpy as np


def shard_range(n_total: int, world: int, rank: int):
    """Contiguous block [lo, hi) of individuals of `rank`: ceil(n/world) per rank, last ranks may be short."""
    per = -(-n_total // world)
    lo = min(n_total, rank * per)
    return lo, min(n_total, lo + per)


def split_individuals(global_idx, n_total: int, world: int):
    """Global individual indices (ascending, as gsl_ran_choose returns them) → per-rank LOCAL index arrays."""
    g = np.asarray(global_idx, np.int64)
    out = []
    for r in range(world):
        lo, hi = shard_range(n_total, world, r)
        out.append((g[(g >= lo) & (g < hi)] - lo).astype(np.int32))
    return out


class DevArray:
    """__cuda_array_interface__ view of a raw device pointer owned by libgarlic_b200."""

    def __init__(self, ptr, shape, typestr):
        self.__cuda_array_interface__ = dict(shape=tuple(shape), typestr=typestr, data=(int(ptr), False), version=2)


def dev_tensor(torch, ptr, shape, typestr, device):
    return torch.as_tensor(DevArray(ptr, shape, typestr), device=device)


def allreduce_counts(dist, counts):
    """counts: int32[4, L0] tensor (device buffer of the library under NCCL, CPU tensor under gloo)."""
    if dist is not None and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(counts, op=dist.ReduceOp.SUM)
    return counts


def allreduce_first_allele_keys(dist, keys):
    """keys: int64 view of the uint64 first-allele keys (all < 2^63, so signed MIN is the unsigned MIN)."""
    if dist is not None and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(keys, op=dist.ReduceOp.MIN)
    return keys


def allgather_thinned(torch, dist, mine, rows_per_rank: int, missing=-9999.0):
    """mine: float64[k, slots] windows of this rank's KDE individuals (k ≤ rows_per_rank).  Returns the
    float64[world*rows_per_rank, slots] stack (MISSING-padded), identical on every rank."""
    slots = mine.shape[1]
    pad = torch.full((rows_per_rank, slots), missing, dtype=torch.float64, device=mine.device)
    if mine.shape[0]:
        pad[:mine.shape[0]] = mine
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return pad
    out = torch.empty((dist.get_world_size() * rows_per_rank, slots), dtype=torch.float64, device=mine.device)
    dist.all_gather_into_tensor(out, pad)
    return out


def merge_roh(per_rank, n_total: int, world: int):
    """per_rank[r]: int32[n_r, 4] rows (local ind, chr, start_idx, stop_idx) sorted by (ind, chr, start).
    → one array with GLOBAL individual indices in the reference's (ind, chr, position) order."""
    parts = []
    for r, a in enumerate(per_rank):
        a = np.asarray(a, np.int32).reshape(-1, 4).copy()
        a[:, 0] += shard_range(n_total, world, r)[0]
        parts.append(a)
    return np.concatenate(parts, axis=0) if parts else np.empty((0, 4), np.int32)


def ld_planes_local(codes_local, ld_individuals, lo):
    """What ld_planes_kernel builds on one rank (wlod.cu): bit j of word j>>6 of the (non-missing, homozygous) planes of
    every SNP, set only for the LD individuals this rank holds (global indices in ld_individuals, rank holds [lo, lo+n)).
    codes_local: uint8[n_local, L] genotype codes.  → int64[L, 2, nw] (the all-reduce payload, as signed words)."""
    n_local, L = codes_local.shape
    nw = (len(ld_individuals) + 63) // 64
    planes = np.zeros((L, 2, nw), np.uint64)
    for j, g in enumerate(ld_individuals):
        k = int(g) - lo
        if 0 <= k < n_local:
            c = codes_local[k]
            bit = np.uint64(1) << np.uint64(j & 63)
            planes[:, 0, j >> 6] |= np.where(c != 3, bit, np.uint64(0))
            planes[:, 1, j >> 6] |= np.where((c == 0) | (c == 2), bit, np.uint64(0))
    return planes.view(np.int64)


def allreduce_ld_planes(dist, planes):
    """The weighted path's exchange (garlic_gpu_ld_band): ranks own disjoint bits, so SUM is OR."""
    dist.all_reduce(planes, op=dist.ReduceOp.SUM)
    return planes


def xchg_slice(n_loci, world, rank):
    """SNP slice rank `rank` owns in the counter exchange (csrc/xchg.cu: lo = L0*rank/N, hi = L0*(rank+1)/N)."""
    return (n_loci * rank) // world, (n_loci * (rank + 1)) // world


def exchange_counts_freq_keep(torch, dist, counts):
    """The algorithm of xchg_freq_keep_kernel with torch.distributed carrying the bytes (tests, world_size > 1 on CPU):
    every rank sums the two counter rows of ALL ranks over the slice it owns, evaluates freq = nalleles / total and the
    keep predicate there, and every rank ends up with the summed rows (in place), freq[] and keep[] of the whole SNP axis.
    counts: int32[>=2][L0] local counters (rows 0, 1 are exchanged).  -> (freq float64[L0], keep bool[L0])"""
    world, rank = dist.get_world_size(), dist.get_rank()
    L0 = counts.shape[1]
    theirs = [torch.empty_like(counts[:2]) for _ in range(world)]
    dist.all_gather(theirs, counts[:2].contiguous())          # what the peer pointers give the kernel
    lo, hi = xchg_slice(L0, world, rank)
    na = sum(t[0, lo:hi].to(torch.int64) for t in theirs)
    tot = sum(t[1, lo:hi].to(torch.int64) for t in theirs)
    f = torch.where(tot == 0, torch.zeros(hi - lo, dtype=torch.float64), na.to(torch.float64) / tot.to(torch.float64))
    k = (f > 0) & (f < 1)
    mine = torch.stack([na.to(torch.float64), tot.to(torch.float64), f, k.to(torch.float64)])   # this rank's slice
    sizes = [xchg_slice(L0, world, r)[1] - xchg_slice(L0, world, r)[0] for r in range(world)]
    width = max(sizes)
    pad = torch.zeros((4, width), dtype=torch.float64)
    pad[:, :hi - lo] = mine
    parts = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(parts, pad)                               # the stores into every rank's block
    freq = torch.empty(L0, dtype=torch.float64)
    keep = torch.empty(L0, dtype=torch.bool)
    for r in range(world):
        a, b = xchg_slice(L0, world, r)
        counts[0, a:b] = parts[r][0, :b - a].to(counts.dtype)
        counts[1, a:b] = parts[r][1, :b - a].to(counts.dtype)
        freq[a:b] = parts[r][2, :b - a]
        keep[a:b] = parts[r][3, :b - a] != 0
    return freq, keep
