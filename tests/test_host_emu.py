"""CPU emulation of the product's walker / segment table / stitching (walk.cuh, segments.h compiled
for the host by tests/host_emu.cpp) against the literal oracle.  Guards the closed-form logic in a
container without a GPU; the CUDA kernels run the same source on the B200 (tests/test_gpu_*.py)."""
import ctypes as C
import math
import os
import subprocess

import numpy as np
import pytest

from oracle import oracle as orc
from tests.common import arg, flatten, load_case, oracle_roh_idx, oracle_windows_matrix

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
_EMU = None


def emu():
    global _EMU
    if _EMU is None:
        so = os.path.join(HERE, "_hostemu.so")
        srcs = [os.path.join(HERE, "host_emu.cpp")] + [os.path.join(ROOT, "garlic_b200", "csrc", f)
                                                       for f in ("walk.cuh", "segments.h", "common.cuh", "bound.cuh")]
        if not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
            subprocess.check_call(["g++", "-O2", "-ffp-contract=off", "-std=c++17", "-shared", "-fPIC",
                                   "-o", so, srcs[0]])
        _EMU = C.CDLL(so)
    return _EMU


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def thr_of(frac, W):
    t = frac * W
    t = t if t >= 1 else 1
    t = t if t <= W else W
    return int(math.ceil(t))


def emu_roh(F, W, cutoff, thr, chunk, tol, max_gap=200000, use_gl=False):
    cap = 1 << 16
    out = np.zeros((cap, 4), np.int32)
    n_amb = C.c_int(0)
    n_items = C.c_int(0)
    n = emu().emu_call_roh(_p(F["rows"]), C.c_int64(F["row_words"]), _p(F["lut"]), _p(F["gl"] if use_gl else None),
                           C.c_int64(F["L"] + 4160), _p(F["freq"]), C.c_int(F["N"]), C.c_int(len(F["chr_off"]) - 1),
                           _p(F["chr_off"]), _p(F["pos"]), _p(F["cen"]), C.c_int(max_gap), C.c_int(W),
                           C.c_double(cutoff), C.c_int(thr), C.c_int(chunk), C.c_double(tol), _p(out), C.c_int(cap),
                           C.byref(n_amb), C.byref(n_items))
    return [tuple(int(v) for v in r) for r in out[:n]], n_amb.value, n_items.value


def emu_windows(F, W, chunk, step, max_gap=200000, use_gl=False):
    nchr = len(F["chr_off"]) - 1
    slots = int(sum((F["chr_off"][c + 1] - F["chr_off"][c] + step - 1) // step for c in range(nchr)))
    out = np.full((F["N"], slots), orc.MISSING)
    emu().emu_windows(_p(F["rows"]), C.c_int64(F["row_words"]), _p(F["lut"]), _p(F["gl"] if use_gl else None),
                      C.c_int64(F["L"] + 4160), _p(F["freq"]), C.c_int(F["N"]), C.c_int(nchr), _p(F["chr_off"]),
                      _p(F["pos"]), _p(F["cen"]), C.c_int(max_gap), C.c_int(W), C.c_int(chunk), C.c_int(step),
                      _p(out), C.c_int64(slots))
    return out


CASES = ["lod_0", "lod_1", "lod_2", "lod_3", "lod_small", "auto_overlap_hg19"]


@pytest.mark.parametrize("name", CASES)
@pytest.mark.parametrize("chunk", [0, 256])
def test_roh_matches_oracle(name, chunk):
    ds, args = load_case(name)
    W = arg(args, "--winsize", cast=int)
    err = arg(args, "--error", cast=float)
    cutoff = arg(args, "--lod-cutoff", cast=float)
    ov = arg(args, "--overlap-frac", 0.25, float)
    res = orc.run_pipeline(ds, W, err, cutoff, ov, auto_overlap="--auto-overlap-frac" in args)
    F = flatten(res, err)
    got, n_amb, n_items = emu_roh(F, W, cutoff, thr_of(res["overlap_frac"], W), chunk, 0.0 if chunk == 0 else 1e-9)
    assert got == oracle_roh_idx(res)
    assert n_amb == 0


@pytest.mark.parametrize("name,W", [("lod_0", 25), ("lod_small", 50), ("lod_0", 70), ("lod_small", 33),
                                     ("lod_small", 32), ("lod_0", 200)])
def test_windows_exact_chain_bitwise(name, W):
    """Whole-segment items reproduce the reference's running sums bit for bit (same LUT)."""
    ds, args = load_case(name)
    res = orc.run_pipeline(ds, W, 0.002, None)
    F = flatten(res, 0.002)
    got = emu_windows(F, W, 0, 1)
    want = oracle_windows_matrix(res)
    assert got.shape == want.shape
    assert np.array_equal(got, want)


@pytest.mark.parametrize("W,chunk", [(25, 256), (50, 512), (10, 256)])
def test_windows_chunked_within_tolerance(W, chunk):
    ds, args = load_case("lod_0")
    res = orc.run_pipeline(ds, W, 0.002, None)
    F = flatten(res, 0.002)
    got = emu_windows(F, W, chunk, 1)
    want = oracle_windows_matrix(res)
    assert np.array_equal(got == orc.MISSING, want == orc.MISSING)
    ok = want != orc.MISSING
    assert np.max(np.abs(got[ok] - want[ok])) < 1e-11


def test_thinned_layout():
    ds, args = load_case("lod_0")
    res = orc.run_pipeline(ds, 25, 0.002, None)
    F = flatten(res, 0.002)
    got = emu_windows(F, 25, 0, 25)
    want = oracle_windows_matrix(res, 25)
    assert np.array_equal(got, want)


def test_gl_mode_roh_and_windows():
    ds, args = load_case("gl_pl")
    W = 30
    res = orc.run_pipeline(ds, W, None, 1.0, 0.25)
    F = flatten(res, None, use_gl=True)
    got, n_amb, _ = emu_roh(F, W, 1.0, thr_of(0.25, W), 0, 0.0, use_gl=True)
    assert got == oracle_roh_idx(res)
    w = emu_windows(F, W, 0, 1, use_gl=True)
    assert np.array_equal(w, oracle_windows_matrix(res))


def test_ambiguity_detection_and_cutoff_on_window_value():
    """A cutoff placed exactly on a window value must be flagged by the chunked pass."""
    ds, args = load_case("lod_small")
    W = 50
    res = orc.run_pipeline(ds, W, 0.001, None)
    win = res["chroms"][0]["win"]
    vals = win[win != orc.MISSING]
    cutoff = float(np.sort(vals)[len(vals) // 2])
    res = orc.run_pipeline(ds, W, 0.001, cutoff)
    F = flatten(res, 0.001)
    got, n_amb, _ = emu_roh(F, W, cutoff, thr_of(0.25, W), 256, 1e-9)
    assert n_amb >= 1
    got_exact, n_amb2, _ = emu_roh(F, W, cutoff, thr_of(0.25, W), 0, 0.0)
    assert got_exact == oracle_roh_idx(res)


def emu_piece_bounds(F, W):
    n_pieces = (F["L"] + 255) // 256
    pmax = np.zeros((n_pieces, F["N"]), np.uint32)
    rc = emu().emu_bound(_p(F["rows"]), C.c_int64(F["row_words"]), _p(F["lut"]), C.c_longlong(F["L"]), C.c_int(F["N"]),
                         C.c_int(W), _p(pmax), C.c_int(n_pieces))
    assert rc == 0
    return pmax


@pytest.mark.parametrize("name,item_pieces,W_over", [("lod_2", 1, None), ("lod_small", 1, None), ("auto_overlap_hg19", 2, None), ("lod_2", 4, None),
                                                     ("lod_small", 1, 16), ("lod_small", 1, 17), ("lod_small", 1, 24), ("lod_small", 1, 30),
                                                     ("lod_small", 2, 31), ("lod_small", 1, 250), ("lod_small", 1, 300), ("lod_2", 2, 530)])
def test_pruning_bound_never_drops_a_flagged_window(name, item_pieces, W_over):
    """bound.cuh: every (individual, item) pair that holds a window >= cutoff - tol must survive the pruning
    bound; and the bound must actually prune something on data with planted ROH."""
    ds, args = load_case(name)
    W = W_over or arg(args, "--winsize", cast=int)
    err = arg(args, "--error", cast=float)
    assert W >= 16
    res = orc.run_pipeline(ds, W, err, None, cm="--cm" in args)
    F = flatten(res, err)
    win = oracle_windows_matrix(res)
    vals = np.sort(win[win != orc.MISSING])
    pmax = emu_piece_bounds(F, W)
    for cutoff in (float(vals[int(0.9 * len(vals))]), float(vals[int(0.5 * len(vals))]), 2.0, float(vals[-1])):
        cap = 8192
        out = np.zeros((cap, F["N"]), np.uint8)
        bounds = np.zeros((cap, 3), np.int32)
        n = emu().emu_select(_p(pmax), C.c_int(F["N"]), C.c_int(len(F["chr_off"]) - 1), _p(F["chr_off"]), _p(F["pos"]),
                             _p(F["cen"]), C.c_int(200000), C.c_int(W), C.c_double(cutoff), C.c_double(1e-9),
                             C.c_int(item_pieces), _p(out), _p(bounds), C.c_int(cap))
        assert n > 0
        dropped_flagged = 0
        for i in range(n):
            w0, own_hi, we = bounds[i]
            hi = min(own_hi, we)
            has = (win[:, w0:hi] >= cutoff - 1e-9) & (win[:, w0:hi] != orc.MISSING)
            has = has.any(axis=1)
            dropped_flagged += int((has & (out[i] == 0)).sum())
        assert dropped_flagged == 0
        if cutoff == float(vals[-1]):
            assert out[:n].mean() < 0.9      # something is pruned when almost nothing passes the cutoff


@pytest.mark.parametrize("W", [16, 17, 18, 23, 30, 31, 32, 33, 49, 50, 64, 100, 177, 208, 209, 210, 240, 400])
def test_pruning_bound_dominates_every_window_of_its_block(W):
    """The stored piece maximum is an upper bound (fixed point, >> 2 rounded up) of every window starting in the piece,
    for every window-size class the kernel is instantiated for; the tail maximum covers the piece's last C2 blocks."""
    ds, args = load_case("lod_small")
    err = 0.001
    res = orc.run_pipeline(ds, W, err, None)
    F = flatten(res, err)
    lut, codes, L, N = F["lut"], F["codes"], F["L"], F["N"]
    val = np.take_along_axis(np.broadcast_to(lut[None, :L], (N, L, 4)), codes[:, :, None].astype(np.int64), 2)[:, :, 0]
    cs = np.concatenate([np.zeros((N, 1)), np.cumsum(val, 1)], 1)
    win = cs[:, W:] - cs[:, :-W]                      # every window start, chromosome borders ignored (a superset)
    pmax = emu_piece_bounds(F, W)
    all_ = (pmax & 0xffff).astype(np.uint16).view(np.int16).astype(np.int64) * 4 / 256.0
    tail = (pmax >> 16).astype(np.uint16).view(np.int16).astype(np.int64) * 4 / 256.0
    c2 = (min(W, 209) + 14) >> 4
    for p in range(pmax.shape[0]):
        lo, hi = 256 * p, min(256 * (p + 1), win.shape[1])
        if lo >= hi:
            continue
        assert np.all(win[:, lo:hi].max(1) <= all_[p] + 1e-9), (W, p)
        tl = 256 * (p + 1) - 16 * c2
        if tl < hi:
            assert np.all(win[:, max(lo, tl):hi].max(1) <= tail[p] + 1e-9), (W, p)


def test_compaction_plan_equals_column_gather():
    """bound.cuh:plan_half (K3 as funnel shifts / masked segments of 16-SNP half-words) == dropping the columns."""
    rng = np.random.default_rng(5)
    N, L0 = 37, 5000
    codes = rng.integers(0, 4, (N, L0)).astype(np.uint8)
    for drop in (0.0, 0.02, 0.5, 0.97):
        keep = rng.random(L0) >= drop
        keep[100:400] = drop < 0.9            # a long kept run and, with heavy dropping, a long dropped one
        keep[1000:1700] = False
        src = np.flatnonzero(keep).astype(np.int32)
        L = len(src)
        in_words = (((L0 + 4160 + 31) >> 5) + 3) & ~1
        rows_in = synth_pack(codes, in_words)
        out_words = (((L + 4160 + 31) >> 5) + 3) & ~1
        rows_out = np.full((N, out_words), 0xFFFFFFFFFFFFFFFF, np.uint64)
        emu().emu_squeeze(_p(rows_in), C.c_int64(in_words), _p(src), C.c_longlong(L), C.c_int(N), _p(rows_out),
                          C.c_int64(out_words))
        want = synth_pack(codes[:, keep], out_words)
        assert np.array_equal(rows_out, want), drop


def synth_pack(codes, row_words):
    from garlic_b200 import synth
    return synth.pack_codes(codes, row_bytes=row_words * 8).view(np.uint64).reshape(codes.shape[0], row_words)
