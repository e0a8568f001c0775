#!/usr/bin/env python3
"""Measure the non-headline BASELINE configs (C3 wLOD, C4 GL/PL, C5 multi-winsize shard) on one B200 at sizes that
fit one GPU comfortably, with a parity check of a few individuals against the reference functions
(oracle/_ref/ref_driver) or the C port.  Prints one JSON object per config; summarised in profiles/.

    python tools/run_configs.py [c3] [c4] [c5]
"""
import ctypes
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402
from garlic_b200 import synth  # noqa: E402
from garlic_b200.api import GarlicGPU  # noqa: E402
from oracle import oracle as orc  # noqa: E402


def timed(g, fn, reps=3):
    fn()
    g.sync()
    t = []
    for _ in range(reps):
        t0 = time.perf_counter()
        out = fn()
        g.sync()
        t.append((time.perf_counter() - t0) * 1e3)
    return out, float(np.median(t))


def setup(n_ind, L0, seed):
    dev = torch.device("cuda", 0)
    names, chr_off0, pos0, cens = synth.make_positions_genomewide(seed, L0)
    row_bytes = ((L0 + 3) // 4 + 15) // 16 * 16
    rows = bench.make_rows_torch(torch, dev, n_ind, L0, seed, 1000, row_bytes)
    torch.cuda.synchronize()                     # the library copies on its own stream
    torch.cuda.empty_cache()
    g = GarlicGPU(0)
    g.set_shape(n_ind, L0, chr_off0, pos0)
    g.put_packed_dev(rows.data_ptr(), row_bytes)
    g.count_packed()
    return g, rows, names, chr_off0, pos0, cens


def oracle_sample(rows, n_s, L0, keep, freq, pos0, chr_off0, names, cens):
    codes = bench.unpack_rows(rows[:n_s].cpu().numpy(), L0)
    return bench.cpu_chroms(codes, keep, freq, pos0, chr_off0, names, cens)


def bench_peak_hbm():
    try:
        return float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        return 6650.0


def units_of(chr_off, W, n):
    return n * int(sum(max(0, chr_off[c + 1] - chr_off[c] - W + 1) for c in range(len(chr_off) - 1)))


def c5(n_ind=20000, L0=600_000):
    """C5 shard: --winsize-multi 30 50 70 on the KDE subsample (pass 1 x3), then pass 2 at the selected size."""
    g, rows, names, chr_off0, pos0, cens = setup(n_ind, L0, 5)
    freq, keep, L = g.filter()
    cen_arr = np.array([cens["chr" + nm] for nm in names], np.int32)
    g.set_tables(0.001, 200000, cen_arr)
    kde = np.linspace(0, n_ind - 1, 20).astype(np.int32)
    res = dict(config="C5 shard: %d ind x %d SNPs, --winsize-multi 30 50 70" % (n_ind, L0), loci_used=int(L))
    for W in (30, 50, 70):
        _, ms = timed(g, lambda: g.windows(W, W, individuals=kde, exact=False))
        res["pass1_W%d_ms" % W] = ms
    kept = g.get_kept_index()
    chr_off = np.searchsorted(kept, chr_off0)
    for W in (30, 50, 70):
        roh, ms = timed(g, lambda: g.call_roh(W, 2.0, 0.25))
        st = g.last_stats()
        res["pass2_W%d" % W] = dict(ms=ms, kernel_ms=st["kernel_ms"], bound_ms=st["squeeze_ms"], roh=len(roh), units=st["units"],
                                    units_per_s_kernel=st["units"] / (st["kernel_ms"] / 1e3), ambiguous=st["ambiguous_pairs"],
                                    candidate_pair_fraction=(st["candidate_pairs"] / st["all_pairs"]) if st["candidate_pairs"] >= 0 else None)
    # parity of 8 individuals at W = 30 and W = 70 against the reference functions
    n_s = 8
    chroms = oracle_sample(rows, n_s, L0, keep.copy(), freq.copy(), pos0, chr_off0, names, cens)
    pos_k = pos0[keep]
    for W in (30, 70):
        cs = bench.CpuSample(chroms, n_s, W, 0.001, 2.0, 0.25, 200000, 1)
        _, roh_cpu = cs.run()
        roh = g.call_roh(W, 2.0, 0.25)
        got = sorted((int(r[0]), int(r[1]), int(pos_k[r[2]]), int(pos_k[r[3]])) for r in roh if r[0] < n_s)
        res["parity_8_individuals_W%d" % W] = "identical ROH (%d), cpu kind=%s" % (len(got), cs.kind) if got == sorted(roh_cpu) else "MISMATCH"
    g.close()
    return res


def c4(n_ind=500, L0=2_000_000):
    """C4 at reduced L: per-genotype likelihoods (--tgls --gl-type PL), --winsize 200."""
    g, rows, names, chr_off0, pos0, cens = setup(n_ind, L0, 4)
    dev = torch.device("cuda", 0)
    gen = torch.Generator(device=dev)
    gen.manual_seed(44)
    vals = torch.tensor([0.0004, 0.004, 0.04, 0.5, 3.0, 0.0, 150.0], dtype=torch.float64, device=dev)
    gl = torch.empty((n_ind, L0), dtype=torch.float64, device=dev)         # C4: 40 GB; filled 25 individuals at a time
    for i0 in range(0, n_ind, 25):
        n = min(25, n_ind - i0)
        idx = torch.randint(0, 7, (n, L0), generator=gen, device=dev)      # (multinomial is not reproducible at this size)
        idx = torch.where(idx >= 5, torch.randint(0, 7, (n, L0), generator=gen, device=dev), idx)
        gl[i0:i0 + n] = vals[idx]
    del idx
    torch.cuda.synchronize()
    torch.cuda.empty_cache()
    g._ck(g.lib.garlic_gpu_put_gl_dev(g.h, ctypes.c_void_p(gl.data_ptr()), 2))
    n_s = min(16 if L0 > 4_000_000 else 40, n_ind)
    gl_sample = gl[:n_s].cpu().numpy()
    del gl
    torch.cuda.empty_cache()                                               # the library holds its own copy now
    (freq, keep, L), ms_f = timed(g, lambda: g.filter(), reps=1)
    cen_arr = np.array([cens["chr" + nm] for nm in names], np.int32)
    g.set_tables(None, 200000, cen_arr)
    W = 200
    res = dict(config="C4%s: %d ind x %d SNPs, PL likelihoods, W=%d" % ("" if L0 >= 10_000_000 else " reduced", n_ind, L0, W), loci_used=int(L),
               filter_ms=ms_f)
    kde = np.linspace(0, n_ind - 1, 20).astype(np.int32)
    _, ms = timed(g, lambda: g.windows(W, W, individuals=kde, exact=False))
    res["pass1_ms"] = ms
    roh, ms = timed(g, lambda: g.call_roh(W, 5.0, 0.25))
    st = g.last_stats()
    bytes_unit = 8.25
    res["pass2"] = dict(ms=ms, kernel_ms=st["kernel_ms"], roh=len(roh), units=st["units"],
                        units_per_s_kernel=st["units"] / (st["kernel_ms"] / 1e3),
                        hbm_gbs_algorithmic=st["units"] * bytes_unit / (st["kernel_ms"] / 1e3) / 1e9,
                        hbm_gbs_at_8_bytes=st["units"] * 8.0 / (st["kernel_ms"] / 1e3) / 1e9, ambiguous=st["ambiguous_pairs"])
    peak = bench_peak_hbm()
    res["pass2"]["roofline"] = dict(bound="hbm", unit="GB/s", achieved=res["pass2"]["hbm_gbs_at_8_bytes"], peak=peak,
                                    frac=res["pass2"]["hbm_gbs_at_8_bytes"] / peak,
                                    algorithmic="8 bytes (one fp64 per-genotype LOD) per individual-window",
                                    peak_source="MEASURED_PEAKS.json hbm_gbs (burst copy)")
    codes = bench.unpack_rows(rows[:n_s].cpu().numpy(), L0)
    chroms = bench.cpu_chroms(codes, keep.copy(), freq.copy(), pos0, chr_off0, names, cens)
    glh = orc.gl_error(gl_sample, "PL")
    for c, ch in enumerate(chroms):
        lo, hi = int(chr_off0[c]), int(chr_off0[c + 1])
        ch["gl"] = np.ascontiguousarray(glh[:, lo:hi][:, keep[lo:hi]].T)
    from oracle import refdrv
    pos_k = pos0[keep]
    if refdrv.available():
        out = refdrv.run(chroms, n_s, W, None, cutoff=5.0, overlap_frac=0.25, dump_windows=False)
        want = sorted((r[0], r[1], r[2], r[3]) for r in out["roh"])
        got = sorted((int(r[0]), int(r[1]), int(pos_k[r[2]]), int(pos_k[r[3]])) for r in roh if r[0] < n_s)
        res["parity_%d_individuals" % n_s] = "identical ROH (%d) vs reference functions" % len(got) if got == want else "MISMATCH %d vs %d" % (len(got), len(want))
    g.close()
    return res


FP64_DMMA_PEAK = 37.1   # TFLOP/s, mma.sync.m8n8k4.f64 on this pool's B200 (tools/fp64_peak.cu, profiles/r01m)


def c3(n_ind=5000, L0=200_000, n_ld=500, exact_check=True):
    """C3 at reduced L: --weighted wLOD with hr2 LD band over an LD subsample, --cm, W=72."""
    g, rows, names, chr_off0, pos0, cens = setup(n_ind, L0, 3)
    cen_arr = np.array([cens["chr" + nm] for nm in names], np.int32)
    C = len(names)
    map_pos = [pos0[chr_off0[c]:chr_off0[c + 1]][2:-2:5].astype(np.int64) for c in range(C)]
    map_cm = [np.round(p * 1.2e-6, 9) for p in map_pos]
    chr_param = np.array([[map_pos[c][0], map_pos[c][-1], cen_arr[c][0], cen_arr[c][1]] for c in range(C)], np.int32)
    freq, keep, L = g.filter(True, chr_param)
    kept = g.get_kept_index()
    pos = pos0[kept]
    chr_off = np.searchsorted(kept, chr_off0)
    from garlic_b200.pipeline import interpolate_map
    gpos = np.empty(L)
    for c in range(C):
        gpos[chr_off[c]:chr_off[c + 1]], _ = interpolate_map(pos[chr_off[c]:chr_off[c + 1]], map_pos[c], map_cm[c])
    g.set_tables(0.001, 200000, cen_arr, gpos)
    g.set_wlod(1e-9, 7)
    W = 72
    ld_ind = np.sort(np.random.default_rng(3).choice(n_ind, n_ld, replace=False)).astype(np.int32)
    _, ms_ld = timed(g, lambda: g.ld_band(W, ld_ind), reps=2)
    res = dict(config="C3%s: %d ind x %d SNPs, --weighted --cm --ld-subsample %d, W=%d" % ("" if L0 >= 1_000_000 else " reduced", n_ind, L0, n_ld, W),
               loci_used=int(L), ld_band_ms=ms_ld, ld_pair_evaluations=float(L) * W * W * n_ld)
    roh, ms = timed(g, lambda: g.call_roh(W, 1.0, 0.25, weighted=True), reps=2)
    st = g.last_stats()
    res["pass2_wlod"] = dict(ms=ms, kernel_ms=st["kernel_ms"], roh=len(roh), units=st["units"],
                             units_per_s_kernel=st["units"] / (st["kernel_ms"] / 1e3),
                             fp64_gflops=st["units"] * 2 * W / (st["kernel_ms"] / 1e3) / 1e9, ambiguous=st["ambiguous_pairs"])
    res["pass2_wlod"]["roofline"] = dict(bound="tensor", unit="TFLOP/s", achieved=res["pass2_wlod"]["fp64_gflops"] / 1e3, peak=FP64_DMMA_PEAK,
                                         frac=res["pass2_wlod"]["fp64_gflops"] / 1e3 / FP64_DMMA_PEAK,
                                         peak_source="fp64 mma.sync m8n8k4 peak measured on this pool by tools/fp64_peak.cu (no fp64 entry in MEASURED_PEAKS.json)")
    if exact_check:
        # the tensor-core pass against exact mul-then-add sums on the same handle (whole result, every individual)
        roh_exact, ms_x = timed(g, lambda: g.call_roh(W, 1.0, 0.25, weighted=True, exact=True), reps=1)
        res["exact_kernel_ms"] = g.last_stats()["kernel_ms"]
        res["parity_all_individuals"] = "tensor-core pass == exact sums (%d ROH)" % len(roh) if np.array_equal(roh, roh_exact) else "MISMATCH"
    # … and against the REFERENCE's own functions (calculateGenoFreq + calcHR2LD + calcwLODWindows + assembleROHWindows,
    # oracle/_ref/ref_driver) on the last chromosomes (as many as keep the reference's L·W²·n_LD pair loop at about a
    # minute of host time): every individual, the same LD individuals, bit-identical ROH required
    from oracle import refdrv
    if refdrv.available():
        budget, c0 = 30_000, C
        while c0 > 0 and (chr_off[C] - chr_off[c0 - 1]) <= budget:
            c0 -= 1
        c0 = min(c0, C - 1)
        chroms = []
        keep_h, freq_h = keep.copy(), freq.copy()
        for c in range(c0, C):
            lo0, hi0 = int(chr_off0[c]), int(chr_off0[c + 1])
            b0, b1 = lo0 // 4, (hi0 + 3) // 4
            sub = rows[:, b0:b1].cpu().numpy()
            codes = bench.unpack_rows(sub, (b1 - b0) * 4)[:, lo0 - 4 * b0:hi0 - 4 * b0]
            k = keep_h[lo0:hi0]
            lo, hi = int(chr_off[c]), int(chr_off[c + 1])
            chroms.append(dict(name="chr" + names[c], cen=cens["chr" + names[c]], pos=np.ascontiguousarray(pos[lo:hi]),
                               freq=np.ascontiguousarray(freq_h[lo0:hi0][k]), gpos=np.ascontiguousarray(gpos[lo:hi]), gl=None,
                               geno=np.ascontiguousarray(codes[:, k].T.astype(np.int8))))
        t0 = time.perf_counter()
        out = refdrv.run(chroms, n_ind, W, 0.001, cutoff=1.0, overlap_frac=0.25, weighted=True, cm=True, mu=1e-9, M=7,
                         threads=min(32, os.cpu_count() or 1), ld_individuals=ld_ind, dump_windows=False)
        want = sorted((r[0], r[1] + c0, r[2], r[3]) for r in out["roh"])
        got = sorted((int(r[0]), int(r[1]), int(pos[r[2]]), int(pos[r[3]])) for r in roh if r[1] >= c0)
        n_snps = int(chr_off[C] - chr_off[c0])
        res["parity_vs_reference_functions"] = (
            "identical ROH (%d) on chromosomes %s (%d SNPs), all %d individuals, LD over the same %d individuals; reference "
            "calcHR2LD + calcwLODWindows + assembleROHWindows took %.1f s on the host" % (
                len(got), "+".join(names[c0:]), n_snps, n_ind, n_ld, time.perf_counter() - t0)) if got == want else \
            "MISMATCH: gpu %d vs reference %d ROH on chromosomes %s" % (len(got), len(want), "+".join(names[c0:]))
    g.close()
    return res


if __name__ == "__main__":
    # e.g.  c4   c4:500:10000000   c3:5000:1000000   c5:12500:600000   (name[:individuals[:SNPs]])
    which = sys.argv[1:] or ["c5", "c4", "c3"]
    for w in which:
        parts = w.split(":")
        print(json.dumps({"c3": c3, "c4": c4, "c5": c5}[parts[0]](*[int(x) for x in parts[1:]])), flush=True)
