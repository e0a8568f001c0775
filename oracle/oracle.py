"""ctypes wrapper + pipeline glue around oracle/liboracle.so (garlic_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  Nothing under garlic_b200/ imports this.

``run_pipeline`` restates the order of operations of the reference's main()
(garlic-main.cpp:216-406) for the hot path: code → freq → filter → [map interpolation, homFreq,
LD] → windows → [thinning] → assembly, per chromosome exactly as the reference loops.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None
MISSING = -9999.0


def lib():
    global _LIB
    if _LIB is None:
        so = os.path.join(_HERE, "liboracle.so")
        src = os.path.join(_HERE, "garlic_oracle.c")
        if (not os.path.exists(so)) or os.path.getmtime(so) < os.path.getmtime(src):
            subprocess.check_call(["make", "-C", _HERE, "liboracle.so"], stdout=subprocess.DEVNULL)
        L = C.CDLL(so)
        L.orc_lod.restype = C.c_double
        L.orc_lod.argtypes = [C.c_int, C.c_double, C.c_double]
        L.orc_gl_error.restype = C.c_double
        L.orc_gl_error.argtypes = [C.c_double, C.c_int]
        L.orc_thin.restype = C.c_size_t
        L.orc_assemble.restype = C.c_int
        L.orc_interpolate_map.restype = C.c_int
        L.orc_select_overlap_frac.restype = C.c_double
        L.orc_select_overlap_frac.argtypes = [C.c_double, C.c_int]
        L.orc_select_winsize_weighted.restype = C.c_int
        L.orc_select_winsize_weighted.argtypes = [C.c_double]
        L.orc_in_gap.restype = C.c_int
        _LIB = L
    return _LIB


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


GL_TYPES = {"GQ": 0, "GL": 1, "PL": 2}


def code_tped(alleles: np.ndarray, missing="0"):
    """alleles uint8[L0,N,2] → geno int8[L0,N], nalleles, total, one_allele, freq."""
    L0, N, _ = alleles.shape
    alleles = np.ascontiguousarray(alleles, np.uint8)
    geno = np.empty((L0, N), np.int8)
    na = np.empty(L0, np.int32)
    tot = np.empty(L0, np.int32)
    one = np.empty(L0, np.uint8)
    freq = np.empty(L0, np.float64)
    lib().orc_code_tped(_p(alleles), C.c_int(L0), C.c_int(N), C.c_int(ord(missing)),
                        _p(geno), _p(na), _p(tot), _p(one), _p(freq))
    return geno, na, tot, one, freq


def hom_freq(geno):
    L, N = geno.shape
    out = np.empty(L, np.float64)
    with np.errstate(all="ignore"):
        lib().orc_hom_freq(_p(np.ascontiguousarray(geno)), C.c_int(L), C.c_int(N), _p(out))
    return out


def keep_mask(freq, pos, oob=False, scaf_first=0, scaf_last=0, cen=(0, 0)):
    L = len(freq)
    keep = np.empty(L, np.uint8)
    lib().orc_keep_mask(_p(np.ascontiguousarray(freq)), _p(np.ascontiguousarray(pos, np.int32)),
                        C.c_int(L), C.c_int(int(oob)), C.c_int(int(scaf_first)), C.c_int(int(scaf_last)),
                        C.c_int(int(cen[0])), C.c_int(int(cen[1])), _p(keep))
    return keep.astype(bool)


def gl_error(values, gl_type):
    v = np.ascontiguousarray(values, np.float64)
    out = np.empty_like(v)
    lib().orc_gl_error_array(_p(v), C.c_size_t(v.size), C.c_int(GL_TYPES[gl_type]), _p(out))
    return out


def lod_lut(freq, error):
    L = len(freq)
    lut = np.empty((L, 4), np.float64)
    lib().orc_lod_lut(_p(np.ascontiguousarray(freq)), C.c_int(L), C.c_double(error), _p(lut))
    return lut


def lod_matrix(geno, freq, err):
    """per-genotype lod() for per-genotype error rates: geno int8[L,N], freq[L], err[L,N] → float64[L,N]."""
    L, N = geno.shape
    out = np.empty((L, N), np.float64)
    lib().orc_lod_matrix(_p(np.ascontiguousarray(geno, np.int8)), _p(np.ascontiguousarray(freq, np.float64)),
                         _p(np.ascontiguousarray(err, np.float64)), C.c_int(L), C.c_int(N), _p(out))
    return out


def calc_lod(geno, freq, pos, W, error, max_gap, cen, gl=None):
    L, N = geno.shape
    win = np.empty((N, L), np.float64)
    lib().orc_calc_lod(_p(np.ascontiguousarray(geno)), _p(np.ascontiguousarray(freq)),
                       _p(np.ascontiguousarray(pos, np.int32)), C.c_int(L), C.c_int(N), C.c_int(W),
                       C.c_double(error), C.c_int(max_gap), C.c_int(cen[0]), C.c_int(cen[1]),
                       _p(None if gl is None else np.ascontiguousarray(gl)), _p(win))
    return win


def calc_hr2_ld(geno, homf, W, ind_index):
    L, N = geno.shape
    LD = np.empty((L, W), np.float64)
    idx = np.ascontiguousarray(ind_index, np.int32)
    lib().orc_calc_hr2_ld(_p(np.ascontiguousarray(geno)), _p(np.ascontiguousarray(homf)), C.c_int(L),
                          C.c_int(N), C.c_int(W), _p(idx), C.c_int(len(idx)), _p(LD))
    return LD


def calc_r2_ld(geno, first_copy, freq, W, ind_index):
    """--phased: r2 between haplotypes (garlic-data.cpp:426-471,585-617); first_copy uint8[L,N]."""
    L, N = geno.shape
    LD = np.empty((L, W), np.float64)
    idx = np.ascontiguousarray(ind_index, np.int32)
    lib().orc_calc_r2_ld(_p(np.ascontiguousarray(geno)), _p(np.ascontiguousarray(first_copy, np.uint8)),
                         _p(np.ascontiguousarray(freq)), C.c_int(L), C.c_int(N), C.c_int(W), _p(idx), C.c_int(len(idx)), _p(LD))
    return LD


def wlod_weights(pos, gpos, mu, M):
    L = len(pos)
    a = np.empty(L)
    b = np.empty(L)
    lib().orc_wlod_weights(_p(np.ascontiguousarray(pos, np.int32)), _p(np.ascontiguousarray(gpos)),
                           C.c_int(L), C.c_double(mu), C.c_int(M), _p(a), _p(b))
    return a, b


def calc_wlod(geno, freq, pos, gpos, W, error, max_gap, cen, LD, mu, M, gl=None):
    L, N = geno.shape
    win = np.empty((N, L), np.float64)
    with np.errstate(all="ignore"):
        lib().orc_calc_wlod(_p(np.ascontiguousarray(geno)), _p(np.ascontiguousarray(freq)),
                            _p(np.ascontiguousarray(pos, np.int32)), _p(np.ascontiguousarray(gpos)),
                            C.c_int(L), C.c_int(N), C.c_int(W), C.c_double(error), C.c_int(max_gap),
                            C.c_int(cen[0]), C.c_int(cen[1]),
                            _p(None if gl is None else np.ascontiguousarray(gl)),
                            _p(np.ascontiguousarray(LD)), C.c_double(mu), C.c_int(M), _p(win))
    return win


def thin(win, ind_list, step):
    N, L = win.shape
    idx = np.ascontiguousarray(ind_list, np.int32)
    out = np.empty(len(idx) * ((L + step - 1) // step), np.float64)
    n = lib().orc_thin(_p(np.ascontiguousarray(win)), C.c_int(L), _p(idx), C.c_int(len(idx)),
                       C.c_int(step), _p(out))
    return out[:n].copy()


def assemble(win_row, pos, gpos, cutoff, W, max_gap, overlap_frac, cm, cen):
    L = len(win_row)
    cap = L // 2 + 2
    a = np.empty(cap, np.int32)
    b = np.empty(cap, np.int32)
    ln = np.empty(cap, np.float64)
    n = lib().orc_assemble(_p(np.ascontiguousarray(win_row)), _p(np.ascontiguousarray(pos, np.int32)),
                           _p(None if gpos is None else np.ascontiguousarray(gpos)), C.c_int(L),
                           C.c_double(cutoff), C.c_int(W), C.c_int(max_gap), C.c_double(overlap_frac),
                           C.c_int(int(cm)), C.c_int(cen[0]), C.c_int(cen[1]), _p(a), _p(b), _p(ln),
                           C.c_int(cap))
    return a[:n].copy(), b[:n].copy(), ln[:n].copy()


def interpolate_map(pos, spos, sgen):
    gpos = np.empty(len(pos), np.float64)
    n = lib().orc_interpolate_map(_p(np.ascontiguousarray(pos, np.int32)), C.c_int(len(pos)),
                                  _p(np.ascontiguousarray(spos, np.int32)),
                                  _p(np.ascontiguousarray(sgen, np.float64)), C.c_int(len(spos)), _p(gpos))
    if n < 0:
        raise ValueError("site outside of map scaffold")
    return gpos, n


def calc_density(n_loci, chroms):
    """garlic-data.cpp:318-328.  chroms: list of dicts with pos, cen."""
    length = 0.0
    for ch in chroms:
        length += int(ch["pos"][-1]) - int(ch["pos"][0]) + 1 - (ch["cen"][1] - ch["cen"][0])
    return float(n_loci) / length


def chr_label(name):
    """garlic-data.cpp:1886-1891 checkChrName."""
    return name if name[0] == "c" else "chr" + name


# ------------------------------------------------------------------------------------------------
def run_pipeline(ds, W, error=None, cutoff=None, overlap_frac=0.25, max_gap=200000, weighted=False,
                 cm=False, mu=1e-9, M=7, ld_individuals=None, kde_individuals=None, thin_step=None,
                 auto_overlap=False, keep_windows=True, phased=False):
    """Oracle run over a synth.Dataset.  Returns a dict with per-chromosome arrays and the ROH
    list [(ind, chr_index, start_bp, stop_bp, length)] in the reference's (ind, chr, pos) order."""
    geno0, na, tot, one, freq0 = code_tped(ds.alleles, ds.tped_missing)
    use_gl = ds.gl is not None
    gl0 = gl_error(ds.gl, ds.gl_type) if use_gl else None
    N = ds.n_ind
    chroms = []
    n_used = 0
    n_interp = 0
    for c, name in enumerate(ds.chr_names):
        lo, hi = int(ds.chr_offsets[c]), int(ds.chr_offsets[c + 1])
        label = chr_label(name)
        cen = ds.centromeres.get(label, (0, 0))
        pos = ds.pos[lo:hi]
        if weighted or cm:
            sp, sg = ds.map_pos[c], ds.map_cm[c]
            keep = keep_mask(freq0[lo:hi], pos, True, sp[0], sp[-1], cen)
        else:
            keep = keep_mask(freq0[lo:hi], pos)
        fc = None
        if phased:      # firstCopy[i] = (alleleStr1 == oneAllele), garlic-data.cpp:129
            fc = np.ascontiguousarray((ds.alleles[lo:hi, :, 0] == one[lo:hi, None])[keep]).astype(np.uint8)
        ch = dict(name=label, cen=cen, keep=keep, lo=lo, first_copy=fc, pos=np.ascontiguousarray(pos[keep]),
                  geno=np.ascontiguousarray(geno0[lo:hi][keep]), freq=np.ascontiguousarray(freq0[lo:hi][keep]),
                  gl=np.ascontiguousarray(gl0[lo:hi][keep]) if use_gl else None, gpos=None)
        if weighted or cm:
            ch["gpos"], k = interpolate_map(ch["pos"], sp, sg)
            n_interp += k
        if weighted:
            ch["homf"] = hom_freq(ch["geno"])
        n_used += int(keep.sum())
        chroms.append(ch)
    out = dict(freq0=freq0, nalleles=na, total=tot, one_allele=one, geno0=geno0, chroms=chroms,
               n_used=n_used, n_interp=n_interp, gl0=gl0)
    if auto_overlap:
        dens = calc_density(n_used, chroms)
        overlap_frac = float(lib().orc_select_overlap_frac(dens, W))
        out["density"] = dens
    out["overlap_frac"] = overlap_frac
    ldi = np.arange(N, dtype=np.int32) if ld_individuals is None else np.asarray(ld_individuals, np.int32)
    for ch in chroms:
        if weighted:
            ch["LD"] = calc_r2_ld(ch["geno"], ch["first_copy"], ch["freq"], W, ldi) if phased else \
                calc_hr2_ld(ch["geno"], ch["homf"], W, ldi)
            ch["win"] = calc_wlod(ch["geno"], ch["freq"], ch["pos"], ch["gpos"], W,
                                  -1.0 if error is None else error, max_gap, ch["cen"], ch["LD"], mu, M, ch["gl"])
        else:
            ch["win"] = calc_lod(ch["geno"], ch["freq"], ch["pos"], W, -1.0 if error is None else error,
                                 max_gap, ch["cen"], ch["gl"])
    if thin_step is not None:
        kdi = np.arange(N, dtype=np.int32) if kde_individuals is None else np.asarray(kde_individuals, np.int32)
        out["thinned"] = np.concatenate([thin(ch["win"], kdi, thin_step) for ch in chroms])
    if cutoff is not None:
        roh = []
        for i in range(N):
            for ci, ch in enumerate(chroms):
                a, b, ln = assemble(ch["win"][i], ch["pos"], ch["gpos"], cutoff, W, max_gap, overlap_frac,
                                    cm, ch["cen"])
                for k in range(len(a)):
                    roh.append((i, ci, int(ch["pos"][a[k]]), int(ch["pos"][b[k]]), float(ln[k]), int(a[k]), int(b[k])))
        out["roh"] = roh
    if not keep_windows:
        for ch in chroms:
            ch.pop("win", None)
    return out


COLORS = ["228,26,28", "77,175,74", "55,126,184", "152,78,163", "255,127,0", "255,255,51",
          "166,86,40", "247,129,191", "153,153,153"]



# ------------------------------------------------------------------------------------------------
# KDE and cutoff heuristic (garlic-kde.cpp), restated with the EXACT Gauss transform
# ------------------------------------------------------------------------------------------------
def compute_kde(data, M=512):
    """computeKDE (garlic-kde.cpp:14-101) with nrd0 (:130-140): sorted data, gsl_stats_sd (N-1), gsl quantiles at
    f(n-1) with linear interpolation, h = 0.9 min(sd, iqr/1.34) n^-0.2; M targets (i+1)/M (max-min)+min over
    [min-3h, max+3h]; y = sum_i (1/n) exp(-(t-x_i)^2/h^2) — the sum FIGTree evaluates (to eps = 1e-2 in the reference,
    exactly with FIGTREE_EVAL_DIRECT) — divided by sum(y)*spacing.  -> (t, y, h)"""
    x = np.sort(np.asarray(data, np.float64))
    n = len(x)
    sd = np.std(x, ddof=1)

    def q(f):
        idx = f * (n - 1)
        lo = int(idx)
        d = idx - lo
        return x[lo] if lo == n - 1 else (1 - d) * x[lo] + d * x[lo + 1]
    h = 0.9 * min(sd, (q(0.75) - q(0.25)) / 1.34) * n ** -0.2
    mn, mx = x[0] - 3 * h, x[-1] + 3 * h
    t = (np.arange(1, M + 1) / float(M)) * (mx - mn) + mn
    y = np.array([np.exp(-((ti - x) / h) ** 2).sum() / n for ti in t])
    y /= y.sum() * (t[1] - t[0])
    return t, y, h


def min_between_modes(t, y, W):
    """get_min_btw_modes (garlic-kde.cpp:142-234), behaviour for behaviour (including the i == 1 case)."""
    size, win = len(y), 20
    m = size - win
    umax, ucnt, index = np.zeros(m), np.zeros(m), 0
    for i in range(m):
        seg = y[i:i + win]
        mxv = seg[np.argmax(seg)] if seg.max() > np.finfo(float).tiny else y[i - 1]
        if i == 1:
            umax[i] = mxv
            ucnt[i] += 1
        elif umax[index] == mxv:
            ucnt[index] += 1
        else:
            index += 1
            umax[index] = mxv
            ucnt[index] += 1
    c1, c2 = int(ucnt[0]), 0
    for i in range(1, m):
        if c1 <= ucnt[i]:
            c2, c1 = c1, int(ucnt[i])
        elif c2 <= ucnt[i]:
            c2 = int(ucnt[i])
    vals = [umax[i] for i in range(m) if ucnt[i] == c1 or ucnt[i] == c2]
    first = second = -1.0
    for v in vals:
        if first <= v:
            second, first = first, v
        elif second <= v:
            second = v
    li = ri = -1
    for i in range(size):
        if y[i] == first:
            li = i
        if y[i] == second:
            ri = i
    if ri < li:
        li, ri = ri, li
    mi = int(np.argmin(y[li:ri + 1])) + li
    return t[mi] if abs(t[mi] / W) < 1 else 0.0


def _fmt_g6(x):
    """C++ ostream default formatting of a double (precision 6, %g)."""
    return "%g" % x


def format_bed(roh, ind_ids, chr_labels, bounds, pop, cm=False, version="1.1.6a"):
    """garlic-roh.cpp:574-644 writeROHData."""
    lines = []
    by_ind = {}
    for r in roh:
        by_ind.setdefault(r[0], []).append(r)
    for i, iid in enumerate(ind_ids):
        lines.append('track name="Ind: %s Pop:%s ROH" description="Ind: %s Pop:%s ROH from GARLIC v%s" '
                     'visibility=2 itemRgb="On"' % (iid, pop, iid, pop, version))
        for r in by_ind.get(i, []):
            size = r[4]
            k = 0
            while k < len(bounds) and not (size < bounds[k]):
                k += 1
            cls = chr(ord("A") + k)
            color = COLORS[k if k <= 8 else 8]
            c = chr_labels[r[1]]
            if c[0] not in "cC":
                c = "chr" + c
            sz = _fmt_g6(size) if cm else str(int(size))
            lines.append("%s\t%d\t%d\t%s\t%s\t.\t0\t0\t%s" % (c, r[2], r[3], cls, sz, color))
    return "\n".join(lines) + "\n"
