"""Parity of the CUDA path (through the C ABI, include/garlic_b200.h) against the CPU oracle on the
golden cases, and against the reference binary's committed outputs.  Needs a GPU (B200).

Bars (BASELINE.json north_star): packed genotypes, allele counts and ROH intervals bit-exact;
window LOD / wLOD values within 1e-9 relative (tolerance written in each test)."""
import numpy as np
import pytest

from garlic_b200 import synth
from garlic_b200.pipeline import HotPath
from oracle import oracle as orc
from tests.common import (arg, arg_list, flatten, golden_text, load_case, oracle_roh_idx,
                          oracle_windows_matrix)

pytestmark = pytest.mark.gpu

RTOL_WINDOWS = 1e-9     # north_star: "Window LOD/wLOD values must agree within 1e-9 relative"


def run_oracle(name, **kw):
    ds, args = load_case(name)
    W = arg(args, "--winsize", cast=int)
    err = arg(args, "--error", None, float)
    cutoff = arg(args, "--lod-cutoff", None, float)
    ov = arg(args, "--overlap-frac", 0.25, float)
    weighted = "--weighted" in args
    cm = "--cm" in args
    res = orc.run_pipeline(ds, W, err, cutoff, ov, weighted=weighted, cm=cm,
                           auto_overlap="--auto-overlap-frac" in args, **kw)
    return ds, args, res, dict(W=W, err=err, cutoff=cutoff, weighted=weighted, cm=cm)


def close_windows(got, want, rtol=RTOL_WINDOWS):
    miss_w = (want == orc.MISSING)
    assert np.array_equal(got == orc.MISSING, miss_w)
    nan_w = np.isnan(want)
    assert np.array_equal(np.isnan(got), nan_w)
    ok = ~miss_w & ~nan_w
    scale = np.maximum(np.abs(want[ok]), 1e-3)
    assert np.max(np.abs(got[ok] - want[ok]) / scale, initial=0.0) <= rtol


# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", ["lod_0", "lod_small", "wlod_cm", "auto_cutoff"])
def test_k1_coding_counts_and_packed_matrix(name):
    ds, args = load_case(name)
    geno, na, tot, one, freq = orc.code_tped(ds.alleles)
    hp = HotPath()
    g = hp.g
    g.set_shape(ds.n_ind, ds.n_loci, ds.chr_offsets, ds.pos)
    for s0 in range(0, ds.n_loci, 1024):           # several chunks
        g.put_alleles(ds.alleles[s0:s0 + 1024], s0)
    g.code_alleles()
    c_na, c_tot, c_hom, c_nm = g.get_counts()
    assert np.array_equal(c_na, na) and np.array_equal(c_tot, tot)
    assert np.array_equal(g.get_one_allele(), one)
    codes = geno.astype(np.uint8)
    assert np.array_equal(c_nm, (codes != 3).sum(1))
    assert np.array_equal(c_hom, ((codes == 0) | (codes == 2)).sum(1))
    packed = g.get_genotypes(False)
    assert np.array_equal(synth.unpack_codes(packed, ds.n_loci), codes.T)
    f, keep, L = g.filter()
    assert np.array_equal(f, freq)                  # bit-exact double(nalleles)/double(total)
    assert np.array_equal(keep, (freq > 0) & (freq < 1))
    # the committed reference .freq (6 significant digits) agrees too
    lines = golden_text(name, "out.freq").splitlines()[1:]
    assert all(l.split("\t")[4] == "%g" % f[i] for i, l in enumerate(lines))
    packed2 = g.get_genotypes(True)
    assert np.array_equal(synth.unpack_codes(packed2, L), codes[keep].T)
    hp.close()


def test_k2_count_packed_matches_coding_path():
    ds, args = load_case("lod_0")
    geno, na, tot, one, freq = orc.code_tped(ds.alleles)
    codes = geno.astype(np.uint8).T.copy()
    hp = HotPath()
    g = hp.g
    g.set_shape(ds.n_ind, ds.n_loci, ds.chr_offsets, ds.pos)
    g.put_packed(synth.pack_codes(codes))
    # half-missing calls are not representable in 2 bits: the loader supplies them as corrections
    base_na = np.where(codes == 3, 0, codes).sum(0).astype(np.int32)
    base_tot = (2 * (codes != 3).sum(0)).astype(np.int32)
    g.count_packed(na - base_na, tot - base_tot)
    c_na, c_tot, c_hom, c_nm = g.get_counts()
    assert np.array_equal(c_na, na) and np.array_equal(c_tot, tot)
    assert np.array_equal(c_nm, (codes != 3).sum(0))
    assert np.array_equal(c_hom, ((codes == 0) | (codes == 2)).sum(0))
    g.count_packed()
    c_na, c_tot, _, _ = g.get_counts()
    assert np.array_equal(c_na, base_na) and np.array_equal(c_tot, base_tot)
    hp.close()


def test_k2_many_rows_counter_flush():
    """> 255·8 rows per block exercises the 8-bit partial-counter flush."""
    rng = np.random.default_rng(5)
    N, L = 5000, 700
    codes = rng.integers(0, 4, (N, L)).astype(np.uint8)
    hp = HotPath()
    g = hp.g
    g.set_shape(N, L, np.array([0, 300, L]), np.arange(L) * 1000 + 1000)
    g.put_packed(synth.pack_codes(codes))
    g.count_packed()
    c_na, c_tot, c_hom, c_nm = g.get_counts()
    assert np.array_equal(c_na, np.where(codes == 3, 0, codes).sum(0))
    assert np.array_equal(c_nm, (codes != 3).sum(0))
    assert np.array_equal(c_hom, ((codes == 0) | (codes == 2)).sum(0))
    hp.close()


@pytest.mark.parametrize("name", ["lod_0", "lod_3", "lod_small", "auto_overlap_hg19"])
def test_lut_and_windows(name):
    ds, args, res, p = run_oracle(name)
    hp = HotPath().load(ds, error=p["err"])
    assert hp.L == res["n_used"]
    F = flatten(res, p["err"])
    lut = hp.g.get_lut()
    # device log10 vs host libm: a few ulp at most
    assert np.allclose(lut, F["lut"][:hp.L], rtol=1e-14, atol=1e-16)
    want = oracle_windows_matrix(res)
    # (1) product path (device-built table), exact chains and chunked: 1e-9 relative
    close_windows(hp.g.windows(p["W"], 1, exact=True), want)
    close_windows(hp.g.windows(p["W"], 1, exact=False), want)
    # (2) same table as the oracle → whole-segment chains are bit-identical to the reference recurrence
    hp.g.set_lut(F["lut"][:hp.L])
    assert np.array_equal(hp.g.windows(p["W"], 1, exact=True), want)
    hp.close()


@pytest.mark.parametrize("name", ["lod_0", "lod_1", "lod_2", "lod_3", "lod_small", "auto_overlap_hg19", "lod_cm"])
@pytest.mark.parametrize("exact", [False, True])
def test_roh_unweighted_bit_exact(name, exact):
    ds, args, res, p = run_oracle(name)
    hp = HotPath().load(ds, error=p["err"], cm=p["cm"])
    got = hp.roh(p["W"], p["cutoff"], res["overlap_frac"], exact=exact, cm=p["cm"])
    assert [(r[0], r[1], r[5], r[6]) for r in got] == oracle_roh_idx(res)
    # and the BED text equals the reference binary's committed output byte for byte
    bounds = arg_list(args, "--size-bounds")
    bed = orc.format_bed(got, ds.ind_ids, hp.labels, bounds, ds.pop, cm=p["cm"])
    assert bed == golden_text(name, "out.roh.bed")
    hp.close()


@pytest.mark.parametrize("name", ["gl_pl", "gl_gl", "gl_gq"])
def test_gl_path(name):
    ds, args, res, p = run_oracle(name)
    hp = HotPath().load(ds, error=None)
    close_windows(hp.g.windows(p["W"], 1, exact=True), oracle_windows_matrix(res))
    for exact in (False, True):
        got = hp.roh(p["W"], p["cutoff"], res["overlap_frac"], exact=exact)
        assert [(r[0], r[1], r[5], r[6]) for r in got] == oracle_roh_idx(res)
        bed = orc.format_bed(got, ds.ind_ids, hp.labels, arg_list(args, "--size-bounds"), ds.pop)
        assert bed == golden_text(name, "out.roh.bed")
    hp.close()


def _tgls_text(tokens_by_line, rng):
    """Raw value columns of a tgls file (what follows the 4th field): tokens separated by random blanks / tabs."""
    text, off = bytearray(), [0]
    for toks in tokens_by_line:
        line = ""
        for t in toks:
            line += rng.choice([" ", "\t", "  ", " \t "]) + t
        line += rng.choice(["", " ", "\r"])
        text += line.encode()
        off.append(len(text))
    return bytes(text), np.array(off, np.int64)


def test_k0_tgls_text_on_device_equals_host_strtod():
    """K0-GL (garlic_gpu_put_tgls_text): the likelihood columns converted on the GPU equal strtod of every token bit for
    bit — fast-path tokens (phred integers, short decimals, exponents) on the device, the others (19+ digits, huge
    exponents, 16-digit mantissas, hex) through the host list — for a whole shard and for a shard that starts at
    individual 7; token counts per line come back for the column check (garlic-data.cpp:1531-1554)."""
    rng = np.random.default_rng(11)
    n_ind, L0 = 37, 300
    fmts = [lambda r: str(int(r.integers(0, 256))), lambda r: "%.4f" % r.uniform(0, 3), lambda r: "%g" % r.uniform(1e-6, 1e3),
            lambda r: "-%.3f" % r.uniform(0, 12), lambda r: "%.6e" % r.uniform(1e-12, 1e12), lambda r: "+%d" % r.integers(0, 99),
            lambda r: "0.%018d" % r.integers(0, 10 ** 18), lambda r: "%.17g" % r.uniform(0, 1), lambda r: "1e%d" % r.integers(-330, 310),
            lambda r: "000%d.500" % r.integers(0, 999), lambda r: ".5", lambda r: "7.", lambda r: "0x1.8p3", lambda r: "0", lambda r: "-0.0",
            lambda r: "123456789012345678901234", lambda r: "9007199254740993", lambda r: "1E5"]
    toks = [[fmts[int(rng.integers(0, len(fmts)))](rng) for _ in range(n_ind)] for _ in range(L0)]
    want = np.array([[float.fromhex(t) if t.startswith("0x") else float(t) for t in line] for line in toks]).T   # [ind][snp]
    text, off = _tgls_text(toks, rng)
    from garlic_b200.api import GarlicGPU
    pos = np.arange(1, L0 + 1, dtype=np.int32) * 100
    for lo, n in ((0, n_ind), (7, 20)):
        g = GarlicGPU(0)
        g.set_shape(n, L0, np.array([0, L0], np.int64), pos, ind_offset=lo)
        for s0 in range(0, L0, 128):                       # blocks of lines, as the driver streams them
            s1 = min(L0, s0 + 128)
            ntok = g.put_tgls_text(text, off[s0:s1 + 1], "PL", snp0=s0)
            assert np.all(ntok == n_ind)
        got = g.get_gl()
        assert got.tobytes() == np.ascontiguousarray(want[lo:lo + n]).tobytes()
        g.close()
    # a short and a long line are reported through the token counts
    toks2 = [toks[0][:-2], toks[1] + ["5"]]
    text2, off2 = _tgls_text(toks2, rng)
    g = GarlicGPU(0)
    g.set_shape(n_ind, L0, np.array([0, L0], np.int64), pos)
    assert list(g.put_tgls_text(text2, off2, "PL")) == [n_ind - 2, n_ind + 1]
    g.close()


@pytest.mark.parametrize("name", ["gl_pl", "gl_gq"])
def test_gl_path_from_tgls_text(name):
    """The GL golden cases with the likelihood matrix ingested as text: same windows and ROH as the oracle."""
    ds, args, res, p = run_oracle(name)
    hp = HotPath()
    gl = ds.gl
    ds.gl = None
    try:
        g = hp.g
        # HotPath.load without the matrix, the text goes in between the genotype upload and the filter
        rng = np.random.default_rng(3)
        toks = [[("%d" % v if float(v).is_integer() else repr(float(v))) for v in row] for row in np.asarray(gl)]   # [snp][ind]
        text, off = _tgls_text(toks, rng)
        orig_filter = g.filter

        def filter_with_text(*a, **k):
            ntok = g.put_tgls_text(text, off, ds.gl_type)
            assert np.all(ntok == ds.n_ind)
            return orig_filter(*a, **k)
        g.filter = filter_with_text
        hp.load(ds, error=None)
    finally:
        ds.gl = gl
    close_windows(hp.g.windows(p["W"], 1, exact=True), oracle_windows_matrix(res))
    got = hp.roh(p["W"], p["cutoff"], res["overlap_frac"])
    assert [(r[0], r[1], r[5], r[6]) for r in got] == oracle_roh_idx(res)
    hp.close()


def test_weighted_ld_wlod_and_roh():
    ds, args, res, p = run_oracle("wlod_cm")
    hp = HotPath().load(ds, weighted=True, cm=True, error=p["err"])
    assert hp.L == res["n_used"]
    assert np.array_equal(hp.gpos, np.concatenate([c["gpos"] for c in res["chroms"]]))
    homf = hp.g.get_hom_freq()
    assert np.array_equal(homf, np.concatenate([c["homf"] for c in res["chroms"]]), equal_nan=True)
    ld = hp.g.ld_band(p["W"], None, want_ld=True)
    want_ld = np.concatenate([c["LD"] for c in res["chroms"]], axis=0)
    # integer popcounts + IEEE division/multiplication in the reference's order: bit-exact
    assert np.array_equal(ld, want_ld, equal_nan=True)
    close_windows(hp.g.windows(p["W"], 1, weighted=True), oracle_windows_matrix(res))
    got = hp.roh(p["W"], p["cutoff"], res["overlap_frac"], weighted=True, cm=True)
    assert [(r[0], r[1], r[5], r[6]) for r in got] == oracle_roh_idx(res)
    bed = orc.format_bed(got, ds.ind_ids, hp.labels, arg_list(args, "--size-bounds"), ds.pop, cm=True)
    assert bed == golden_text("wlod_cm", "out.roh.bed")
    hp.close()


@pytest.mark.parametrize("W", [25, 32, 72])
def test_weighted_ld_subsample(W):
    ds, args = load_case("wlod_cm")
    sub = np.array([0, 3, 4, 9, 10, 17, 20], np.int32)
    res = orc.run_pipeline(ds, W, 0.001, 0.5, 0.25, weighted=True, cm=True, ld_individuals=sub)
    hp = HotPath().load(ds, weighted=True, cm=True, error=0.001)
    ld = hp.g.ld_band(W, sub, want_ld=True)
    assert np.array_equal(ld, np.concatenate([c["LD"] for c in res["chroms"]], axis=0), equal_nan=True)
    close_windows(hp.g.windows(W, 1, weighted=True), oracle_windows_matrix(res))
    for exact in (False, True):      # tensor-core pass with exact re-evaluation / exact sums everywhere
        got = hp.roh(W, 0.5, 0.25, weighted=True, cm=True, exact=exact)
        assert len(got) > 0 and [(r[0], r[1], r[5], r[6]) for r in got] == oracle_roh_idx(res)
    hp.close()


def test_thinned_windows_for_kde_subsample():
    ds, args = load_case("auto_cutoff")
    W = 30
    sub = np.array([1, 5, 6, 20, 39], np.int32)
    res = orc.run_pipeline(ds, W, 0.001, None, thin_step=W, kde_individuals=sub)
    hp = HotPath().load(ds, error=0.001)
    got = hp.thinned(W, W, sub)
    assert got.shape == res["thinned"].shape
    assert np.max(np.abs(got - res["thinned"]) / np.maximum(np.abs(res["thinned"]), 1e-3)) <= RTOL_WINDOWS
    # all individuals: count equals the reference log's "KDE with N points"
    n = int(golden_text("auto_cutoff", "out.log").split("KDE with ")[1].split(" ")[0])
    assert len(hp.thinned(W, W)) == n
    hp.close()


def test_kde_on_device_matches_restatement():
    """garlic_gpu_kde (computeKDE on the device): nrd0 bandwidth by radix selection of the quantiles, 512
    targets, exact Gauss transform — of the window matrix pass 1 left on the GPU (MISSING slots skipped in place) and of
    host values — against the restatement of garlic-kde.cpp:14-140 on the oracle's thinned windows; same cutoff from
    get_min_btw_modes; two runs bit-identical (fixed reduction order)."""
    ds, args = load_case("auto_cutoff")
    W = 30
    res = orc.run_pipeline(ds, W, 0.001, None, thin_step=W)
    t, y, h = orc.compute_kde(res["thinned"])
    hp = HotPath().load(ds, error=0.001)
    data = hp.thinned(W, W)                                   # leaves the thinned matrix on the device
    x1, y1, n1, h1 = hp.g.kde()
    assert n1 == len(res["thinned"]) == len(data)
    assert abs(h1 - h) <= 1e-12 * h
    assert np.allclose(x1, t, rtol=1e-12, atol=1e-12)
    assert np.max(np.abs(y1 - y)) <= 1e-9 * np.max(y)
    assert orc.min_between_modes(x1, y1, W) == pytest.approx(orc.min_between_modes(t, y, W), rel=1e-12)
    x2, y2, n2, h2 = hp.g.kde()
    assert np.array_equal(y1, y2) and h1 == h2
    # host values (several GPUs: the gathered thinned windows), odd target count, duplicates and negative values
    rng = np.random.default_rng(5)
    v = np.concatenate([rng.normal(-8, 2, 3001), rng.normal(3, 1, 1500), np.full(40, 1.25)])
    for m in (512, 333, 1024):
        t3, y3, h3 = orc.compute_kde(v, m)
        x4, y4, n4, h4 = hp.g.kde(v, m)
        assert n4 == len(v) and abs(h4 - h3) <= 1e-12 * h3
        assert np.allclose(x4, t3, rtol=1e-12, atol=1e-12) and np.max(np.abs(y4 - y3)) <= 1e-9 * np.max(y3)
    hp.close()


def test_cutoff_on_a_window_value_is_resolved_exactly():
    """A cutoff equal to an actual window value: the chunked pass must detect the ambiguity and the
    exact re-evaluation must reproduce the reference's decision."""
    ds, args = load_case("lod_small")
    W = 50
    res0 = orc.run_pipeline(ds, W, 0.001, None)
    F = flatten(res0, 0.001)
    win = res0["chroms"][0]["win"]
    vals = np.sort(win[win != orc.MISSING])
    cutoff = float(vals[len(vals) // 2])
    res = orc.run_pipeline(ds, W, 0.001, cutoff)
    hp = HotPath().load(ds, error=0.001)
    hp.g.set_lut(F["lut"][:hp.L])
    got = hp.roh(W, cutoff, 0.25, exact=False)
    assert hp.g.last_stats()["ambiguous_pairs"] >= 1
    assert [(r[0], r[1], r[5], r[6]) for r in got] == oracle_roh_idx(res)
    hp.close()


def test_tiny_panel_frequencies_widen_the_ambiguity_tolerance():
    """--freq-file frequencies around 1e-25 give |lod| > 20: the bound on the table's entries behind the ambiguity
    tolerance follows the smallest frequency supplied (capi.cu:lod_bound), so the chunked / pruned pass with a cutoff ON a
    window value still marks the pair ambiguous and equals whole-segment exact chains; windows match a direct fp64 sum
    over the device's own table."""
    from garlic_b200.api import GarlicGPU
    names, offs, pos, cens = synth.make_positions_genomewide(9, 30000)
    codes = synth.make_codes(9, 96, 30000)
    cen = np.array([cens["chr" + n] for n in names], np.int32)
    g = GarlicGPU(0)
    g.set_shape(96, 30000, offs, pos)
    g.put_packed(synth.pack_codes(codes))
    g.count_packed()
    freq, keep, L = g.filter()
    f = freq.copy()
    poly = np.flatnonzero((f > 0) & (f < 1))
    f[poly[::53]] = 1e-25
    f[poly[7::211]] = 1.0 - 1e-13
    g.filter(freq_override=f)
    g.set_tables(0.001, 200000, cen)
    lut = g.get_lut()
    assert np.abs(lut).max() > 20.0
    W = 50
    win = g.windows(W, 1, exact=True)
    # a window that holds a tiny-frequency SNP, of an individual homozygous there: its value is the cutoff
    big = np.argwhere((win != orc.MISSING) & (win > 20.0))
    assert len(big) > 0
    i, t = big[len(big) // 2]
    cutoff = float(win[i, t])
    fast = g.call_roh(W, cutoff, 0.25)
    st = g.last_stats()
    assert st["ambiguous_pairs"] >= 1
    exact = g.call_roh(W, cutoff, 0.25, exact=True)
    assert len(exact) > 0 and np.array_equal(fast, exact)
    # the device's windows against a direct sum over its own table
    kept = g.get_kept_index()
    ck = codes[:, kept]
    val = np.take_along_axis(np.broadcast_to(lut[None], (96, L, 4)), ck[:, :, None].astype(np.int64), 2)[:, :, 0]
    direct = np.lib.stride_tricks.sliding_window_view(val, W, axis=1).sum(-1)
    ok = win[:, :direct.shape[1]] != orc.MISSING
    d = np.abs(win[:, :direct.shape[1]][ok] - direct[ok]) / np.maximum(np.abs(direct[ok]), 1e-3)
    assert d.max() <= 1e-9
    g.close()


def test_empty_and_ragged_inputs():
    # a chromosome shorter than the window, a single individual, nothing above the cutoff
    # (one individual: only heterozygous / half-missing sites survive the 0<freq<1 filter)
    ds = synth.make_dataset(seed=3, n_ind=1, chr_sizes=(40, 2500), centromere=None)
    res = orc.run_pipeline(ds, 60, 0.001, 1e9, 0.25)
    hp = HotPath().load(ds, error=0.001)
    assert hp.L == res["n_used"] and len(res["chroms"][0]["pos"]) < 60 < len(res["chroms"][1]["pos"])
    assert hp.roh(60, 1e9, 0.25) == []
    # every valid window passes (cutoff above the MISSING sentinel: at or below -9999 the reference
    # itself overruns inWin[], garlic-roh.cpp:450-453, and the C ABI rejects it)
    res = orc.run_pipeline(ds, 60, 0.001, -5000.0, 0.25)
    got = hp.roh(60, -5000.0, 0.25)
    assert len(got) >= 1 and [(r[0], r[1], r[5], r[6]) for r in got] == oracle_roh_idx(res)
    with pytest.raises(Exception):
        hp.roh(60, -1e9, 0.25)
    hp.close()
    # 33 individuals (one lane of a second warp), window larger than every chromosome → no windows
    ds = synth.make_dataset(seed=4, n_ind=33, chr_sizes=(50, 45), centromere=None)
    hp = HotPath().load(ds, error=0.001)
    assert hp.roh(64, 0.0, 0.25) == []
    assert np.all(hp.g.windows(64, 1) == orc.MISSING)
    hp.close()


def test_full_size_properties_c2_shape():
    """Size-independent properties at a larger shape (2,000 × 60k): exact == chunked ROH,
    idempotence, per-individual independence (a subset of rows gives the same ROH)."""
    names, offs, pos, cens = synth.make_positions_genomewide(2, 60000)
    codes = synth.make_codes(2, 512, 60000)

    class DS:
        pass
    ds = DS()
    ds.chr_names, ds.chr_offsets, ds.pos, ds.centromeres = names, offs, pos, cens
    ds.gl = None
    hp = HotPath().load(ds, error=0.001, packed_rows=synth.pack_codes(codes))
    a = hp.g.call_roh(50, 2.0, 0.25, exact=False)
    b = hp.g.call_roh(50, 2.0, 0.25, exact=True)
    c = hp.g.call_roh(50, 2.0, 0.25, exact=False)
    assert np.array_equal(a, b) and np.array_equal(a, c) and len(a) > 100
    # sortedness (ind, chr, start) and disjointness within an individual
    key = a[:, 0].astype(np.int64) * (1 << 32) + a[:, 2]
    assert np.all(np.diff(key) > 0)
    hp.close()


def test_repeated_runs_are_identical_and_equal_exact_chains():
    """The same count → filter → tables → call_roh sequence, repeated on one handle, returns identical ROH every
    time (no race between the stream-ordered phases, pruning pass included) and equals whole-segment chains."""
    names, offs, pos, cens = synth.make_positions_genomewide(9, 80000)
    codes = synth.make_codes(9, 300, 80000)
    rows = synth.pack_codes(codes)
    hp = HotPath()
    g = hp.g
    g.set_shape(300, 80000, offs, pos)
    g.put_packed(rows)
    cen = np.array([cens["chr" + n] for n in names], np.int32)
    outs = []
    for _ in range(6):
        g.count_packed()
        g.filter()
        g.set_tables(0.001, 200000, cen)
        outs.append(g.call_roh(50, 2.0, 0.25).copy())
    st = g.last_stats()
    assert 0 <= st["candidate_pairs"] < st["all_pairs"]          # the pruning pass ran and pruned
    assert all(np.array_equal(o, outs[0]) for o in outs) and len(outs[0]) > 100
    assert np.array_equal(outs[0], g.call_roh(50, 2.0, 0.25, exact=True))
    hp.close()


def _emu_bounds(codes_kept, lut, L, W):
    """bound.cuh on the CPU (tests/host_emu.cpp): the piece maxima the fused kernel must reproduce bit for bit."""
    import ctypes as C
    from tests.test_host_emu import emu, _p
    N = codes_kept.shape[0]
    row_words = (((L + 4160 + 31) >> 5) + 3) & ~1
    rows = synth.pack_codes(codes_kept, row_bytes=row_words * 8).view(np.uint64).reshape(N, row_words)
    lut_pad = np.zeros((L + 4160, 4))
    lut_pad[:L] = lut
    n_pieces = (L + 255) // 256
    pmax = np.zeros((n_pieces, N), np.uint32)
    rc = emu().emu_bound(_p(rows), C.c_int64(row_words), _p(lut_pad), C.c_longlong(L), C.c_int(N), C.c_int(W), _p(pmax),
                         C.c_int(n_pieces))
    assert rc == 0
    return pmax


@pytest.mark.parametrize("n_ind,L0,W,W2", [(300, 80000, 50, 100), (77, 30011, 32, 209), (513, 20000, 64, 33), (130, 25000, 30, 16),
                                           (90, 22000, 17, 300), (64, 40000, 250, 24)])
def test_fused_compaction_and_bound_equal_host_emulation(n_ind, L0, W, W2):
    """squeeze.cu: the compacted matrix equals the column gather, and the piece maxima written by the fused pass
    (window size W, first consumer = thinned windows) and by the bound-only pass (W2, same data) equal the CPU
    emulation of bound.cuh bit for bit; ragged individual counts and SNP counts included."""
    names, offs, pos, cens = synth.make_positions_genomewide(11, L0)
    codes = synth.make_codes(11, n_ind, L0)
    hp = HotPath()
    g = hp.g
    g.set_shape(n_ind, L0, offs, pos)
    g.put_packed(synth.pack_codes(codes))
    g.count_packed()
    freq, keep, L = g.filter()
    keep = keep.copy()
    g.set_tables(0.001, 200000, np.array([cens["chr" + n] for n in names], np.int32))
    g.windows(W, W, individuals=np.arange(0, n_ind, 37, dtype=np.int32), exact=False)    # compaction + bound(W), fused
    st = g.last_stats()
    assert st["squeeze_ms"] > 0
    packed = g.get_genotypes(True)
    assert np.array_equal(synth.unpack_codes(packed, L), codes[:, keep])
    lut = g.get_lut()
    want = _emu_bounds(codes[:, keep], lut, L, W)
    assert np.array_equal(g.piece_bounds(W), want)
    want2 = _emu_bounds(codes[:, keep], lut, L, W2)
    assert np.array_equal(g.piece_bounds(W2), want2)                                     # bound only, compacted rows
    hp.close()


@pytest.mark.parametrize("W,cutoff", [(50, 2.0), (32, 0.5), (100, 5.0), (209, 20.0), (60, -3.0), (10, 1.0), (16, 1.0), (30, 1.5), (31, 2.0),
                                      (250, 25.0), (400, 40.0), (600, 50.0)])
def test_pruned_pass_equals_exact_chains_and_unpruned_pass(W, cutoff):
    """Pass 2 over the candidates the bound leaves == whole-segment exact chains == the pass without pruning, for
    several window-size classes and cutoffs (a negative cutoff makes nearly every pair a candidate)."""
    import os
    names, offs, pos, cens = synth.make_positions_genomewide(4, 70000)
    codes = synth.make_codes(4, 330, 70000)
    rows = synth.pack_codes(codes)
    cen = np.array([cens["chr" + n] for n in names], np.int32)
    outs = {}
    for mode in ("pruned", "unpruned"):
        if mode == "unpruned":
            os.environ["GARLIC_NO_PRUNE"] = "1"
        try:
            hp = HotPath()
        finally:
            os.environ.pop("GARLIC_NO_PRUNE", None)
        g = hp.g
        g.set_shape(330, 70000, offs, pos)
        g.put_packed(rows)
        g.count_packed()
        g.filter()
        g.set_tables(0.001, 200000, cen)
        outs[mode] = g.call_roh(W, cutoff, 0.25).copy()
        st = g.last_stats()
        if mode == "pruned" and W < 16:
            assert st["candidate_pairs"] < 0               # below the bound's range: every pair is walked
            outs["exact"] = g.call_roh(W, cutoff, 0.25, exact=True).copy()
        elif mode == "pruned":
            assert 0 <= st["candidate_pairs"] <= st["all_pairs"]
            if cutoff >= 2.0 and 32 <= W <= 209:          # (beyond 209 the optimistic remainder weakens the bound)
                assert st["candidate_pairs"] < 0.5 * st["all_pairs"]
            outs["exact"] = g.call_roh(W, cutoff, 0.25, exact=True).copy()
        else:
            assert st["candidate_pairs"] < 0
        hp.close()
    assert len(outs["exact"]) > (50 if W <= 400 else -1)
    assert np.array_equal(outs["pruned"], outs["exact"])
    assert np.array_equal(outs["unpruned"], outs["exact"])


def _weighted_handle(n_ind, L0, seed, gl=False):
    names, offs, pos, cens = synth.make_positions_genomewide(seed, L0, n_chr=3)
    codes = synth.make_codes(seed, n_ind, L0)

    class DS:
        pass
    ds = DS()
    ds.chr_names, ds.chr_offsets, ds.pos, ds.centromeres = names, offs, pos, cens
    C = len(names)
    ds.map_pos = [pos[offs[c]:offs[c + 1]][2:-2:5].astype(np.int64) for c in range(C)]
    ds.map_cm = [np.round(p * 1.2e-6, 9) for p in ds.map_pos]
    ds.gl = None
    if gl:
        rng = np.random.default_rng(seed + 100)
        ds.gl = rng.choice(np.array([0.0004, 0.004, 0.04, 0.5, 3.0, 0.0, 150.0]), size=(L0, n_ind))
        ds.gl_type = "PL"
    return HotPath().load(ds, weighted=True, cm=True, error=None if gl else 0.001, packed_rows=synth.pack_codes(codes))


@pytest.mark.parametrize("W,gl", [(25, False), (72, False), (33, False), (40, True)])
def test_weighted_tensor_core_pass_equals_exact_sums(W, gl):
    """Pass 2 of the weighted path: the tolerance-checked DMMA pass (wlod_mma_kernel) returns the same ROH as the
    exact mul-then-add kernel, also when the cutoff sits exactly on a window value (ambiguous pairs are re-walked
    exactly), for W below / above one flag word, a ragged last individual group, and per-genotype likelihoods."""
    hp = _weighted_handle(77, 30000, 21 + W, gl)
    g = hp.g
    ld_ind = np.arange(0, 77, 2, dtype=np.int32)
    g.ld_band(W, ld_ind)
    win = g.windows(W, 1, weighted=True, individuals=np.array([5], np.int32))[0]
    vals = np.sort(win[(win != orc.MISSING) & ~np.isnan(win)])
    for cutoff in (float(vals[int(len(vals) * 0.97)]), 0.5):
        a = g.call_roh(W, cutoff, 0.25, weighted=True, exact=False).copy()
        st = g.last_stats()
        b = g.call_roh(W, cutoff, 0.25, weighted=True, exact=True)
        assert np.array_equal(a, b) and len(a) > 0
        if cutoff != 0.5:
            assert st["ambiguous_pairs"] >= 1       # individual 5 holds a window equal to the cutoff
    hp.close()


@pytest.mark.parametrize("W", [200, 50, 16, 31, 330])
def test_gl_ring_walker_equals_generic_walker(W):
    """GL mode, pass 2: gl_walk_kernel (bulk-copied shared-memory ring over the lane-interleaved likelihood matrix)
    returns the same ROH as the generic per-lane walker and as whole-segment chains; 77 individuals leave a ragged
    last group; cutoff on an actual window value forces exact re-evaluation of an ambiguous pair."""
    import os
    names, offs, pos, cens = synth.make_positions_genomewide(31, 40000, n_chr=3)
    codes = synth.make_codes(31, 77, 40000)

    class DS:
        pass
    ds = DS()
    ds.chr_names, ds.chr_offsets, ds.pos, ds.centromeres = names, offs, pos, cens
    rng = np.random.default_rng(5)
    ds.gl = rng.choice(np.array([0.0004, 0.004, 0.04, 0.5, 3.0, 0.0, 150.0]), size=(40000, 77))
    ds.gl_type = "PL"
    hp = HotPath().load(ds, error=None, packed_rows=synth.pack_codes(codes))
    g = hp.g
    win = g.windows(W, 1, individuals=np.array([70], np.int32), exact=True)[0]
    vals = np.sort(win[win != orc.MISSING])
    thin = g.windows(W, W, individuals=np.array([3, 70], np.int32), exact=False)       # thinned pass 1 (direct sums)
    full = g.windows(W, 1, individuals=np.array([3, 70], np.int32), exact=True)
    base = 0
    for c in range(3):
        n = int(hp.chr_off[c + 1] - hp.chr_off[c])
        a, b = thin[:, base:base + (n + W - 1) // W], full[:, hp.chr_off[c]:hp.chr_off[c + 1]][:, ::W]
        ok = b != orc.MISSING
        assert np.array_equal(a == orc.MISSING, ~ok)
        assert np.max(np.abs(a[ok] - b[ok]) / np.maximum(np.abs(b[ok]), 1e-3), initial=0.0) <= RTOL_WINDOWS
        base += (n + W - 1) // W
    for cutoff in (float(vals[int(len(vals) * 0.98)]), 1.0):
        a = g.call_roh(W, cutoff, 0.25, exact=False).copy()
        st = g.last_stats()
        b = g.call_roh(W, cutoff, 0.25, exact=True).copy()
        os.environ["GARLIC_NO_GL_RING"] = "1"
        try:
            c = g.call_roh(W, cutoff, 0.25, exact=True).copy()
            d = g.call_roh(W, cutoff, 0.25, exact=False).copy()
        finally:
            del os.environ["GARLIC_NO_GL_RING"]
        assert np.array_equal(a, b) and np.array_equal(a, c) and np.array_equal(a, d) and len(a) > 0
        if cutoff != 1.0:
            assert st["ambiguous_pairs"] >= 1
    hp.close()


@pytest.mark.parametrize("W,n_ind,L0", [(200, 77, 40000), (330, 200, 60000), (64, 40, 30000)])
def test_gl_relay_warps_equal_single_warp(W, n_ind, L0):
    """GL mode, pass 2: the relay kernel (K warps sharing one ring, base + in-block prefix) returns the same ROH for
    every K as the single-warp ring walker, repeatedly (cache-resident data makes the bulk copies land early: the
    hand-over and slot-reuse ordering is what this exercises)."""
    import os
    names, offs, pos, cens = synth.make_positions_genomewide(31, L0, n_chr=3)
    codes = synth.make_codes(31, n_ind, L0)

    class DS:
        pass
    ds = DS()
    ds.chr_names, ds.chr_offsets, ds.pos, ds.centromeres = names, offs, pos, cens
    rng = np.random.default_rng(5)
    ds.gl = rng.choice(np.array([0.0004, 0.004, 0.04, 0.5, 3.0, 0.0, 150.0]), size=(L0, n_ind))
    ds.gl_type = "PL"
    hp = HotPath().load(ds, error=None, packed_rows=synth.pack_codes(codes))
    g = hp.g
    try:
        os.environ["GARLIC_GL_WARPS"] = "1"
        ref = g.call_roh(W, 1.0, 0.25).copy()
        assert len(ref) > 50
        for K in (2, 3, 4):
            os.environ["GARLIC_GL_WARPS"] = str(K)
            for _ in range(3):
                assert np.array_equal(g.call_roh(W, 1.0, 0.25), ref)
    finally:
        del os.environ["GARLIC_GL_WARPS"]
    assert np.array_equal(g.call_roh(W, 1.0, 0.25), ref)          # default K
    assert np.array_equal(g.call_roh(W, 1.0, 0.25, exact=True), ref)
    hp.close()


def test_k0_tped_text_tokeniser_equals_allele_upload():
    """K0: raw tped genotype columns (single blanks, tabs, runs of blanks, CR, leading blanks) tokenised on the GPU give
    the same packed matrix, counts and "1" alleles as uploading the allele characters; two shards of individuals
    each keep their own columns; a short line is reported through the non-blank count."""
    from garlic_b200.api import GarlicGPU
    ds, args = load_case("lod_small")
    L0, N = ds.n_loci, ds.n_ind
    rng = np.random.default_rng(11)
    seps = [b" ", b"\t", b"  ", b" \t "]
    lines, off = [], [0]
    for l in range(L0):
        a = ds.alleles[l].reshape(-1)
        style = l % 4
        if style == 0:
            tail = b" " + b" ".join(bytes([x]) for x in a)
        elif style == 1:
            tail = b"\t" + b"\t".join(bytes([x]) for x in a) + b"\r"
        else:
            tail = b"".join(seps[int(rng.integers(0, 4))] + bytes([x]) for x in a) + (b"  " if style == 3 else b"")
        lines.append(tail)
        off.append(off[-1] + len(tail))
    text = b"".join(lines)
    ref = GarlicGPU(0)
    ref.set_shape(N, L0, ds.chr_offsets, ds.pos)
    ref.put_alleles(ds.alleles, 0)
    ref.code_alleles()
    want_geno, want_counts, want_one = ref.get_genotypes(), [c.copy() for c in ref.get_counts()], ref.get_one_allele().copy()
    ref.close()
    g = GarlicGPU(0)
    g.set_shape(N, L0, ds.chr_offsets, ds.pos)
    blk = 992                                                   # several uploads, multiple of 32
    for s0 in range(0, L0, blk):
        n = min(blk, L0 - s0)
        nb = g.put_tped_text(text, off[s0:s0 + n + 1], s0)
        assert np.all(nb == 2 * N)
    g.code_alleles()
    assert np.array_equal(g.get_genotypes(), want_geno)
    assert all(np.array_equal(a, b) for a, b in zip(g.get_counts(), want_counts))
    assert np.array_equal(g.get_one_allele(), want_one)
    g.close()
    # a shard holding the middle half of the individuals: only its columns, same codes
    lo, hi = N // 4, N // 4 + N // 2
    g = GarlicGPU(0)
    g.set_shape(hi - lo, L0, ds.chr_offsets, ds.pos, ind_offset=lo)
    g.put_tped_text(text, off, 0)
    g.code_alleles()
    sub = g.get_genotypes()
    # the "1" allele of a shard is the first non-missing character among ITS individuals (the cross-GPU MIN all-reduce
    # restores the global one); compare codes where both agree on it
    one_s = g.get_one_allele()
    same = one_s == want_one
    def codes(rows):                                            # packed rows → uint8[n][L0] genotype codes
        b = rows[:, :(L0 + 3) // 4]
        return np.stack([(b >> (2 * k)) & 3 for k in range(4)], axis=2).reshape(b.shape[0], -1)[:, :L0]
    assert same.mean() > 0.5 and np.array_equal(codes(sub)[:, same], codes(want_geno)[lo:hi][:, same])
    # truncated line: one allele short
    short = text[:off[1] - 2]
    nb = g.put_tped_text(short + b" ", [0, len(short) + 1], 0)
    assert nb[0] == 2 * N - 1
    g.close()


@pytest.mark.parametrize("n_ld,W", [(500, 40), (130, 25), (64, 33), (350, 72)])
def test_ld_band_fused_kernel_many_individuals(n_ld, W):
    """K6 with several plane words per SNP (the C3 shape: 500 LD individuals = 8 words): the fused pair + window-sum kernel
    is bit-identical to the oracle's calcHR2LD restatement (garlic-data.cpp:474-527,558-583) and to the unfused
    kernels (ordered pair matrix in global memory, GARLIC_LD_UNFUSED=1)."""
    import os
    ds = synth.make_dataset(seed=21, n_ind=520, chr_sizes=(420, 380), with_map=True, n_roh=2, roh_snps=(40, 120))
    sub = np.sort(np.random.default_rng(2).choice(520, n_ld, replace=False)).astype(np.int32)
    res = orc.run_pipeline(ds, W, 0.001, 0.5, 0.25, weighted=True, cm=True, ld_individuals=sub)
    want = np.concatenate([c["LD"] for c in res["chroms"]], axis=0)
    hp = HotPath().load(ds, weighted=True, cm=True, error=0.001)
    got = hp.g.ld_band(W, sub, want_ld=True)
    assert np.array_equal(got, want, equal_nan=True)
    os.environ["GARLIC_LD_UNFUSED"] = "1"
    try:
        assert np.array_equal(hp.g.ld_band(W, sub, want_ld=True), want, equal_nan=True)
    finally:
        os.environ.pop("GARLIC_LD_UNFUSED")
    close_windows(hp.g.windows(W, 1, weighted=True), oracle_windows_matrix(res))
    got_roh = hp.roh(W, 0.5, 0.25, weighted=True, cm=True)
    assert [(r[0], r[1], r[5], r[6]) for r in got_roh] == oracle_roh_idx(res)
    hp.close()


def test_phased_r2_ld_band_wlod_and_roh():
    """--weighted --phased: the LD band from r2 between haplotypes (first-copy bits read from the allele block, four
    bit-planes, popcounts) is bit-identical to the oracle's restatement of calcR2LD; windows within 1e-9, ROH and BED
    equal the reference binary's."""
    ds, args = load_case("wlod_phased")
    W = arg(args, "--winsize", cast=int)
    cutoff = arg(args, "--lod-cutoff", None, float)
    res = orc.run_pipeline(ds, W, 0.001, cutoff, 0.25, weighted=True, cm=True, phased=True)
    hp = HotPath().load(ds, weighted=True, cm=True, error=0.001)
    hp.g.set_phased(True)
    ld = hp.g.ld_band(W, None, want_ld=True)
    assert np.array_equal(ld, np.concatenate([c["LD"] for c in res["chroms"]], axis=0), equal_nan=True)
    close_windows(hp.g.windows(W, 1, weighted=True), oracle_windows_matrix(res))
    for exact in (False, True):
        got = hp.roh(W, cutoff, res["overlap_frac"], weighted=True, cm=True, exact=exact)
        assert [(r[0], r[1], r[5], r[6]) for r in got] == oracle_roh_idx(res)
    bed = orc.format_bed(got, ds.ind_ids, hp.labels, arg_list(args, "--size-bounds"), ds.pop, cm=True)
    assert bed == golden_text("wlod_phased", "out.roh.bed")
    # an LD subsample, and hr2 again after switching back
    sub = np.array([0, 3, 4, 9, 10, 17, 20], np.int32)
    res2 = orc.run_pipeline(ds, W, 0.001, cutoff, 0.25, weighted=True, cm=True, phased=True, ld_individuals=sub)
    assert np.array_equal(hp.g.ld_band(W, sub, want_ld=True), np.concatenate([c["LD"] for c in res2["chroms"]], axis=0), equal_nan=True)
    hp.g.set_phased(False)
    res3 = orc.run_pipeline(ds, W, 0.001, cutoff, 0.25, weighted=True, cm=True)
    assert np.array_equal(hp.g.ld_band(W, None, want_ld=True), np.concatenate([c["LD"] for c in res3["chroms"]], axis=0), equal_nan=True)
    hp.close()
    # pre-packed genotypes carry no phase
    codes = synth.make_codes(3, 8, 4000)
    names, offs, pos, cens = synth.make_positions_genomewide(3, 4000, n_chr=2)

    class DS:
        pass
    d2 = DS()
    d2.chr_names, d2.chr_offsets, d2.pos, d2.centromeres, d2.gl = names, offs, pos, cens, None
    d2.map_pos = [pos[offs[c]:offs[c + 1]][::5].astype(np.int64) for c in range(2)]
    d2.map_cm = [p * 1.2e-6 for p in d2.map_pos]
    hp = HotPath().load(d2, weighted=True, cm=True, error=0.001, packed_rows=synth.pack_codes(codes))
    hp.g.set_phased(True)
    with pytest.raises(Exception):
        hp.g.ld_band(25, None)
    hp.close()


@pytest.mark.parametrize("weighted", [False, True])
def test_largest_window_sizes(weighted):
    """Window sizes up to the C ABI's limit (4096): the flag-history ring then needs more than the default 48 KB of
    dynamic shared memory; fast pass == exact pass, and a few ROH come out of the planted homozygous stretches."""
    names, offs, pos, cens = synth.make_positions_genomewide(41, 60000, n_chr=2, big_gap_frac=0.0)
    codes = synth.make_codes(41, 40, 60000, n_roh=3, roh_snps=(6000, 9000))

    class DS:
        pass
    ds = DS()
    ds.chr_names, ds.chr_offsets, ds.pos, ds.centromeres, ds.gl = names, offs, pos, {}, None
    ds.map_pos = [pos[offs[c]:offs[c + 1]][::5].astype(np.int64) for c in range(2)]
    ds.map_cm = [p * 1.2e-6 for p in ds.map_pos]
    hp = HotPath().load(ds, weighted=weighted, cm=weighted, error=0.001, packed_rows=synth.pack_codes(codes), max_gap=10 ** 9)
    for W in ((1600, 4096) if not weighted else (1600,)):
        if weighted:
            hp.g.ld_band(W, np.arange(0, 40, 4, dtype=np.int32))
        win = hp.g.windows(W, 1, weighted=weighted, individuals=np.array([1], np.int32))[0]
        v = np.sort(win[(win != orc.MISSING) & ~np.isnan(win)])
        cutoff = float(v[int(0.9 * len(v))])
        a = hp.g.call_roh(W, cutoff, 0.25, weighted=weighted, exact=False)
        b = hp.g.call_roh(W, cutoff, 0.25, weighted=weighted, exact=True)
        assert np.array_equal(a, b) and len(a) > 0
    hp.close()
