"""Host-side (non-GPU) logic of the product mirrored from the reference: map interpolation, heuristics,
overlap threshold.  Checked against the oracle and the reference binary's logged values."""
import numpy as np

from garlic_b200 import pipeline
from garlic_b200.api import overlap_threshold
from oracle import oracle as orc
from tests.common import load_case, log_value


def test_interpolate_map_matches_oracle():
    ds, args = load_case("wlod_cm")
    res = orc.run_pipeline(ds, 25, 0.001, None, weighted=False, cm=True)
    for c, ch in enumerate(res["chroms"]):
        g, k = pipeline.interpolate_map(ch["pos"], ds.map_pos[c], ds.map_cm[c])
        assert np.array_equal(g, ch["gpos"])
    assert int(log_value("wlod_cm", "Number of genetic map locations interpolated:")) == res["n_interp"]


def test_density_and_overlap_fraction_match_reference_log():
    ds, args = load_case("auto_overlap_hg19")
    res = orc.run_pipeline(ds, 60, 0.001, None, auto_overlap=True, keep_windows=False)
    d = pipeline.calc_density(res["n_used"], [c["pos"] for c in res["chroms"]], [c["cen"] for c in res["chroms"]])
    assert d == res["density"]
    f = pipeline.select_overlap_frac(d, 60)
    assert "%g" % f == log_value("auto_overlap_hg19", "Selected overlap fraction:")
    assert f == res["overlap_frac"]
    assert pipeline.select_winsize_weighted(3.6e-4) == orc.lib().orc_select_winsize_weighted(3.6e-4) == 72


def test_overlap_threshold_clamp():
    assert overlap_threshold(0.25, 50) == 13      # 12.5 → SNP counts are integers
    assert overlap_threshold(0.0, 50) == 1
    assert overlap_threshold(1.0, 50) == 50
    assert overlap_threshold(2.0, 50) == 50
    assert overlap_threshold(0.3, 25) == 8


def test_bad_pair_bit_map_reproduces_the_reference_missing_pattern():
    """The closed form pass 1 tests on the device (a bit per SNP pair that no window may span: gap, centromere, chromosome
    start) against the MISSING pattern of the oracle's windows (calcLOD, garlic-roh.cpp:55-123), for data with gaps,
    centromeres strictly between / swallowing SNPs, hg19 centromeres, and several window sizes."""
    import pytest
    for name, W in (("lod_2", 50), ("lod_small", 10), ("lod_small", 2), ("auto_overlap_hg19", 60), ("lod_0", 25), ("lod_3", 33)):
        ds, args = load_case(name)
        res = orc.run_pipeline(ds, W, 0.001, None)
        chr_off = np.concatenate([[0], np.cumsum([len(c["pos"]) for c in res["chroms"]])])
        pos = np.concatenate([c["pos"] for c in res["chroms"]])
        bad = pipeline.bad_pair_bits(pos, chr_off, [c["cen"] for c in res["chroms"]], 200000)
        ok = pipeline.window_valid(bad, W)
        want = np.concatenate([c["win"][0] != orc.MISSING for c in res["chroms"]])
        assert len(want) == len(ok)
        assert np.array_equal(ok, want), (name, W)
        assert (~ok).sum() > 0


def test_decimal_fast_path_is_exact_or_hands_over():
    """K0-GL's conversion rule (csrc/ingest.cu:parse_decimal, restated in pipeline.parse_decimal_fast): whatever it does not
    hand to strtod it converts exactly — checked against Python's correctly rounded float() on 200,000 random tokens of the
    shapes likelihood files hold, and on the edge cases of the rule."""
    rng = np.random.default_rng(17)
    n_fast = 0
    for _ in range(200000):
        kind = rng.integers(0, 8)
        if kind == 0:
            tok = str(int(rng.integers(0, 10 ** int(rng.integers(1, 17)))))
        elif kind == 1:
            tok = "%.*f" % (int(rng.integers(0, 12)), rng.uniform(-300, 300))
        elif kind == 2:
            tok = "%.*e" % (int(rng.integers(0, 17)), rng.uniform(-1, 1) * 10.0 ** int(rng.integers(-30, 30)))
        elif kind == 3:
            tok = "%d.%0*d" % (rng.integers(0, 1000), int(rng.integers(1, 18)), rng.integers(0, 10 ** 9))
        elif kind == 4:
            tok = "%.17g" % rng.uniform(0, 1)
        elif kind == 5:
            tok = "0.%s%d" % ("0" * int(rng.integers(0, 20)), rng.integers(1, 10 ** 6))
        elif kind == 6:
            tok = "%de%d" % (rng.integers(0, 10 ** 15), rng.integers(-25, 25))
        else:
            tok = "%s%d%s" % (["", "+", "-"][int(rng.integers(0, 3))], rng.integers(0, 256), ["", ".", ".0", ".50"][int(rng.integers(0, 4))])
        v, hard = pipeline.parse_decimal_fast(tok)
        if not hard:
            n_fast += 1
            assert v == float(tok) and np.signbit(v) == np.signbit(float(tok)), tok
    assert n_fast > 100000
    for tok, want_hard in (("0", False), ("-0.0", False), ("150", False), ("0.0004", False), ("1e22", False), ("1e23", True),
                           ("123456789012345", False), ("1234567890123456", True), ("9007199254740993", True), ("nan", True),
                           ("inf", True), ("0x10", True), ("1e", True), (".", True), ("", True), ("1.5x", True), ("1e-22", False),
                           ("0.000000000000000000000001", True), ("5e-324", True)):
        v, hard = pipeline.parse_decimal_fast(tok)
        assert hard == want_hard, tok
        if not hard:
            assert v == float(tok), tok
