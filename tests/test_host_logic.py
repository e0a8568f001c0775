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
