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
