"""oracle/hypothesis.py against the vectors produced by the reference's own _hypothesis_barycenter_core
(tests/golden/make_golden_hyp.py; fl/backend/operators/hypothesis.py:51-115).  CPU only."""
import glob
import os

import numpy as np
import pytest

from conftest import GOLDEN, golden, rel_err

HYP_CASES = sorted(os.path.basename(p) for p in glob.glob(os.path.join(GOLDEN, "hyp_*.npz")))


def test_cases_present():
    assert len(HYP_CASES) >= 3


@pytest.mark.parametrize("case", HYP_CASES)
def test_oracle_hypothesis_barycenter_matches_reference(case):
    from oracle import hypothesis as oh
    g = golden(case)
    o = oh.hypothesis_barycenter(g["L_stack"], g["h_stack"], g["z_lin_stack"], g["weights"], float(g["weight_floor"]),
                                 float(g["eps_psd"]), float(g["eps_lift"]))
    assert np.array_equal(o["weights_normalized"], g["weights_normalized"]) and o["floor_adjustment"] == g["floor_adjustment"]
    assert rel_err(o["L"], g["L"]) < 1e-13 and rel_err(o["h"], g["h"]) < 1e-14 and rel_err(o["z_lin"], g["z_lin"]) < 1e-14
    assert rel_err(o["psd_cert"][2:5], g["psd_cert"][2:5]) < 1e-9 and o["psd_cert"][5] == g["psd_cert"][5]
    assert abs(o["psd_cert"][0] - g["psd_cert"][0]) < 1e-9 * (1.0 + np.max(np.abs(g["L"]))) * 1e-3
    # the per-hypothesis means solve systems with condition numbers up to 1e9 (1e17 in the degenerate case, where the
    # lift decides): the spread diagnostic agrees to cond * 1e-16
    assert abs(o["spread_proxy"] - g["spread_proxy"]) < 1e-6 * abs(g["spread_proxy"])
    # property the reference's tests pin for the PSD projection (test/test_primitives.py:53-94): no negative direction
    # beyond what float64 resolves for a matrix of this norm
    assert np.linalg.eigvalsh(0.5 * (o["L"] + o["L"].T)).min() >= -1e-14 * np.abs(o["L"]).max()
