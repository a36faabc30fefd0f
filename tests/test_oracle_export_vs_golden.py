"""oracle/export.py against the vectors produced by the reference's own extract_primitive_map_view,
renderable_batch_from_view and _build_pointcloud2_from_view (tests/golden/make_golden_export.py).  CPU only."""
import glob
import os

import numpy as np
import pytest

from conftest import GOLDEN, golden, rel_err

EXPORT_CASES = sorted(os.path.basename(p) for p in glob.glob(os.path.join(GOLDEN, "export_*.npz")))


def atlas_of(g):
    from gc_slam_b200 import synth
    atl = synth.synthetic_atlas(int(g["n_surf"]), int(g["m_tile"]), int(g["seed"]), scan_seq=30)
    if int(np.min(g["slot_counts"])) == 0:
        atl["tiles"][sorted(atl["tiles"].keys())[0]]["valid_mask"][:] = False
    return atl


def test_cases_present():
    assert len(EXPORT_CASES) >= 2


@pytest.mark.parametrize("case", EXPORT_CASES)
def test_oracle_export_matches_reference(case):
    from oracle import export as oe
    g = golden(case)
    o = oe.export_map_points(atlas_of(g))
    assert np.array_equal(o["primitive_ids"], g["primitive_ids"]) and np.array_equal(o["last_supported_scan_seq"], g["last_supported_scan_seq"])
    assert np.array_equal(o["cloud"], g["cloud"]) and int(g["point_step"]) == 16 and int(g["width"]) == o["mu_world"].shape[0]
    assert rel_err(o["mu_world"], g["mu_world"]) < 1e-13 and rel_err(o["Sigma_world"], g["Sigma_world"]) < 1e-13
    assert rel_err(o["Lambda_world"], g["Lambda_world"]) < 1e-12 and np.array_equal(o["eta"], g["eta"])
    assert np.array_equal(o["mass"], g["mass"]) and np.array_equal(o["color"], g["color"])
    # newest first, ties by primitive id
    rec, pid = o["last_supported_scan_seq"], o["primitive_ids"]
    assert np.all((rec[:-1] > rec[1:]) | ((rec[:-1] == rec[1:]) & (pid[:-1] < pid[1:])))
