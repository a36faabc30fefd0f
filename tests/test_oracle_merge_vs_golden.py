"""oracle/merge.py against the vectors produced by the reference's own primitive_map_merge_reduce
(tests/golden/make_golden_merge.py), including the tile of the reference's known-answer test
(test/test_primitive_map_merge_reduce.py:77-99).  CPU only."""
import glob
import os

import numpy as np
import pytest

from conftest import GOLDEN, golden, rel_err

MERGE_CASES = sorted(os.path.basename(p) for p in glob.glob(os.path.join(GOLDEN, "merge_*.npz")))
FIELDS = ("Lambdas", "thetas", "etas", "weights", "timestamps", "created_timestamps", "last_supported_scan_seq",
          "last_update_scan_seq", "primitive_ids", "valid_mask", "colors", "cam_mass", "lidar_mass", "rgb_cam_accum",
          "rgb_cam_denom", "rgb")
EXACT = ("timestamps", "created_timestamps", "last_supported_scan_seq", "last_update_scan_seq", "primitive_ids", "valid_mask",
         "cam_mass", "lidar_mass", "rgb_cam_accum", "rgb_cam_denom", "weights")


def tile_of(g):
    t = {k: np.array(g["in_" + k]) for k in FIELDS}
    t.update(tile_id=int(g["tile_id"]), count=int(g["count_in"]), next_local_id=int(g["next_local_id"]))
    return t


def kwargs_of(g):
    return dict(merge_threshold=float(g["kw_merge_threshold"]), max_pairs=int(g["kw_max_pairs"]),
                max_tile_size=int(g["kw_max_tile_size"]))


def test_cases_present():
    assert len(MERGE_CASES) >= 4


@pytest.mark.parametrize("case", MERGE_CASES)
def test_oracle_merge_matches_reference(case):
    from oracle import merge as om
    g = golden(case)
    t, n, status = om.merge_reduce_tile(tile_of(g), **kwargs_of(g))
    assert n == int(g["n_merged"])
    assert status == ("budget_cap" if "merge_reduce_budget_cap" in list(g["triggers"]) else ("merged" if n else "noop"))
    for k in EXACT:
        assert np.array_equal(t[k], g["out_" + k]), k
    for k in ("Lambdas", "thetas", "etas", "colors", "rgb"):
        assert rel_err(t[k], g["out_" + k]) < 1e-12, k


def test_reference_known_answer():
    """test/test_primitive_map_merge_reduce.py:77-99: the close pair merges into slot 0 with weight 2, slot 1 is freed."""
    from oracle import merge as om
    g = golden("merge_reference_test_tile.npz")
    t, n, _ = om.merge_reduce_tile(tile_of(g), merge_threshold=0.5, max_pairs=1, max_tile_size=10)
    assert n == 1 and bool(t["valid_mask"][0]) and not bool(t["valid_mask"][1]) and np.isclose(t["weights"][0], 2.0)
    assert int(g["total_count"]) == 2 and bool(g["frobenius_applied"]) and float(g["realized"]) == 1.0
