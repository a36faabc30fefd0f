"""oracle/prim_path.py's per-tile map operators against the vectors produced by the reference's own
primitive_map_fuse / insert_masked / cull / forget and block_associations_for_fuse
(tests/golden/make_golden_mapops.py).  CPU only."""
import glob
import os

import numpy as np
import pytest

from conftest import GOLDEN, golden, rel_err

FIELDS = ("Lambdas", "thetas", "etas", "weights", "timestamps", "created_timestamps", "last_supported_scan_seq",
          "last_update_scan_seq", "primitive_ids", "valid_mask", "colors", "cam_mass", "lidar_mass", "rgb_cam_accum",
          "rgb_cam_denom", "rgb")
EXACT = ("timestamps", "created_timestamps", "last_supported_scan_seq", "last_update_scan_seq", "primitive_ids", "valid_mask")
FLOAT = tuple(k for k in FIELDS if k not in EXACT)

FUSE_CASES = sorted(os.path.basename(p) for p in glob.glob(os.path.join(GOLDEN, "mapops_fuse_*.npz")))
INSERT_CASES = sorted(os.path.basename(p) for p in glob.glob(os.path.join(GOLDEN, "mapops_insert_*.npz")))
CULL_CASES = sorted(os.path.basename(p) for p in glob.glob(os.path.join(GOLDEN, "mapops_cull_*.npz")))


def atlas_of(g, next_global_id=10 ** 6):
    t = {k: np.array(g["in_" + k]) for k in FIELDS}
    t.update(tile_id=int(g["tile_id"]), count=int(g["count_in"]), next_local_id=int(g["next_local_id"]))
    return dict(tiles={int(g["tile_id"]): t}, next_global_id=next_global_id, total_count=int(g["count_in"]),
                m_tile=int(t["weights"].shape[0]))


def check_tile(out, g, tol):
    for k in EXACT:
        assert np.array_equal(out[k], g["out_" + k]), k
    for k in FLOAT:
        assert rel_err(out[k], g["out_" + k]) < tol, k


def fuse_args(g):
    full = bool(g["full"])
    return dict(slots=g["slots"], lam=g["lam"], th=g["th"], eta=g["eta"], w=g["w"], resp=g["resp"],
                vm=g["vm"] if full else None, col=g["col"] if full else None, src=g["src"] if full else None)


def test_cases_present():
    assert len(FUSE_CASES) >= 2 and len(INSERT_CASES) >= 3 and len(CULL_CASES) >= 3


@pytest.mark.parametrize("case", FUSE_CASES)
def test_oracle_fuse_matches_reference(case):
    from oracle import prim_path as op
    g = golden(case)
    a = fuse_args(g)
    atl = atlas_of(g)
    n = op.primitive_map_fuse(atl, int(g["tile_id"]), a["slots"], a["lam"], a["th"], a["eta"], a["w"], a["resp"],
                              float(g["timestamp"]), int(g["scan_seq"]), a["vm"], a["col"], a["src"])
    assert n == int(g["n_fused"]) and float(g["realized"]) == float(n) and float(g["predicted"]) == float(len(a["slots"]))
    check_tile(atl["tiles"][int(g["tile_id"])], g, 1e-13)


@pytest.mark.parametrize("case", INSERT_CASES)
def test_oracle_insert_masked_matches_reference(case):
    from oracle import prim_path as op
    g = golden(case)
    full = bool(g["full"])
    atl = atlas_of(g, int(g["next_global_id_in"]))
    n, ids, slots = op.primitive_map_insert_masked(atl, int(g["tile_id"]), g["lam"], g["th"], g["eta"], g["w"], float(g["timestamp"]),
                                                   g["vnew"], int(g["scan_seq"]), _lambda(),
                                                   g["col"] if full else None, g["src"] if full else None)
    assert n == int(g["n_inserted"]) and np.array_equal(ids, g["new_ids"])
    assert atl["next_global_id"] == int(g["next_global_id"]) and atl["total_count"] == int(g["total_count"])
    t = atl["tiles"][int(g["tile_id"])]
    assert t["count"] == int(g["count_out"])
    check_tile(t, g, 1e-15)
    assert bool(g["exact"]) == bool(np.all(g["vnew"]))
    if not bool(g["exact"]):
        assert [str(x) for x in g["triggers"]] == ["insert_unfilled_budget"]


def _lambda():
    from gc_slam_b200 import constants
    return constants.GC_RECENCY_DECAY_LAMBDA


@pytest.mark.parametrize("case", CULL_CASES)
def test_oracle_cull_matches_reference(case):
    from oracle import prim_path as op
    g = golden(case)
    atl = atlas_of(g)
    maxp = None if int(g["maxp"]) < 0 else int(g["maxp"])
    n, mass = op.primitive_map_cull(atl, int(g["tile_id"]), float(g["thr"]), maxp)
    assert n == int(g["n_culled"]) and abs(mass - float(g["mass_dropped"])) <= 1e-13 * max(1.0, abs(mass))
    assert atl["total_count"] == int(g["total_count"])
    t = atl["tiles"][int(g["tile_id"])]
    assert t["count"] == int(g["count_out"])
    check_tile(t, g, 1e-15)
    assert bool(g["exact"]) == (n == 0)
    if n:
        from gc_slam_b200 import constants
        assert [str(x) for x in g["triggers"]] == ["budgeting", "mass_drop"]
        assert abs(float(g["mass_epsilon_ratio"]) - mass / (np.sum(g["in_weights"]) + constants.GC_EPS_MASS)) < 1e-14


def test_oracle_forget_matches_reference():
    from oracle import prim_path as op
    g = golden("mapops_forget.npz")
    atl = atlas_of(g)
    op.primitive_map_forget(atl, int(g["tile_id"]), float(g["gamma"]))
    check_tile(atl["tiles"][int(g["tile_id"])], g, 1e-16)
    assert bool(g["exact"]) and abs(float(g["predicted"]) - (1.0 - float(g["gamma"]))) < 1e-16


def test_oracle_block_associations_match_reference():
    from oracle import prim_path as op
    g = golden("mapops_block_assoc.npz")
    assoc = dict(responsibilities=g["responsibilities"], candidate_tile_ids=g["candidate_tile_ids"], candidate_slots=g["candidate_slots"])
    mi, ct, cs, rs, vr = op.block_associations_for_fuse(assoc, g["valid_mask"], int(g["block"]))
    assert np.array_equal(mi, g["out_meas_idx"]) and np.array_equal(ct, g["out_tile_ids"]) and np.array_equal(cs, g["out_slots"])
    assert np.array_equal(rs, g["out_resp"]) and np.array_equal(vr, g["out_valid_rows"])
