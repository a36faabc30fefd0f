"""
The oracle's per-scan primitive path, chained over four consecutive scans from an empty map, against the golden recorded
from the reference's own ``process_scan_single_hypothesis`` (tests/golden/make_golden_pipeline.py: the unmodified loop of
fl/backend/pipeline.py run on the NumPy shim, its operator calls recorded).  This pins the GLUE -- the order of the
operators, which pose goes where (stencil centre = predicted pose, linearisation at z_lin, map update at z_t), the fuse
blocks, the insert scores, cull / forget -- that the per-operator goldens cannot see.
"""
import sys

import numpy as np

from conftest import GOLDEN, Gold, rel_err

sys.path.insert(0, GOLDEN)
from pipeline_chain_inputs import scan_inputs  # noqa: E402

from oracle import bin_path as ob  # noqa: E402
from oracle import imu as oi  # noqa: E402
from oracle import prim_path as op  # noqa: E402


class Scan:
    """Golden view of one scan of the chain: keys are prefixed s<k>_ in the file."""

    def __init__(self, G, k):
        self.G, self.p = G, f"s{k}_"

    def __getitem__(self, key):
        return self.G.g[self.p + key]

    def has_full(self, key):
        return self.G.has_full(self.p + key)

    def eq(self, key, arr):
        self.G.eq(self.p + key, arr)

    def close(self, key, arr, tol):
        self.G.close(self.p + key, arr, tol)


def check_tiles(S, active, tile_of, counts=None):
    for a, tid in enumerate(active):
        t = tile_of(tid)
        S.eq(f"tile{tid}_valid", t["valid_mask"]); S.eq(f"tile{tid}_ids", t["primitive_ids"])
        S.eq(f"tile{tid}_last", t["last_supported_scan_seq"])
        assert t["count"] == int(S[f"tile{tid}_count"])
        if counts is not None:
            assert counts[a] == t["count"]
        S.close(f"tile{tid}_weights", t["weights"], 1e-9)
        for f in ("Lambdas", "thetas", "etas", "timestamps", "rgb", "cam_mass", "lidar_mass"):
            assert rel_err(t[f].astype(np.float64).sum(axis=0), S[f"tile{tid}_{f}_sum"]) < 1e-8, (tid, f)


def test_oracle_chain_matches_the_reference_loop():
    G = Gold("pipeline_chain.npz")
    g = G.g
    atlas = op.create_empty_atlas(int(g["m_tile"]))
    for k in range(1, int(g["n_scans"]) + 1):
        S = Scan(G, k)
        x = scan_inputs(k - 1)
        # step 3 (pipeline.py:436-483): the twist the loop deskews with, from the IMU window it saw
        tw = oi.imu_scan_twist(x["imu_t"], x["gyro"], x["accel"], x["t0"], x["t1"], float(S["imu_sigma_warp"]), S["imu_rotvec0"],
                               S["imu_gyro_bias"], S["imu_accel_bias"], S["imu_gravity_W"])
        assert rel_err(tw["xi_body"], S["xi"]) < 1e-9 and abs(tw["ess"] - float(S["ess_imu"])) < 1e-9 * float(S["ess_imu"])
        # step 1 + 5 of the loop (pipeline.py:400-418, 569-587)
        rs, _ = ob.point_budget_resample(x["points"], x["timestamps"], x["weights"], x["ring"], x["tag"], int(g["cap"]))
        S.close("rs_points", rs["points"], 1e-14); S.close("rs_weights", rs["weights"], 1e-14)
        dk, _ = ob.deskew_constant_twist(rs["points"], rs["timestamps"], rs["weights"], x["t0"], x["t1"], S["xi"])
        S.close("dk_points", dk["points"], 1e-12); S.close("dk_weights", dk["weights"], 1e-12)
        cam = x["cam"]
        base = op.batch_from_camera_splats(cam["positions"], cam["covariances"], cam["directions"], cam["kappas"], cam["weights"],
                                           cam["timestamps"], cam["colors"], int(g["n_feat"]), int(g["n_surfel"]))
        batch, _, _ = op.extract_lidar_surfels(dk["points"], rs["timestamps"], dk["weights"], base, n_surfel=int(g["n_surfel"]),
                                               n_feat=int(g["n_feat"]))
        assert batch["n_lidar_valid"] == int(S["mb_n_lidar"]) and batch["n_camera_valid"] == int(S["mb_n_cam"])
        S.eq("mb_valid", batch["valid_mask"]); S.eq("mb_sources", batch["sources"]); S.eq("mb_source_indices", batch["source_indices"])
        for ko, kg in (("Lambdas", "mb_Lambdas"), ("thetas", "mb_thetas"), ("etas", "mb_etas"), ("weights", "mb_weights"),
                       ("timestamps", "mb_timestamps"), ("colors", "mb_colors")):
            S.close(kg, batch[ko], 1e-8)
        # stencil from the PREDICTED pose (pipeline.py:802-829), inflation, view
        active = op.stencil_tile_ids(S["pose_pred"][:3])
        assert active == [int(t) for t in S["active"]]
        atlas, inf = op.recency_inflate(atlas, active, x["scan_seq"])
        got = np.array([inf["staleness_inflation_strength"], inf["staleness_cov_inflation_trace"], inf["stale_precision_downscale_total"]])
        assert rel_err(got, S["inf_stats"]) < 1e-11 or float(np.max(np.abs(S["inf_stats"]))) == 0.0 == float(np.max(np.abs(got)))
        view = op.extract_atlas_map_view(atlas, active, int(g["m_view"]))
        S.eq("view_slots", view["candidate_slots"]); S.eq("view_tids", view["candidate_tile_ids"])
        S.eq("view_valid", view["valid_mask"]); S.eq("view_ids", view["primitive_ids"])
        assert int(np.sum(view["valid_mask"])) == int(S["view_n_valid"])
        S.close("view_pos", view["positions"] * view["valid_mask"][:, None], 1e-9)
        # association, evidence linearised at z_lin (pipeline.py:855-877, 998-1010)
        assoc, c_as = op.associate_primitives_ot(batch, view, scan_seq=x["scan_seq"])
        S.eq("as_pool", assoc["candidate_pool_indices"]); S.eq("as_tids", assoc["candidate_tile_ids"])
        S.eq("as_slots", assoc["candidate_slots"])
        # (the chain accumulates rounding from scan to scan: 2e-9 on the costs of scan 4; the bar is 1e-5)
        S.close("as_cost", assoc["cost_matrix"], 1e-8); S.close("as_resp", assoc["responsibilities"], 1e-7)
        S.close("as_row", assoc["row_masses"], 1e-7)
        if int(S["as_has_ot"]):
            ot = np.array([c_as["marginal_defect_a"], c_as["marginal_defect_b"], c_as["transport_mass_total"], c_as["sum_a"],
                           c_as["sum_m"], c_as["sum_novel"], c_as["p95_a"], c_as["nonzero_a"], c_as["b_recency_p95"]])
            assert np.max(np.abs(ot - S["as_ot"]) / (np.abs(S["as_ot"]) + 1e-12)) < 1e-8
        vpe, _ = op.visual_pose_evidence(assoc, batch, view, S["z_lin"])
        assert rel_err(vpe["L_pose"], S["vp_L"]) < 1e-8 and rel_err(vpe["h_pose"], S["vp_h"]) < 1e-7
        assert abs(vpe["total_weighted_cost"] - float(S["vp_cost"])) <= 1e-8 * abs(float(S["vp_cost"]))
        # what the loop hands to the fusion (build_visual_pose_evidence_22d)
        assert rel_err(vpe["L_pose"], S["L22"]) < 1e-8 and rel_err(vpe["h_pose"], S["h22"]) < 1e-7
        # step 12b at z_t (pipeline.py:1233-1447)
        atlas, st = op.map_update(atlas, batch, assoc, active, S["z_t"], x["scan_seq"], x["t1"], k_insert=int(g["k_ins"]))
        assert st["fused_count"] == int(S["fused_count"]) == int(S["mu_counts"][0]) == int(S["res_counts"][0])
        assert st["insert_count_total"] == int(S["n_ins"]) == int(S["mu_counts"][1])
        assert st["evicted_count"] == int(S["n_cull"]) == int(S["mu_counts"][2]) and int(S["n_merged"]) == 0
        for name, i in (("fused_mass_total", 0), ("insert_mass_total", 1), ("insert_mass_p95", 2), ("evicted_mass_total", 3)):
            assert abs(st[name] - float(S["mu_cert"][i])) <= 1e-9 * abs(float(S["mu_cert"][i])) + 1e-15, name
        assert np.array_equal(np.stack(st["new_ids"]), S["new_ids"])
        assert atlas["next_global_id"] == int(S["next_global_id"]) and atlas["total_count"] == int(S["total_count"])
        check_tiles(S, active, lambda tid: atlas["tiles"][tid])
