"""
The CUDA primitive path chained over four consecutive scans from an empty map (pytest -m gpu), against the golden
recorded from the reference's own ``process_scan_single_hypothesis`` (tests/golden/make_golden_pipeline.py): the fused
entry ``primitives.lidar_evidence_primitives`` and the hypothesis batch (H = 1 and H = 3: hypothesis 0 is the recorded
one) receive what the reference's loop handed to its operators -- stencil from the predicted pose, linearisation at
z_lin, map update at z_t -- and must reproduce every output of every scan, the map included.
Integer outputs bit-exact; floating outputs far inside the 1e-5 bar (the chain accumulates rounding from scan to scan).
"""
import sys

import numpy as np
import pytest

from conftest import GOLDEN, Gold, rel_err
from test_oracle_pipeline_chain import Scan, check_tiles

sys.path.insert(0, GOLDEN)
from pipeline_chain_inputs import scan_inputs  # noqa: E402

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")


def _np(t):
    return t.detach().cpu().numpy()


@pytest.fixture(scope="module")
def mods():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from gc_slam_b200 import hypothesis_batch, operators, primitives
    return primitives, operators, hypothesis_batch


def _check_scan(P, S, g, x, out_deskew, batch, inf, view, assoc, c_as, vpe, mu_res, c_mu, amap, active):
    dk = out_deskew
    S.close("dk_points", _np(dk.points), 1e-12); S.close("dk_weights", _np(dk.weights), 1e-12)
    assert batch.n_lidar_valid == int(S["mb_n_lidar"]) and batch.n_camera_valid == int(S["mb_n_cam"])
    S.eq("mb_valid", _np(batch.valid_mask).astype(bool)); S.eq("mb_sources", _np(batch.sources))
    S.eq("mb_source_indices", _np(batch.source_indices))
    for got, key in ((batch.Lambdas, "mb_Lambdas"), (batch.thetas, "mb_thetas"), (batch.etas, "mb_etas"),
                     (batch.weights, "mb_weights"), (batch.timestamps, "mb_timestamps"), (batch.colors, "mb_colors")):
        S.close(key, _np(got), 1e-7)
    got = np.array([inf.staleness_inflation_strength, inf.staleness_cov_inflation_trace, inf.stale_precision_downscale_total])
    assert rel_err(got, S["inf_stats"]) < 1e-10 or float(np.max(np.abs(S["inf_stats"]))) == 0.0 == float(np.max(np.abs(got)))
    vm = _np(view.valid_mask).astype(bool)
    S.eq("view_slots", _np(view.candidate_slots)); S.eq("view_tids", _np(view.candidate_tile_ids))
    S.eq("view_valid", vm); S.eq("view_ids", _np(view.primitive_ids))
    assert view.n_valid == int(S["view_n_valid"])
    S.close("view_pos", _np(view.positions) * vm[:, None], 1e-8)
    S.eq("as_pool", _np(assoc.candidate_pool_indices)); S.eq("as_tids", _np(assoc.candidate_tile_ids))
    S.eq("as_slots", _np(assoc.candidate_slots))
    S.close("as_cost", _np(assoc.cost_matrix), 1e-7); S.close("as_resp", _np(assoc.responsibilities), 1e-6)
    S.close("as_row", _np(assoc.row_masses), 1e-6)
    if int(S["as_has_ot"]):
        ot = c_as.ot
        got = np.array([ot.marginal_defect_a, ot.marginal_defect_b, ot.transport_mass_total, ot.sum_a, ot.sum_m, ot.sum_novel,
                        ot.p95_a, ot.nonzero_a, ot.b_recency_p95])
        assert np.max(np.abs(got - S["as_ot"]) / (np.abs(S["as_ot"]) + 1e-12)) < 1e-6
    else:
        assert c_as.ot is None
    assert rel_err(_np(vpe.L_pose), S["L22"]) < 1e-6 and rel_err(_np(vpe.h_pose), S["h22"]) < 1e-5
    assert abs(vpe.total_weighted_cost - float(S["vp_cost"])) <= 1e-7 * abs(float(S["vp_cost"]))
    assert vpe.n_associations == int(S["vp_n_assoc"])
    mu = c_mu.map_update
    assert mu_res.n_fused == int(S["fused_count"]) and mu_res.n_inserted == int(S["n_ins"]) and mu_res.n_culled == int(S["n_cull"])
    assert (mu.fused_count, mu.insert_count_total, mu.evicted_count) == tuple(int(v) for v in S["mu_counts"])
    for got, i in ((mu.fused_mass_total, 0), (mu.insert_mass_total, 1), (mu.insert_mass_p95, 2), (mu.evicted_mass_total, 3)):
        assert abs(got - float(S["mu_cert"][i])) <= 1e-8 * abs(float(S["mu_cert"][i])) + 1e-15, i
    assert np.array_equal(_np(mu_res.new_ids), S["new_ids"])
    assert amap.next_global_id == int(S["next_global_id"]) and amap.total_count == int(S["total_count"])
    check_tiles(S, active, amap.download_tile, counts=mu_res.tile_counts)


def _scan_args(P, ops, S, g, x):
    # step 3 (pipeline.py:436-483): the twist the loop deskews with, from the IMU window it saw
    from gc_slam_b200 import imu
    tw = imu.imu_scan_twist(x["imu_t"], x["gyro"], x["accel"], x["t0"], x["t1"], float(S["imu_sigma_warp"]), S["imu_rotvec0"],
                            S["imu_gyro_bias"], S["imu_accel_bias"], S["imu_gravity_W"])
    assert rel_err(_np(tw.xi_body)[0], S["xi"]) < 1e-9 and abs(float(tw.ess[0]) - float(S["ess_imu"])) < 1e-9 * float(S["ess_imu"])
    rs, _, _ = ops.point_budget_resample(x["points"], x["timestamps"], x["weights"], x["ring"], x["tag"], int(g["cap"]))
    S.close("rs_points", _np(rs.points), 1e-14); S.close("rs_weights", _np(rs.weights), 1e-14)
    cam = x["cam"]
    base = P.measurement_batch_from_camera_splats(cam["positions"], cam["covariances"], cam["directions"], cam["kappas"],
                                                  cam["weights"], cam["timestamps"], cam["colors"], int(g["n_feat"]), int(g["n_surfel"]))
    active = P.ma_hex_stencil_tile_ids(S["pose_pred"][:3], 2.0, 1, 0)
    assert active == [int(t) for t in S["active"]]
    return rs, base, active


def test_fused_entry_chain_matches_the_reference_loop(mods):
    P, ops, _ = mods
    G = Gold("pipeline_chain.npz")
    g = G.g
    amap = P.AtlasMap(m_tile=int(g["m_tile"]), n_tiles_cap=48)
    for k in range(1, int(g["n_scans"]) + 1):
        S, x = Scan(G, k), scan_inputs(k - 1)
        rs, base, active = _scan_args(P, ops, S, g, x)
        out = P.lidar_evidence_primitives(rs.points, rs.timestamps, rs.weights, x["t0"], x["t1"], S["xi"], amap, active, S["z_lin"],
                                          x["scan_seq"], base_batch=base, z_t=S["z_t"], ess_imu=float(S["ess_imu"]),
                                          m_tile_view=int(g["m_view"]), map_update_kwargs=dict(k_insert_tile=int(g["k_ins"])))
        mu_res, c_mu, _ = out["map_update"]
        _check_scan(P, S, g, x, out["deskew"][0], out["surfels"][0], out["recency_inflate"][3], out["map_view"],
                    out["association"][0], out["association"][1], out["pose_evidence"][0], mu_res, c_mu, amap, active)


@pytest.mark.parametrize("H", [1, 3])
def test_hypothesis_batch_chain_matches_the_reference_loop(mods, H):
    """Hypothesis 0 of the batch is the recorded hypothesis; the others (perturbed) must not disturb it or the map."""
    P, ops, HB = mods
    G = Gold("pipeline_chain.npz")
    g = G.g
    amap = P.AtlasMap(m_tile=int(g["m_tile"]), n_tiles_cap=64)
    for k in range(1, int(g["n_scans"]) + 1):
        S, x = Scan(G, k), scan_inputs(k - 1)
        rs, base, active = _scan_args(P, ops, S, g, x)
        d = np.zeros((H, 6)); d[1:, 0] = 0.05 * np.arange(1, H); d[1:, 5] = 0.01 * np.arange(1, H)
        xi = np.tile(S["xi"], (H, 1)) + 0.1 * d
        res = HB.lidar_evidence_primitives_batched(rs.points, rs.timestamps, rs.weights, x["t0"], x["t1"], xi, amap,
                                                   np.tile(S["pose_pred"], (H, 1)) + d, x["scan_seq"], base_batch=base,
                                                   m_tile_view=int(g["m_view"]), ess_imu=float(S["ess_imu"]),
                                                   map_update_kwargs=dict(k_insert_tile=int(g["k_ins"])),
                                                   z_lin_poses=np.tile(S["z_lin"], (H, 1)) + d, z_t=S["z_t"], update_map=True)
        u = res.unit(0)
        mu_res, c_mu, _ = u["map_update"]
        _check_scan(P, S, g, x, u["deskew"][0], u["surfels"][0], u["recency_inflate"][3], u["map_view"], u["association"][0],
                    u["association"][1], u["pose_evidence"][0], mu_res, c_mu, amap, active)
        assert res.L_pose.shape == (H, 22, 22) and bool(torch.isfinite(res.L_pose).all())
