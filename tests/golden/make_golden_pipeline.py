"""
Reference-DRIVEN golden of the per-scan primitive path: the reference's own ``process_scan_single_hypothesis``
(fl/backend/pipeline.py:316-1600, unmodified, executed on the NumPy JAX shim -- see make_golden.py) is run for three
consecutive scans of one synthetic room, map and belief carried from scan to scan, and what ITS operator calls received
and returned is recorded by wrapping the names the function looks up in its module (the function body is untouched).

Why: the other primitive goldens re-drive the glue between the operators (which operator is called in which order with
which arguments: pipeline.py:778-877, 998-1010, 1233-1447) from a restatement; here the reference's loop itself decides
the stencil, the twist, the linearisation pose, the update pose z_t, the fuse blocks, the insert scores and the
cull / forget order.  Reference budgets (n_feat 512, n_surfel 1,024, m_tile 50,000, view 1,024, K_INSERT 64, cap 8,192).

Inputs of the path that the pipeline derives from sensors outside this path (IMU preintegration -> twist, IMU + odometry
evidence -> linearisation pose, fused belief -> z_t) are stored as the small vectors the reference computed; the scans
and camera splats are regenerated from gc_slam_b200.synth by seed.  Large outputs are stored in the compact form of
conftest.check_compact.  The tests chain the three scans through the oracle (CPU) and through the CUDA path (GPU),
starting from an empty map.

    python tests/golden/make_golden_pipeline.py        (about 15 s)
"""
import hashlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("GCS_REFERENCE", "/root/reference")
sys.path.insert(0, os.path.join(ROOT, "oracle", "jax_shim"))
sys.path.insert(0, os.path.join(REF, "fl_ws", "src", "fl_slam_poc"))
sys.path.insert(0, ROOT)

sys.path.insert(0, HERE)
from pipeline_chain_inputs import N_RAW, SCANS, FULL_STRIDE, X_ANCHOR, scan_inputs  # noqa: E402

SPIED = ["point_budget_resample", "deskew_constant_twist", "extract_lidar_surfels", "primitive_map_recency_inflate",
         "extract_atlas_map_view", "associate_primitives_ot", "visual_pose_evidence", "build_visual_pose_evidence_22d",
         "pose_update_frobenius_recompose", "primitive_map_fuse", "primitive_map_insert_masked", "primitive_map_cull",
         "primitive_map_forget", "primitive_map_merge_reduce", "ma_hex_stencil_tile_ids", "block_associations_for_fuse",
         "smooth_window_weights", "preintegrate_imu_relative_pose_jax"]


def A(x):
    return np.asarray(x)


def main():
    import jax.numpy as jnp
    from fl_slam_poc.backend import pipeline as PL
    from fl_slam_poc.backend.operators.inverse_wishart_jax import process_noise_state_to_Q_jax
    from fl_slam_poc.backend.operators.measurement_noise_iw_jax import measurement_noise_mean_jax
    from fl_slam_poc.backend.structures import (create_datasheet_measurement_noise_state,
                                                create_datasheet_process_noise_state)
    from fl_slam_poc.backend.structures import primitive_map as pm
    from fl_slam_poc.backend.structures.measurement_batch import measurement_batch_from_camera_splats
    from fl_slam_poc.common.belief import BeliefGaussianInfo

    from gc_slam_b200 import synth

    log = []

    def spy(name):
        fn = getattr(PL, name)

        def wrapped(*a, **kw):
            r = fn(*a, **kw)
            log.append((name, a, kw, r))
            return r
        wrapped.__name__ = name
        return wrapped

    for nm in SPIED:
        setattr(PL, nm, spy(nm))

    cfg = PL.PipelineConfig()
    mns = create_datasheet_measurement_noise_state(lidar_sigma_meas=0.01)
    cfg.Sigma_meas = measurement_noise_mean_jax(mns, idx=2)
    cfg.Sigma_g = measurement_noise_mean_jax(mns, idx=0)
    cfg.Sigma_a = measurement_noise_mean_jax(mns, idx=1)
    cfg.lidar_origin_base = jnp.asarray(synth.lidar_origin_base())
    Q = process_noise_state_to_Q_jax(create_datasheet_process_noise_state())
    # a prior away from the tile boundaries at the origin, tight enough that the synthetic IMU / odometry do not throw the
    # pose across tiles from scan to scan (the scans re-observe the same room, so later scans associate with the map)
    belief = BeliefGaussianInfo.create_prior(anchor_id="hyp_0_anchor_0", X_anchor=jnp.asarray(X_ANCHOR), stamp_sec=0.0,
                                             mean=jnp.zeros(22), cov=1e-4 * jnp.eye(22))
    amap = pm.create_empty_atlas_map(m_tile=cfg.primitive_map_max_size)

    out = dict(n_raw=N_RAW, cap=int(cfg.N_POINTS_CAP), n_feat=int(cfg.n_feat), n_surfel=int(cfg.n_surfel),
               m_tile=int(cfg.primitive_map_max_size), m_view=int(cfg.M_TILE_VIEW), k_ins=int(cfg.k_insert_tile),
               n_scans=len(SCANS), full_stride=FULL_STRIDE)
    for k in range(len(SCANS)):
        x = scan_inputs(k)
        cam = x["cam"]
        cb = measurement_batch_from_camera_splats(jnp.asarray(cam["positions"]), jnp.asarray(cam["covariances"]),
                                                  jnp.asarray(cam["directions"]), jnp.asarray(cam["kappas"]),
                                                  jnp.asarray(cam["weights"]), jnp.asarray(cam["timestamps"]),
                                                  jnp.asarray(cam["colors"]), n_feat=cfg.n_feat, n_surfel=cfg.n_surfel)
        del log[:]
        res = PL.process_scan_single_hypothesis(
            belief_prev=belief, raw_points=jnp.asarray(x["points"]), raw_timestamps=jnp.asarray(x["timestamps"]),
            raw_weights=jnp.asarray(x["weights"]), raw_ring=jnp.asarray(x["ring"]), raw_tag=jnp.asarray(x["tag"]),
            imu_stamps=jnp.asarray(x["imu_t"]), imu_gyro=jnp.asarray(x["gyro"]), imu_accel=jnp.asarray(x["accel"]),
            odom_pose=jnp.asarray(x["odom_pose"]), odom_cov_se3=1e-2 * jnp.eye(6), scan_start_time=x["t0"], scan_end_time=x["t1"],
            dt_sec=0.1, t_last_scan=x["t0"] - 0.1, t_scan=x["t1"], Q=Q, config=cfg, odom_twist=jnp.asarray(x["odom_twist"]),
            odom_twist_cov=1e-2 * jnp.eye(6), camera_batch=cb, scan_seq=x["scan_seq"], primitive_map=amap)
        calls = {}
        for name, a, kw, r in log:
            calls.setdefault(name, []).append((a, kw, r))
        p = f"s{k + 1}_"
        o = {}
        # ---- what the loop handed to the operators of this path
        (_, kw, r), = calls["deskew_constant_twist"]
        o["xi"] = A(kw["xi_body"]); o["ess_imu"] = float(kw["ess_imu"])
        o["rs_points"] = A(kw["points"]); o["rs_weights"] = A(kw["weights"]); o["rs_timestamps"] = A(kw["timestamps"])
        o["dk_points"] = A(r[0].points); o["dk_weights"] = A(r[0].weights)
        # step 3 (pipeline.py:436-483): the within-scan IMU window and preintegration that produced this twist
        kw_w = calls["smooth_window_weights"][0][1]
        kw_p = calls["preintegrate_imu_relative_pose_jax"][0][1]
        assert float(kw_w["scan_start_time"]) == x["t0"] and float(kw_w["scan_end_time"]) == x["t1"]
        o["imu_sigma_warp"] = float(kw_w["sigma"]); o["imu_rotvec0"] = A(kw_p["rotvec_start_WB"])
        o["imu_gyro_bias"] = A(kw_p["gyro_bias"]); o["imu_accel_bias"] = A(kw_p["accel_bias"]); o["imu_gravity_W"] = A(kw_p["gravity_W"])
        (_, kw, r), = calls["extract_lidar_surfels"]
        batch = r[0]
        assert kw["base_batch"] is cb and A(kw["timestamps"]).shape == o["rs_timestamps"].shape
        stc = calls["ma_hex_stencil_tile_ids"]
        assert len(stc) == 2 and list(stc[0][2]) == list(stc[1][2])       # active == stencil at the compiled radii
        active = [int(t) for t in stc[0][2]]
        o["stencil_center"] = A(stc[0][1]["center_xyz"]); o["active"] = np.asarray(active, np.int64)
        (_, kw, r), = calls["primitive_map_recency_inflate"]
        assert [int(t) for t in kw["tile_ids"]] == active and int(kw["scan_seq"]) == x["scan_seq"]
        inf = r[3]
        o["inf_stats"] = np.array([inf.staleness_inflation_strength, inf.staleness_cov_inflation_trace,
                                   inf.stale_precision_downscale_total])
        (_, kw, r), = calls["extract_atlas_map_view"]
        view = r
        (_, kw, r), = calls["associate_primitives_ot"]
        assoc, c_as, e_as = r
        assert kw["measurement_batch"] is batch and kw["map_view"] is view
        (_, kw, r), = calls["visual_pose_evidence"]
        vpe = r[0]
        o["z_lin"] = A(kw["z_lin_pose"]).reshape(-1)[:6]
        o["pose_pred"] = A(kw["belief_pred"].mean_world_pose(eps_lift=cfg.eps_lift)).reshape(-1)[:6]
        assert np.allclose(o["pose_pred"][:3], o["stencil_center"])
        (_, kw, r), = calls["build_visual_pose_evidence_22d"]
        o["L22"] = A(r[0]); o["h22"] = A(r[1])
        (_, kw, r), = calls["pose_update_frobenius_recompose"]
        o["z_t"] = A(r[1].mean_world_pose(eps_lift=cfg.eps_lift)).reshape(-1)[:6]
        # ---- measurement batch, view, association, pose evidence
        o.update(mb_Lambdas=A(batch.Lambdas), mb_thetas=A(batch.thetas), mb_etas=A(batch.etas), mb_weights=A(batch.weights),
                 mb_sources=A(batch.sources), mb_source_indices=A(batch.source_indices), mb_valid=A(batch.valid_mask),
                 mb_timestamps=A(batch.timestamps), mb_colors=A(batch.colors), mb_n_lidar=batch.n_lidar_valid,
                 mb_n_cam=batch.n_camera_valid,
                 view_slots=A(view.candidate_slots), view_tids=A(view.candidate_tile_ids), view_valid=A(view.valid_mask),
                 view_pos=A(view.positions) * A(view.valid_mask)[:, None], view_w=A(view.weights), view_ids=A(view.primitive_ids),
                 view_n_valid=int(np.sum(A(view.valid_mask))),
                 as_resp=A(assoc.responsibilities), as_pool=A(assoc.candidate_pool_indices), as_tids=A(assoc.candidate_tile_ids),
                 as_slots=A(assoc.candidate_slots), as_row=A(assoc.row_masses), as_cost=A(assoc.cost_matrix),
                 as_effect=e_as.predicted,
                 as_has_ot=int(c_as.ot is not None),      # None: the empty-view early exit (primitive_association.py:352-389)
                 as_ot=(np.array([c_as.ot.marginal_defect_a, c_as.ot.marginal_defect_b, c_as.ot.transport_mass_total,
                                  c_as.ot.sum_a, c_as.ot.sum_m, c_as.ot.sum_novel, c_as.ot.p95_a, c_as.ot.nonzero_a,
                                  c_as.ot.b_recency_p95]) if c_as.ot is not None else np.zeros(9)),
                 vp_L=A(vpe.L_pose), vp_h=A(vpe.h_pose), vp_cost=vpe.total_weighted_cost, vp_mean_mass=vpe.mean_transported_mass,
                 vp_n_assoc=int(vpe.n_associations))
        # ---- step 12b as the loop ran it
        fuse = calls.get("primitive_map_fuse", [])
        ins = calls.get("primitive_map_insert_masked", [])
        cull = calls.get("primitive_map_cull", [])
        assert [int(c[1]["tile_id"]) for c in ins] == active and [int(c[1]["tile_id"]) for c in cull] == active
        assert all(float(c[1]["timestamp"]) == x["t1"] and int(c[1]["scan_seq"]) == x["scan_seq"] for c in fuse + ins)
        mu_cert = [c for c in res.all_certs if getattr(c, "anchor_id", "") == "map_update"][-1].map_update
        o.update(n_fuse_calls=len(fuse), fused_count=int(sum(c[2][0].n_fused for c in fuse)),
                 n_ins=int(sum(int(c[2][0].n_inserted) for c in ins)), new_ids=np.stack([A(c[2][0].new_ids) for c in ins]),
                 n_cull=int(sum(int(c[2][0].n_culled) for c in cull)), m_cull=float(sum(float(c[2][0].mass_dropped) for c in cull)),
                 n_merged=int(res.n_primitives_merged),
                 res_counts=np.array([res.n_primitives_fused, res.n_primitives_inserted, res.n_primitives_culled]),
                 mu_cert=np.array([getattr(mu_cert, f) for f in ("fused_mass_total", "insert_mass_total", "insert_mass_p95",
                                                                  "evicted_mass_total")], np.float64),
                 mu_counts=np.array([getattr(mu_cert, f) for f in ("fused_count", "insert_count_total", "evicted_count")], np.int64))
        amap = res.primitive_map_updated
        belief = res.belief_updated
        o["next_global_id"] = int(amap.next_global_id); o["total_count"] = int(amap.total_count)
        for tid in active:
            tl = amap.tiles[int(tid)]
            for f in ("Lambdas", "thetas", "etas", "timestamps", "rgb", "cam_mass", "lidar_mass"):
                o[f"tile{tid}_{f}_sum"] = A(getattr(tl, f)).astype(np.float64).sum(axis=0)
            o[f"tile{tid}_weights"] = A(tl.weights); o[f"tile{tid}_valid"] = A(tl.valid_mask)
            o[f"tile{tid}_ids"] = A(tl.primitive_ids); o[f"tile{tid}_last"] = A(tl.last_supported_scan_seq)
            o[f"tile{tid}_count"] = int(tl.count)
        print(f"scan {k + 1}: stencil {active}, n_lidar {batch.n_lidar_valid}, view valid {o['view_n_valid']}, "
              f"fused {o['fused_count']} ({o['n_fuse_calls']} calls), inserted {o['n_ins']}, culled {o['n_cull']}, "
              f"map total {o['total_count']}, |xi| {np.linalg.norm(o['xi']):.4f}, z_t {np.round(o['z_t'], 4)}")
        for key, v in o.items():
            a = np.asarray(v)
            if a.nbytes <= 32768:
                out[p + key] = a
                continue
            a = np.ascontiguousarray(a)
            out[p + key + "_rows"] = a[::FULL_STRIDE]
            out[p + key + "_sha256"] = np.frombuffer(hashlib.sha256(a.tobytes()).digest(), np.uint8)
            out[p + key + "_sum"] = a.astype(np.float64).sum(axis=0)
            out[p + key + "_abssum"] = np.abs(a.astype(np.float64)).sum(axis=0)
    path = os.path.join(HERE, "pipeline_chain.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
