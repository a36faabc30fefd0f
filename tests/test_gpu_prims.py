"""
GPU parity tests of the primitive family (pytest -m gpu): surfel extraction, map view, OT association, pose evidence,
map update -- all through the C-ABI, checked against the reference-generated golden vectors and the NumPy oracle.

Bars (BASELINE.json north_star): bucket contents, surfel order, view slots, candidate pool indices / tile ids / slots,
eviction slots and new primitive ids bit-exact; floating results relative 1e-5 (asserted far tighter: float64 path);
rotation geodesic < 1e-6 rad.
"""
import glob
import os

import numpy as np
import pytest

from conftest import GOLDEN, Gold, geodesic, golden, rel_err
from test_oracle_prim_vs_golden import build_inputs

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")

CASES = sorted(os.path.basename(p) for p in glob.glob(os.path.join(GOLDEN, "prim*_*.npz")))
ESS, SUP = 1, 2


def _np(t):
    return t.detach().cpu().numpy()


@pytest.fixture(scope="module")
def P():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from gc_slam_b200 import primitives
    return primitives


def _gpu_batch_and_view(P, g, rs, dk, atlas_np):
    from gc_slam_b200 import synth
    cfg = P.SurfelExtractionConfig(n_surfel=int(g["n_surfel"]), n_feat=int(g["n_feat"]))
    base = None
    if int(g["n_cam"]):
        cam = synth.camera_splats(int(g["n_cam"]), int(g["seed"]) + 3)
        base = P.measurement_batch_from_camera_splats(cam["positions"], cam["covariances"], cam["directions"], cam["kappas"],
                                                      cam["weights"], cam["timestamps"], cam["colors"], cfg.n_feat, cfg.n_surfel)
    batch, c_sf, e_sf = P.extract_lidar_surfels(dk["points"], rs["timestamps"], dk["weights"], cfg, base, return_bucket=True)
    amap = P.AtlasMap.from_numpy(atlas_np)
    return batch, c_sf, amap


@pytest.mark.parametrize("case", CASES)
def test_primitive_path_vs_golden(P, case):
    G = Gold(case)
    g = G.g
    rs, dk, _, atlas_np = build_inputs(g)
    batch, c_sf, amap = _gpu_batch_and_view(P, g, rs, dk, atlas_np)
    # ---- a10
    G.eq("bucket", _np(batch._bucket)); G.eq("bucket_count", _np(batch._bucket_count))
    assert batch.n_lidar_valid == int(g["mb_n_lidar"]) and batch.n_camera_valid == int(g["mb_n_cam"])
    assert np.array_equal(_np(batch.valid_mask).astype(bool), g["mb_valid"])
    assert np.array_equal(_np(batch.sources), g["mb_sources"]) and np.array_equal(_np(batch.source_indices), g["mb_source_indices"])
    for got, key in ((batch.Lambdas, "mb_Lambdas"), (batch.thetas, "mb_thetas"), (batch.etas, "mb_etas"),
                     (batch.weights, "mb_weights"), (batch.timestamps, "mb_timestamps"), (batch.colors, "mb_colors")):
        G.close(key, _np(got), 1e-7)
    assert c_sf.support.ess_total == g["sf_cert"][ESS] and c_sf.exact is False
    # ---- a11
    scan_seq = int(g["scan_seq"])
    active = [int(x) for x in g["active"]]
    assert P.ma_hex_stencil_tile_ids(g["pose"][:3], 2.0, 1, 0) == active
    amap, _, _, inf = P.primitive_map_recency_inflate(amap, active, scan_seq)
    got = np.array([inf.staleness_inflation_strength, inf.staleness_cov_inflation_trace, inf.stale_precision_downscale_total])
    assert rel_err(got, g["inf_stats"]) < 1e-11
    view = P.extract_atlas_map_view(amap, active, int(g["m_view"]))
    G.eq("view_slots", _np(view.candidate_slots)); G.eq("view_tids", _np(view.candidate_tile_ids))
    vm = _np(view.valid_mask).astype(bool)
    G.eq("view_valid", vm); G.eq("view_ids", _np(view.primitive_ids))
    G.close("view_pos", _np(view.positions) * vm[:, None], 1e-8); G.close("view_cov", _np(view.covariances) * vm[:, None, None], 1e-8)
    G.close("view_dir", _np(view.directions), 1e-12); G.close("view_kappa", _np(view.kappas), 1e-12)
    # ---- a12
    assoc, c_as, e_as = P.associate_primitives_ot(batch, view, P.AssociationConfig(scan_seq=scan_seq))
    pool = _np(assoc.candidate_pool_indices)
    if G.has_full("as_pool"):
        n_mismatch = int(np.sum(np.any(pool != g["as_pool"], axis=1)))
        assert n_mismatch == 0, f"{n_mismatch} rows with a different candidate set"
    G.eq("as_pool", pool)
    G.eq("as_tids", _np(assoc.candidate_tile_ids)); G.eq("as_slots", _np(assoc.candidate_slots))
    G.close("as_cost", _np(assoc.cost_matrix), 1e-8)
    G.close("as_resp", _np(assoc.responsibilities), 1e-8); G.close("as_row", _np(assoc.row_masses), 1e-8)
    ot = c_as.ot
    got = np.array([ot.marginal_defect_a, ot.marginal_defect_b, ot.transport_mass_total, ot.sum_a, ot.sum_m, ot.sum_novel,
                    ot.p95_a, ot.nonzero_a, ot.b_recency_p95])
    assert np.max(np.abs(got - g["as_ot"]) / (np.abs(g["as_ot"]) + 1e-12)) < 1e-7
    assert abs(c_as.support.ess_total - g["as_cert"][ESS]) < 1e-7 * g["as_cert"][ESS]
    assert abs(e_as.predicted - float(g["as_effect"])) < 1e-7 * abs(float(g["as_effect"])) + 1e-13
    assert c_as.compute.largest_tensor_shape == (batch.n_total, 8) and c_as.compute.segment_sum_k == 8
    if case.startswith("prim_p1"):       # MeasurementMassPolicy.WEIGHT_PROPORTIONAL (primitive_association.py:416-421)
        wg = golden("assoc_wprop_p1.npz")
        wa, wc, we = P.associate_primitives_ot(batch, view, P.AssociationConfig(scan_seq=scan_seq,
                                                                              a_policy=P.MeasurementMassPolicy.WEIGHT_PROPORTIONAL))
        assert np.array_equal(_np(wa.candidate_pool_indices), wg["pool"]) and rel_err(_np(wa.cost_matrix), wg["cost"]) < 1e-8
        assert rel_err(_np(wa.responsibilities), wg["resp"]) < 1e-8 and rel_err(_np(wa.row_masses), wg["row"]) < 1e-8
        wot = np.array([wc.ot.marginal_defect_a, wc.ot.marginal_defect_b, wc.ot.transport_mass_total, wc.ot.sum_a, wc.ot.sum_m,
                        wc.ot.sum_novel, wc.ot.p95_a, wc.ot.nonzero_a, wc.ot.b_recency_p95])
        assert np.max(np.abs(wot - wg["ot"]) / (np.abs(wg["ot"]) + 1e-12)) < 1e-7
        assert abs(we.predicted - float(wg["effect"])) < 1e-7 * abs(float(wg["effect"])) + 1e-13
        assert abs(wc.support.support_frac - wg["cert"][SUP]) < 1e-12
        for kk in (4, 16):               # other candidate counts K_ASSOC (AssociationConfig.k_assoc): association, evidence, update
            kg = golden(f"assoc_k{kk}_p1.npz")
            ka, kc, ke = P.associate_primitives_ot(batch, view, P.AssociationConfig(scan_seq=scan_seq, k_assoc=kk))
            assert np.array_equal(_np(ka.candidate_pool_indices), kg["pool"]) and np.array_equal(_np(ka.candidate_slots), kg["slots"])
            assert rel_err(_np(ka.cost_matrix), kg["cost"]) < 1e-8 and rel_err(_np(ka.responsibilities), kg["resp"]) < 1e-8
            assert rel_err(_np(ka.row_masses), kg["row"]) < 1e-8
            kot = np.array([kc.ot.marginal_defect_a, kc.ot.marginal_defect_b, kc.ot.transport_mass_total, kc.ot.sum_a, kc.ot.sum_m,
                            kc.ot.sum_novel, kc.ot.p95_a, kc.ot.nonzero_a, kc.ot.b_recency_p95])
            assert np.max(np.abs(kot - kg["ot"]) / (np.abs(kg["ot"]) + 1e-12)) < 1e-7
            assert abs(ke.predicted - float(kg["effect"])) < 1e-7 * abs(float(kg["effect"])) + 1e-13
            assert kc.compute.largest_tensor_shape == (batch.n_total, kk) and kc.compute.segment_sum_k == kk
            kv, _, _ = P.visual_pose_evidence(ka, batch, view, g["pose"], z_lin_pose=g["pose"])
            assert rel_err(_np(kv.L_pose), kg["vp_L"]) < 1e-7 and rel_err(_np(kv.h_pose), kg["vp_h"]) < 1e-6
            assert abs(kv.total_weighted_cost - float(kg["vp_cost"])) < 1e-7 * abs(float(kg["vp_cost"]))
            # the map update with that K against the oracle's (on copies of the map)
            from oracle import prim_path as op
            from test_oracle_prim_vs_golden import build_inputs as _bi
            o_atlas, _ = op.recency_inflate(_bi(g)[3], active, scan_seq)
            o_assoc = dict(candidate_tile_ids=_np(ka.candidate_tile_ids), candidate_slots=_np(ka.candidate_slots),
                           responsibilities=_np(ka.responsibilities), row_masses=_np(ka.row_masses))
            o_batch = {f: _np(getattr(batch, f)) for f in ("Lambdas", "thetas", "etas", "weights", "colors", "sources")}
            o_batch["valid_mask"] = _np(batch.valid_mask).astype(bool)
            o2, st = op.map_update(o_atlas, o_batch, o_assoc, active, g["z_t"], scan_seq, float(g["ts"]), k_insert=int(g["k_ins"]))
            m2 = P.AtlasMap.from_numpy(atlas_np)
            m2, _, _, _ = P.primitive_map_recency_inflate(m2, active, scan_seq)
            r2, c2, _ = P.map_update_step12b(m2, batch, ka, active, g["z_t"], scan_seq, float(g["ts"]), k_insert_tile=int(g["k_ins"]))
            assert (r2.n_fused, r2.n_inserted, r2.n_culled) == (st["fused_count"], st["insert_count_total"], st["evicted_count"])
            assert np.array_equal(_np(r2.new_ids), np.stack(st["new_ids"]))
            for tid in active:
                t2 = m2.download_tile(tid)
                assert np.array_equal(t2["valid_mask"], o2["tiles"][tid]["valid_mask"]) and rel_err(t2["weights"], o2["tiles"][tid]["weights"]) < 1e-9
                assert rel_err(t2["Lambdas"].sum(axis=0), o2["tiles"][tid]["Lambdas"].sum(axis=0)) < 1e-8
    # ---- a13
    vpe, c_vp, _ = P.visual_pose_evidence(assoc, batch, view, g["pose"], z_lin_pose=g["pose"])
    assert rel_err(_np(vpe.L_pose), g["vp_L"]) < 1e-7 and rel_err(_np(vpe.h_pose), g["vp_h"]) < 1e-6
    assert abs(vpe.total_weighted_cost - float(g["vp_cost"])) < 1e-7 * abs(float(g["vp_cost"]))
    assert abs(vpe.mean_transported_mass - float(g["vp_mean_mass"])) < 1e-8 * abs(float(g["vp_mean_mass"]))
    assert c_vp.frobenius_applied is True and c_vp.approximation_triggers == ["linearization", "ot_soft_correspondence"]
    # ---- a14
    res, c_mu, _ = P.map_update_step12b(amap, batch, assoc, active, g["z_t"], scan_seq, float(g["ts"]),
                                        k_insert_tile=int(g["k_ins"]))
    mu = c_mu.map_update
    assert res.n_fused == int(g["fused_count"]) and res.n_inserted == int(g["n_ins"]) and res.n_culled == int(g["n_cull"])
    assert abs(mu.fused_mass_total - float(g["fused_mass"])) < 1e-9 * abs(float(g["fused_mass"])) + 1e-15
    assert abs(mu.evicted_mass_total - float(g["m_cull"])) < 1e-9 * abs(float(g["m_cull"])) + 1e-15
    assert np.array_equal(_np(res.new_ids), g["new_ids"])
    assert amap.next_global_id == int(g["next_global_id"]) and amap.total_count == int(g["total_count"])
    for a, tid in enumerate(active):
        t = amap.download_tile(tid)
        G.eq(f"tile{tid}_valid", t["valid_mask"]); G.eq(f"tile{tid}_ids", t["primitive_ids"])
        G.eq(f"tile{tid}_last", t["last_supported_scan_seq"])
        assert t["count"] == int(g[f"tile{tid}_count"]) == res.tile_counts[a]
        G.close(f"tile{tid}_weights", t["weights"], 1e-10)
        for f in ("Lambdas", "thetas", "etas", "timestamps", "rgb", "cam_mass", "lidar_mass"):
            assert rel_err(t[f].astype(np.float64).sum(axis=0), g[f"tile{tid}_{f}_sum"]) < 1e-8, f


def test_full_size_vs_oracle_and_determinism(P):
    """Reference budgets (n_feat 512, n_surfel 1024, m_tile 50,000, view 1024, K_INSERT 64) on a 65,536-point scan."""
    from gc_slam_b200 import synth
    from oracle import bin_path as ob
    from oracle import prim_path as op
    n = 65536
    pts, t, w, ring, tag = synth.vlp16_scan(n, 31, t0=synth.EPOCH_T0)
    xi = synth.scan_twist(31)
    dk, _ = ob.deskew_constant_twist(pts, t, w, synth.EPOCH_T0, synth.EPOCH_T0 + 0.1, xi)
    cam = synth.camera_splats(200, 77)
    atlas_np = synth.synthetic_atlas(400000, 50000, 9, scan_seq=20)
    pose = np.array([0.1, -0.2, 0.0, 0.0, 0.0, 0.05])
    active = op.stencil_tile_ids(pose[:3])
    # oracle
    ob_base = op.batch_from_camera_splats(cam["positions"], cam["covariances"], cam["directions"], cam["kappas"],
                                          cam["weights"], cam["timestamps"], cam["colors"])
    o_batch, o_aux, _ = op.extract_lidar_surfels(dk["points"], t, dk["weights"], ob_base)
    o_atlas, _ = op.recency_inflate(atlas_np, active, 21)
    o_view = op.extract_atlas_map_view(o_atlas, active)
    o_assoc, _ = op.associate_primitives_ot(o_batch, o_view, scan_seq=21)
    o_vpe, _ = op.visual_pose_evidence(o_assoc, o_batch, o_view, pose)
    o_atlas2, o_st = op.map_update(o_atlas, o_batch, o_assoc, active, pose, 21, synth.EPOCH_T0 + 0.1)

    def run():
        base = P.measurement_batch_from_camera_splats(cam["positions"], cam["covariances"], cam["directions"], cam["kappas"],
                                                      cam["weights"], cam["timestamps"], cam["colors"])
        batch, _, _ = P.extract_lidar_surfels(dk["points"], t, dk["weights"], P.SurfelExtractionConfig(), base, return_bucket=True)
        amap = P.AtlasMap.from_numpy(atlas_np)
        amap, _, _, _ = P.primitive_map_recency_inflate(amap, active, 21)
        view = P.extract_atlas_map_view(amap, active, 1024)
        assoc, c_as, _ = P.associate_primitives_ot(batch, view, P.AssociationConfig(scan_seq=21))
        vpe, _, _ = P.visual_pose_evidence(assoc, batch, view, pose, z_lin_pose=pose)
        res, c_mu, _ = P.map_update_step12b(amap, batch, assoc, active, pose, 21, synth.EPOCH_T0 + 0.1)
        return batch, view, assoc, vpe, res, amap

    batch, view, assoc, vpe, res, amap = run()
    assert np.array_equal(_np(batch._bucket), o_aux["bucket"])
    assert batch.n_lidar_valid == o_batch["n_lidar_valid"] == 1024
    assert rel_err(_np(batch.Lambdas), o_batch["Lambdas"]) < 1e-7 and rel_err(_np(batch.etas), o_batch["etas"]) < 1e-8
    assert np.array_equal(_np(view.candidate_slots), o_view["candidate_slots"])
    pool = _np(assoc.candidate_pool_indices)
    bad = int(np.sum(np.any(pool != o_assoc["candidate_pool_indices"], axis=1)))
    assert bad == 0, f"{bad} of {pool.shape[0]} rows differ in their candidate set"
    assert rel_err(_np(assoc.responsibilities), o_assoc["responsibilities"]) < 1e-7
    assert rel_err(_np(vpe.L_pose), o_vpe["L_pose"]) < 1e-7 and rel_err(_np(vpe.h_pose)[:3], o_vpe["h_pose"][:3]) < 1e-6
    assert res.n_inserted == o_st["insert_count_total"] and res.n_culled == o_st["evicted_count"] and res.n_fused == o_st["fused_count"]
    assert np.array_equal(_np(res.new_ids), np.stack(o_st["new_ids"]))
    assert np.array_equal(_np(res.insert_slots), np.stack(o_st["insert_slots"]))
    for tid in active:
        tt, ot = amap.download_tile(tid), o_atlas2["tiles"][tid]
        assert np.array_equal(tt["valid_mask"], ot["valid_mask"]) and np.array_equal(tt["primitive_ids"], ot["primitive_ids"])
        assert rel_err(tt["Lambdas"], ot["Lambdas"]) < 1e-9 and rel_err(tt["thetas"], ot["thetas"]) < 1e-9
        assert rel_err(tt["weights"], ot["weights"]) < 1e-12 and np.array_equal(tt["timestamps"], ot["timestamps"])
        assert rel_err(tt["rgb"], ot["rgb"]) < 1e-12
    # determinism: a second run from the same inputs is bit-identical
    batch2, view2, assoc2, vpe2, res2, amap2 = run()
    assert torch.equal(assoc.responsibilities, assoc2.responsibilities) and torch.equal(vpe.L_pose, vpe2.L_pose)
    for name in ("Lambdas", "thetas", "weights", "etas"):
        assert torch.equal(amap.fields[name], amap2.fields[name]), name


def test_empty_and_error_cases(P):
    from gc_slam_b200 import synth
    # empty map: association returns the fixed-shape zero result and an exact certificate
    pts, t, w, _, _ = synth.vlp16_scan(4096, 3, t0=0.0)
    batch, _, _ = P.extract_lidar_surfels(pts, t, w, P.SurfelExtractionConfig(n_surfel=128, n_feat=16))
    amap = P.create_empty_atlas_map(m_tile=2048, n_tiles_cap=8)
    tiles = P.ma_hex_stencil_tile_ids(np.zeros(3))
    view = P.extract_atlas_map_view(amap, tiles, 64)
    assert view.n_valid == 0 and int(view.valid_mask.sum().item()) == 0
    assoc, cert, eff = P.associate_primitives_ot(batch, view)
    assert cert.exact and float(assoc.responsibilities.abs().sum().item()) == 0.0 and eff.predicted == 0.0
    vpe, c2, _ = P.visual_pose_evidence(assoc, batch, view, np.zeros(6))
    assert c2.exact and float(vpe.h_pose.abs().sum().item()) == 0.0
    # first map update on an empty atlas: tiles get created, inserts happen (incl. zero-mass placeholders, quirk Q6)
    from oracle import prim_path as op
    ob = dict(Lambdas=_np(batch.Lambdas), thetas=_np(batch.thetas), etas=_np(batch.etas), weights=_np(batch.weights),
              sources=_np(batch.sources), valid_mask=_np(batch.valid_mask).astype(bool), colors=_np(batch.colors),
              n_feat=batch.n_feat, n_surfel=batch.n_surfel)
    oa = dict(responsibilities=_np(assoc.responsibilities), candidate_tile_ids=_np(assoc.candidate_tile_ids),
              candidate_slots=_np(assoc.candidate_slots), row_masses=_np(assoc.row_masses))
    o_atlas, o_st = op.map_update(op.create_empty_atlas(2048), ob, oa, tiles, np.zeros(6), 1, 0.1, k_insert=16)
    res, c_mu, _ = P.map_update_step12b(amap, batch, assoc, tiles, np.zeros(6), 1, 0.1, k_insert_tile=16)
    assert len(amap.tiles) == 7 and c_mu.map_update.n_active_tiles == 7
    assert res.n_inserted == o_st["insert_count_total"] > 0 and np.array_equal(_np(res.new_ids), np.stack(o_st["new_ids"]))
    for tid in tiles:
        assert np.array_equal(amap.download_tile(tid)["valid_mask"], o_atlas["tiles"][tid]["valid_mask"])
    with pytest.raises(ValueError):
        P.extract_atlas_map_view(amap, tiles, 0)
    with pytest.raises(ValueError):
        P.associate_primitives_ot(batch, view, P.AssociationConfig(b_policy=P.MapMassPolicy.PRIMITIVE_MASS))
    with pytest.raises(ValueError):
        P.extract_lidar_surfels(pts[:10], t, w)


def test_fused_entry_is_bit_identical_to_operator_sequence(P):
    """lidar_evidence_primitives (two host synchronisations) vs the seven stand-alone operators (seven): same C entry
    points in the same order, so every output, certificate scalar and the updated map must be bit-identical; and the
    empty-map early exits must be taken the same way."""
    from gc_slam_b200 import operators as ops, synth
    n = 30000
    pts, t, w, _, _ = synth.vlp16_scan(n, 41, t0=synth.EPOCH_T0)
    xi = synth.scan_twist(41)
    cam = synth.camera_splats(150, 78)
    atlas_np = synth.synthetic_atlas(120000, 50000, 10, scan_seq=20)
    pose = np.array([0.1, -0.2, 0.0, 0.0, 0.0, 0.05])
    active = P.ma_hex_stencil_tile_ids(pose[:3])
    t0, t1 = synth.EPOCH_T0, synth.EPOCH_T0 + 0.1

    def base():
        return P.measurement_batch_from_camera_splats(cam["positions"], cam["covariances"], cam["directions"], cam["kappas"],
                                                      cam["weights"], cam["timestamps"], cam["colors"])

    # operator by operator
    amap_a = P.AtlasMap.from_numpy(atlas_np)
    dk, c_dk, _ = ops.deskew_constant_twist(pts, t, w, t0, t1, xi, 1.0, "GC-RIGHT-01", "a")
    batch, c_sf, _ = P.extract_lidar_surfels(dk.points, dk.timestamps, dk.weights, None, base())
    amap_a, _, _, inf = P.primitive_map_recency_inflate(amap_a, active, 21)
    view = P.extract_atlas_map_view(amap_a, active, 1024)
    assoc, c_as, e_as = P.associate_primitives_ot(batch, view, P.AssociationConfig(scan_seq=21))
    vpe, c_pe, _ = P.visual_pose_evidence(assoc, batch, view, pose, z_lin_pose=pose)
    res, c_mu, _ = P.map_update_step12b(amap_a, batch, assoc, active, pose, 21, t1, inflate_stats=inf)
    # fused
    amap_b = P.AtlasMap.from_numpy(atlas_np)
    out = P.lidar_evidence_primitives(pts, t, w, t0, t1, xi, amap_b, active, pose, 21, base_batch=base())
    f_dk, f_cdk, _ = out["deskew"]
    f_batch, f_csf, _ = out["surfels"]
    f_assoc, f_cas, f_eas = out["association"]
    f_vpe, f_cpe, _ = out["pose_evidence"]
    f_res, f_cmu, _ = out["map_update"]
    assert torch.equal(f_dk.points, dk.points) and f_cdk.support.support_frac == c_dk.support.support_frac
    assert f_batch.n_lidar_valid == batch.n_lidar_valid and torch.equal(f_batch.Lambdas, batch.Lambdas)
    assert f_csf.support.ess_total == c_sf.support.ess_total
    assert out["map_view"].n_valid == view.n_valid and torch.equal(out["map_view"].candidate_slots, view.candidate_slots)
    assert torch.equal(f_assoc.responsibilities, assoc.responsibilities)
    assert torch.equal(f_assoc.candidate_pool_indices, assoc.candidate_pool_indices)
    assert f_cas.ot.transport_mass_total == c_as.ot.transport_mass_total and f_eas.predicted == e_as.predicted
    assert torch.equal(f_vpe.L_pose, vpe.L_pose) and torch.equal(f_vpe.h_pose, vpe.h_pose)
    assert f_cpe.support.ess_total == c_pe.support.ess_total
    assert (f_res.n_fused, f_res.n_inserted, f_res.n_culled) == (res.n_fused, res.n_inserted, res.n_culled)
    assert torch.equal(f_res.new_ids, res.new_ids) and amap_b.next_global_id == amap_a.next_global_id
    assert f_cmu.map_update.fused_mass_total == c_mu.map_update.fused_mass_total
    assert f_cmu.map_update.staleness_inflation_strength == c_mu.map_update.staleness_inflation_strength
    for name in amap_a.fields:
        assert torch.equal(amap_a.fields[name], amap_b.fields[name]), name
    # the fused call counts two host synchronisations (first operator of each group), the operators seven
    syncs = lambda cs: sum(c.compute.device_runtime.host_sync_count_est for c in cs)
    assert syncs([c_dk, c_sf, c_as, c_pe, c_mu]) == 5 and syncs([f_cdk, f_csf, f_cas, f_cpe, f_cmu]) == 2

    # empty map: association / pose evidence take the reference's early exits inside the fused call too
    amap_e = P.create_empty_atlas_map(m_tile=2048, n_tiles_cap=8)
    out_e = P.lidar_evidence_primitives(pts, t, w, t0, t1, xi, amap_e, active, pose, 1,
                                        surfel_config=P.SurfelExtractionConfig(n_surfel=128, n_feat=16), m_tile_view=64,
                                        map_update_kwargs=dict(k_insert_tile=16))
    a_e, c_e, _ = out_e["association"]
    assert c_e.exact and float(a_e.responsibilities.abs().sum().item()) == 0.0
    assert out_e["pose_evidence"][1].exact and len(amap_e.tiles) == 7 and out_e["map_update"][0].n_inserted > 0
