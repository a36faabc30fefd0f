"""
CPU: the NumPy oracle against the golden vectors produced by the reference's own source
(tests/golden/make_golden.py).  This is what pins the oracle.
"""
import glob
import os

import numpy as np
import pytest

from conftest import GOLDEN, geodesic, golden, rel_err
from oracle import bin_path as ob
from oracle import lie

BIN_CASES = sorted(os.path.basename(p) for p in glob.glob(os.path.join(GOLDEN, "bin_*.npz"))
                   if not p.endswith("bin_scalars.npz"))

# cert_scalars layout in make_golden.py
EXACT, ESS, SUP, PSD, MEPS, EMIN, EMAX, COND, NNULL, NLL, DSCORE = range(11)


def _raw(g):
    from gc_slam_b200 import synth
    return synth.vlp16_scan(int(g["n_raw"]), int(g["seed"]), t0=float(g["t0"]))


@pytest.mark.parametrize("case", BIN_CASES)
def test_bin_path_matches_reference(case):
    g = golden(case)
    pts, t, w, ring, tag = _raw(g)
    cap = int(g["cap"])
    rs, c_rs = ob.point_budget_resample(pts, t, w, ring, tag, cap)
    # integer / byte outputs: bit exact
    assert rs["n_output"] == int(g["rs_n_output"])
    assert np.array_equal(rs["ring"], g["rs_ring"]) and np.array_equal(rs["tag"], g["rs_tag"])
    assert np.array_equal(rs["points"], g["rs_points"]) and np.array_equal(rs["timestamps"], g["rs_t"])
    assert rel_err(rs["weights"], g["rs_w"]) < 1e-14
    assert abs(c_rs["ess_total"] - g["rs_cert"][ESS]) <= 1e-12 * g["rs_cert"][ESS]
    assert abs(c_rs["support_frac"] - g["rs_cert"][SUP]) < 1e-15
    assert abs(c_rs["mass_epsilon_ratio"] - g["rs_cert"][MEPS]) <= 1e-12 * g["rs_cert"][MEPS]

    dk, c_dk = ob.deskew_constant_twist(rs["points"], rs["timestamps"], rs["weights"], float(g["t0"]),
                                        float(g["t1"]), g["xi"])
    assert np.max(np.abs(dk["points"] - g["dk_points"])) <= 1e-9 * max(1.0, np.max(np.abs(g["dk_points"])))
    n_sel = rs["n_output"]
    assert rel_err(dk["points"][:n_sel], g["dk_points"][:n_sel]) < 1e-13
    assert rel_err(dk["weights"], g["dk_w"]) < 1e-13
    assert abs(c_dk["support_frac"] - g["dk_cert"][SUP]) < 1e-13

    dirs = ob.ray_directions(dk["points"], g["origin"])
    assert np.max(np.abs(dirs - g["dirs"])) < 1e-12
    sa, c_sa = ob.bin_soft_assign(dirs, g["bin_dirs"], float(g["tau"]))
    resp = sa["responsibilities"]
    assert np.max(np.abs(resp[:: max(1, cap // 64)] - g["resp_rows"])) < 1e-12
    assert rel_err(resp.sum(0), g["resp_colsum"]) < 1e-12
    assert abs(c_sa["ess_total"] - g["sa_cert"][ESS]) < 1e-11 * g["sa_cert"][ESS]
    assert abs(c_sa["support_frac"] - g["sa_cert"][SUP]) < 1e-13
    assert abs(c_sa["effect_predicted"] - float(g["sa_effect"])) < 1e-12

    st, c_st = ob.scan_bin_moment_match(dk["points"], None, dk["weights"], resp, direction_origin=g["origin"])
    for k_o, k_g in (("N", "st_N"), ("s_dir", "st_s_dir"), ("S_dir_scatter", "st_S"), ("p_bar", "st_p_bar"),
                     ("Sigma_p", "st_Sigma_p"), ("kappa_scan", "st_kappa")):
        assert rel_err(st[k_o], g[k_g]) < 1e-10, k_o
    assert abs(c_st["ess_total"] - g["st_cert"][ESS]) < 1e-10 * g["st_cert"][ESS]
    assert abs(c_st["support_frac"] - g["st_cert"][SUP]) < 1e-12
    # psd_projection_delta of an already-PSD matrix is pure round-off noise: bound it, do not match it
    assert c_st["psd_projection_delta"] < 1e-9 and g["st_cert"][PSD] < 1e-9

    ms = {k[4:]: g[k] for k in g.files if k.startswith("map_") and k[4:] in
          ("S_dir", "S_dir_scatter", "N_dir", "N_pos", "sum_p", "sum_ppT")}
    mu_dir, kap, cen, Sc = ob.map_derived_stats(ms)
    assert rel_err(mu_dir, g["map_mu_dir"]) < 1e-12 and rel_err(kap, g["map_kappa"]) < 1e-11
    assert rel_err(cen, g["map_centroid"]) < 1e-12 and rel_err(Sc, g["map_Sigma_c"]) < 1e-10
    fg = ob.apply_forgetting(ms, 0.99)
    assert rel_err(fg["N_dir"], g["map_forgot_N_dir"]) < 1e-15
    assert rel_err(fg["sum_ppT"], g["map_forgot_sum_ppT"]) < 1e-15
    # update_map_stats (archive/bin_atlas.py:137-165) on the increments the generator recorded
    up = ob.update_map_stats(ms, g["st_s_dir"], g["st_S"], g["st_N"], g["upd_inc_N_pos"], g["upd_inc_sum_p"], g["upd_inc_sum_ppT"])
    for k in ("S_dir", "S_dir_scatter", "N_dir", "N_pos", "sum_p", "sum_ppT"):
        assert np.array_equal(up[k], g["upd_" + k]), k

    R_pred = lie.so3_exp(g["pose"][3:6])
    mf, c_mf = ob.matrix_fisher_rotation(R_pred, st["s_dir"], st["S_dir_scatter"], st["N"], ms["S_dir"],
                                         ms["S_dir_scatter"], ms["N_dir"])
    assert geodesic(mf["R_mf"], g["mf_R"]) < 1e-9
    assert rel_err(mf["svd_singular_values"], g["mf_s"]) < 1e-10
    assert rel_err(mf["L_rot"], g["mf_L"]) < 1e-9 and rel_err(mf["h_rot"], g["mf_h"]) < 1e-8
    assert np.max(np.abs(mf["delta_rot"] - g["mf_delta"])) < 1e-10
    assert abs(c_mf["cond"] - g["mf_cert"][COND]) < 1e-8 * g["mf_cert"][COND]
    assert abs(c_mf["nll_per_ess"] - g["mf_cert"][NLL]) < 1e-8 * abs(g["mf_cert"][NLL]) + 1e-15
    assert abs(c_mf["directional_score"] - g["mf_cert"][DSCORE]) < 1e-10 * g["mf_cert"][DSCORE]
    sm = mf["scan_scatter_metrics"]
    got = np.array([sm["linearity"], sm["planarity"], sm["sphericity"], sm["anisotropy"], sm["effective_rank"]])
    assert np.max(np.abs(got - g["mf_scan_metrics"])) < 1e-10
    mm = mf["map_scatter_metrics"]
    got = np.array([mm["linearity"], mm["planarity"], mm["sphericity"], mm["anisotropy"], mm["effective_rank"]])
    assert np.max(np.abs(got - g["mf_map_metrics"])) < 1e-10

    pt, c_pt = ob.planar_translation(g["pose"][:3], st["p_bar"], st["Sigma_p"], st["N"], cen, Sc, ms["N_pos"],
                                     ms["S_dir_scatter"], ms["N_dir"], mf["R_mf"])
    assert rel_err(pt["t_wls"], g["pt_t"]) < 1e-8
    assert rel_err(pt["L_trans"], g["pt_L"]) < 1e-8 and rel_err(pt["h_trans"], g["pt_h"]) < 1e-7
    assert abs(c_pt["nll_per_ess"] - g["pt_cert"][NLL]) < 1e-7 * abs(g["pt_cert"][NLL])
    L, h = ob.combined_lidar_evidence_22d(pt["L_trans"], pt["h_trans"], mf["L_rot"], mf["h_rot"])
    assert rel_err(L, g["L22"]) < 1e-8 and rel_err(h, g["h22"]) < 1e-7
    assert L.shape == (22, 22) and np.count_nonzero(L[6:, :]) == 0


FULL_CASES = sorted(os.path.basename(p) for p in glob.glob(os.path.join(GOLDEN, "binfull_*.npz")))


@pytest.mark.parametrize("case", FULL_CASES)
def test_bin_path_matches_reference_at_bench_size(case):
    """BASELINE config 3 / the bench shape (65,536 points, stride 1, epoch stamps): per-point arrays through digests,
    strided rows and column sums (conftest.check_compact), statistics and evidence in full."""
    from conftest import check_compact
    g = golden(case)
    pts, t, w, ring, tag = _raw(g)
    o = ob.lidar_evidence_bins(pts, t, w, ring, tag, int(g["cap"]), g["xi"], float(g["t0"]), float(g["t1"]), g["origin"],
                               g["bin_dirs"], float(g["tau"]),
                               {k[4:]: g[k] for k in g.files if k.startswith("map_") and k[4:] in
                                ("S_dir", "S_dir_scatter", "N_dir", "N_pos", "sum_p", "sum_ppT")},
                               lie.so3_exp(g["pose"][3:6]), g["pose"][:3])
    rs, dk = o["resample"], o["deskew"]
    for k_o, k_g in (("points", "rs_points"), ("timestamps", "rs_t"), ("ring", "rs_ring"), ("tag", "rs_tag")):
        check_compact(g, k_g, rs[k_o], exact=True)
    check_compact(g, "rs_w", rs["weights"], tol=1e-14)
    check_compact(g, "dk_points", dk["points"], tol=1e-13)
    check_compact(g, "dk_w", dk["weights"], tol=1e-13)
    st = o["stats"]
    for k_o, k_g in (("N", "st_N"), ("s_dir", "st_s_dir"), ("S_dir_scatter", "st_S"), ("p_bar", "st_p_bar"),
                     ("Sigma_p", "st_Sigma_p"), ("kappa_scan", "st_kappa")):
        assert rel_err(st[k_o], g[k_g]) < 1e-10, k_o
    assert geodesic(o["mf"]["R_mf"], g["mf_R"]) < 1e-9
    assert rel_err(o["L"], g["L22"]) < 1e-8 and rel_err(o["h"], g["h22"]) < 1e-7


def test_scalar_tables():
    g = golden("bin_scalars.npz")
    kb = ob.kappa_from_resultant_batch(g["R_bar"])
    assert rel_err(kb, g["kappa_batch"]) < 1e-13
    ks = np.array([ob.kappa_from_resultant_v2(float(r))[0]["kappa"] for r in g["R_bar"]])
    assert rel_err(ks, g["kappa_scalar"]) < 1e-13
    for v, R, lg in zip(g["rotvec"], g["so3_exp"], g["so3_log"]):
        assert np.max(np.abs(lie.so3_exp(v) - R)) < 1e-14
        assert np.max(np.abs(lie.so3_log(R) - lg)) < 1e-12
    for x, e in zip(g["xi6"], g["se3_exp"]):
        assert np.max(np.abs(lie.se3_exp(x) - e)) < 1e-13
