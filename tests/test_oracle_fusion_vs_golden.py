"""oracle/fusion.py (pipeline steps 9-11: tempering, excitation scaling, pose-block conditioning, fusion scale, additive
fusion + PSD projection) against the vectors produced by the reference's own functions
(tests/golden/make_golden_fusion.py).  CPU only."""
import glob
import os

import numpy as np
import pytest

from conftest import GOLDEN, golden, rel_err

FUSION_CASES = sorted(os.path.basename(p) for p in glob.glob(os.path.join(GOLDEN, "fusion_*.npz")))
CFG_KEYS = ("power_beta_min", "power_beta_z_c", "power_beta_exc_c", "alpha_min", "alpha_max", "c0_cond", "eps_mass", "eps_psd")
SCALARS = (("beta", 1e-14), ("dt_asymmetry", 1e-13), ("z_to_xy_ratio", 1e-14), ("s_dt", 1e-13), ("s_ex", 1e-13), ("alpha", 1e-13))
ARRAYS = (("L_evidence", 1e-14), ("h_evidence", 1e-14), ("L_prior_scaled", 1e-14), ("h_prior_scaled", 1e-14), ("h_post", 1e-13))


def cfg_of(g):
    return {k: float(g["cfg_" + k]) for k in CFG_KEYS}


def check_hypothesis(o, g, k, tol_L=1e-10, tol_eig=1e-7):
    """o: dict in the layout of oracle.fusion.evidence_fusion for hypothesis k of golden g."""
    for name, tol in SCALARS:
        a, b = float(o[name]), float(g["out_" + name][k])
        assert abs(a - b) <= tol * max(1.0, abs(b)), (name, a, b)
    for name, tol in ARRAYS:
        assert rel_err(np.asarray(o[name]), g["out_" + name][k]) < tol, name
    # pose-block conditioning: extreme eigenvalues of a 6x6 block spread over ~8 decades
    assert abs(o["pose_eig_max"] - g["out_pose_eig_max"][k]) <= 1e-12 * g["out_pose_eig_max"][k]
    assert abs(o["pose_eig_min"] - g["out_pose_eig_min"][k]) <= tol_eig * g["out_pose_eig_min"][k] + 1e-16 * g["out_pose_eig_max"][k]
    assert int(o["pose_near_null"]) == int(g["out_pose_near_null"][k])
    # posterior: entries to 1e-10 of the matrix norm (eigenvectors of clustered eigenvalues differ, the rebuilt matrix does not)
    assert rel_err(np.asarray(o["L_post"]), g["out_L_post"][k]) < tol_L
    pc = o["psd_cert"]
    scale = float(g["out_post_eig_max"][k])
    assert abs(pc[0] - g["out_psd_projection_delta"][k]) <= 1e-9 * scale
    assert abs(pc[3] - scale) <= 1e-12 * scale
    assert abs(pc[2] - g["out_post_eig_min"][k]) <= 1e-6 * g["out_post_eig_min"][k] + 1e-16 * scale
    assert abs(o["trace_increase"] - g["out_trace_increase"][k]) <= 1e-9 * scale
    ref_nn = int(g["out_post_near_null"][k])
    assert int(pc[5]) == ref_nn or abs(int(pc[5]) - ref_nn) <= _borderline(g["out_L_post"][k], float(g["cfg_eps_psd"]))


def _borderline(L_post, eps_psd):
    """number of eigenvalues of the reference posterior within float64 resolution of the near-null threshold"""
    ev = np.linalg.eigvalsh(0.5 * (L_post + L_post.T))
    return int(np.sum(np.abs(ev - 10.0 * eps_psd) <= 4e-16 * np.max(np.abs(ev))))


def test_cases_present():
    assert len(FUSION_CASES) >= 3


@pytest.mark.parametrize("case", FUSION_CASES)
def test_oracle_fusion_matches_reference(case):
    from oracle import fusion as of
    g = golden(case)
    K = g["L_lidar"].shape[0]
    for k in range(K):
        o = of.evidence_fusion(g["L_lidar"][k], g["h_lidar"][k], g["L_other"][k], g["h_other"][k], g["L_prior"][k], g["h_prior"][k],
                               g["ess_total"][k], g["dt_effect"][k] + g["extrinsic_effect"][k], g["nll_per_ess"][k], cfg_of(g))
        check_hypothesis(o, g, k)
        assert abs(o["ess_to_excitation"] - g["out_fs_ess_to_excitation"][k]) <= 1e-13 * g["out_fs_ess_to_excitation"][k]
    assert [str(x) for x in g["out_fusion_triggers"][0]] == ["InfoFusionAdditive"]
