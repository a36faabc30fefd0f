"""
Device-side hypothesis combine against the reference's own _hypothesis_barycenter_core outputs
(tests/golden/hyp_*.npz, made by tests/golden/make_golden_hyp.py) and against oracle/hypothesis.py on fresh inputs.
Through the C-ABI entry gcs_hypothesis_barycenter.  Tolerances: weights / barycenter 1e-13, PSD-projected matrix
1e-12 of its norm (Jacobi vs LAPACK eigenvectors), spread proxy 1e-6 (solves with condition numbers up to 1e9).
"""
import glob
import os

import numpy as np
import pytest

from conftest import GOLDEN, golden, rel_err

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")

HYP_CASES = sorted(os.path.basename(p) for p in glob.glob(os.path.join(GOLDEN, "hyp_*.npz")))


def _np(t):
    return t.detach().cpu().numpy()


@pytest.fixture(scope="module")
def sh():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    from gc_slam_b200 import sharding
    return sharding


def _check(res, cert, eff, o, K, cond=1e9):
    assert rel_err(_np(res.weights_normalized), o["weights_normalized"]) < 1e-14
    assert abs(res.floor_adjustment - float(o["floor_adjustment"])) < 1e-15
    assert rel_err(_np(res.L), o["L"]) < 1e-12 and rel_err(_np(res.h), o["h"]) < 1e-13
    assert rel_err(_np(res.z_lin), o["z_lin"]) < 1e-13
    L = _np(res.L)
    assert np.array_equal(L, L) and np.max(np.abs(L - L.T)) < 1e-10 * np.max(np.abs(L))
    pc = o["psd_cert"]
    assert abs(cert.conditioning.eig_max - pc[3]) < 1e-12 * pc[3]
    if pc[2] > 1e-12 * pc[3]:
        assert abs(cert.conditioning.eig_min - pc[2]) < 1e-12 * pc[3] + 1e-9 * pc[2]
        assert cert.conditioning.near_null_count == int(pc[5])
    else:
        # eigenvalues below float64 resolution of a matrix of this norm (|L| * 2e-16) are rounding noise in LAPACK and in
        # the Jacobi iteration alike: only their scale is comparable
        assert cert.conditioning.eig_min < 1e-12 * pc[3] and cert.conditioning.near_null_count <= 3
    assert abs(cert.influence.psd_projection_delta - pc[0]) < 1e-12 * np.max(np.abs(o["L"]))
    # the means solve (L_k + eps_lift I) mu = h: their accuracy -- in LAPACK as here -- is cond * 2e-16
    assert abs(eff.predicted - float(o["spread_proxy"])) <= min(0.5, 1e-6 + 10.0 * cond * 2.2e-16) * abs(float(o["spread_proxy"])) + 1e-20
    wn = o["weights_normalized"]
    assert abs(cert.support.ess_total - 1.0 / np.sum(wn ** 2)) < 1e-12 * K
    assert cert.approximation_triggers == ["HypothesisProjection", "I-projection-info-barycenter"] and not cert.exact


@pytest.mark.parametrize("case", HYP_CASES)
def test_combine_vs_reference_golden(sh, case):
    g = golden(case)
    K = g["weights"].shape[0]
    res, cert, eff = sh.hypothesis_barycenter_projection(g["L_stack"], g["h_stack"], g["weights"], g["z_lin_stack"], K_HYP=K,
                                                         HYP_WEIGHT_FLOOR=float(g["weight_floor"]), eps_psd=float(g["eps_psd"]),
                                                         eps_lift=float(g["eps_lift"]))
    cond = max(np.linalg.cond(0.5 * (Lk + Lk.T) + float(g["eps_lift"]) * np.eye(Lk.shape[0])) for Lk in g["L_stack"])
    _check(res, cert, eff, g, K, cond)
    res2, _, eff2 = sh.hypothesis_barycenter_projection(g["L_stack"], g["h_stack"], g["weights"], g["z_lin_stack"])
    assert torch.equal(res.L, res2.L) and eff.predicted == eff2.predicted       # bit-identical reruns


@pytest.mark.parametrize("K,D", [(1, 22), (2, 3), (7, 21), (64, 22), (300, 32), (3, 1)])
def test_combine_vs_oracle_shapes(sh, K, D):
    from gc_slam_b200 import synth
    from oracle import hypothesis as oh
    Ls, hs, zs, w = synth.hypothesis_evidence_stack(K, D, 200 + K + D)
    o = oh.hypothesis_barycenter(Ls, hs, zs, w)
    res, cert, eff = sh.hypothesis_barycenter_projection(Ls, hs, w, zs)
    _check(res, cert, eff, o, K)


def test_combine_of_a_device_resident_plan_and_errors(sh):
    """64 hypotheses of one scan evaluated by a BinPathPlan, combined without leaving the device; equals the host combine."""
    from gc_slam_b200 import operators as ops, synth
    H, n = 64, 8192
    bins = synth.fibonacci_atlas(48)
    pts, t, w, ring, tag = synth.vlp16_scan(n, 5, t0=synth.EPOCH_T0)
    plan = ops.BinPathPlan(1, n, n, n_hyp=H, n_bins=48, tau=0.1, origin=synth.lidar_origin_base(), want_evidence=True,
                           materialize_deskewed=False)
    plan.set_bins(bins, 0.1)
    plan.set_map(synth.random_map_bin_stats(48, 7, bins))
    plan.upload(pts[None], t[None], w[None], ring[None], tag[None], np.array([synth.EPOCH_T0]), np.array([synth.EPOCH_T0 + 0.1]),
                np.stack([synth.scan_twist(60 + k) for k in range(H)]), synth.hypothesis_poses(H, 3), non_blocking=False)
    plan.run()
    out = plan.outputs()
    wts = np.random.default_rng(0).dirichlet(np.ones(H))
    res, cert, _ = sh.hypothesis_barycenter_projection(out.L22, out.h22, wts)
    Lh, hh, wn, adj = sh.hypothesis_barycenter(_np(out.L22), _np(out.h22), wts)
    assert rel_err(_np(res.L), Lh) < 1e-12 and rel_err(_np(res.h), hh) < 1e-13 and res.z_lin is None
    assert cert.compute.device_runtime.host_sync_count_est == 1
    with pytest.raises(ValueError):
        sh.hypothesis_barycenter_projection(out.L22, out.h22, wts[:5])
    with pytest.raises(ValueError):
        sh.hypothesis_barycenter_projection(out.L22, out.h22, wts, K_HYP=4)
    with pytest.raises(ValueError):
        sh.hypothesis_barycenter_projection(np.zeros((2, 40, 40)), np.zeros((2, 40)), np.ones(2))
