"""
Evidence fusion on the device (fusion half: pipeline steps 9-11 for all hypotheses in one launch) against
the reference's own outputs (tests/golden/fusion_*.npz, made by tests/golden/make_golden_fusion.py) and against
oracle/fusion.py on 64 hypotheses.  Through the C-ABI entry gcs_evidence_fusion.  Control-law scalars 1e-13, tempered
evidence / scaled prior 1e-14, posterior information 1e-10 of its norm (Jacobi vs LAPACK eigenvectors).
"""
import numpy as np
import pytest

from conftest import golden, rel_err
from test_oracle_fusion_vs_golden import FUSION_CASES, cfg_of, check_hypothesis

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")


@pytest.fixture(scope="module")
def F():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    from gc_slam_b200 import fusion
    return fusion


def _as_oracle_layout(res, k):
    r = res.rec[k]
    FU = __import__("gc_slam_b200.fusion", fromlist=["FU"]).FU
    return dict(L_post=res.L_post[k].cpu().numpy(), h_post=res.h_post[k].cpu().numpy(), L_evidence=res.L_evidence[k].cpu().numpy(),
                h_evidence=res.h_evidence[k].cpu().numpy(), L_prior_scaled=res.L_prior_scaled[k].cpu().numpy(),
                h_prior_scaled=res.h_prior_scaled[k].cpu().numpy(), beta=r[FU["BETA"]], dt_asymmetry=r[FU["DT_ASYMMETRY"]],
                z_to_xy_ratio=r[FU["Z_TO_XY"]], s_dt=r[FU["S_DT"]], s_ex=r[FU["S_EX"]], alpha=r[FU["ALPHA"]],
                pose_eig_min=r[FU["POSE_EIG_MIN"]], pose_eig_max=r[FU["POSE_EIG_MAX"]], pose_near_null=r[FU["POSE_NEAR_NULL"]],
                psd_cert=np.array([r[FU["PSD_PROJECTION_DELTA"]], r[FU["PSD_SYM_DELTA"]], r[FU["POST_EIG_MIN"]], r[FU["POST_EIG_MAX"]],
                                   r[FU["POST_COND"]], r[FU["POST_NEAR_NULL"]]]),
                trace_increase=r[FU["TRACE_INCREASE"]], ess_to_excitation=r[FU["ESS_TO_EXC"]])


@pytest.mark.parametrize("case", FUSION_CASES)
def test_fusion_vs_reference_golden(F, case):
    g = golden(case)
    K = g["L_lidar"].shape[0]
    res, certs, effects = F.evidence_fusion_batched(g["L_lidar"], g["h_lidar"], g["L_other"], g["h_other"], g["L_prior"], g["h_prior"],
                                                    g["ess_total"], g["dt_effect"] + g["extrinsic_effect"], g["nll_per_ess"],
                                                    config=F.FusionConfig(**cfg_of(g)))
    assert len(certs) == K and len(effects) == K
    for k in range(K):
        check_hypothesis(_as_oracle_layout(res, k), g, k)
        temper, exc_c, scale_c, fuse_c = certs[k]
        assert temper.approximation_triggers == ["PowerTempering"] and abs(temper.influence.power_beta - g["out_beta"][k]) < 1e-13
        assert temper.frobenius_applied == (abs(1.0 - float(g["out_beta"][k])) > 0.0)
        assert exc_c.approximation_triggers == ["ExcitationPriorScaling"]
        assert abs(exc_c.influence.dt_scale - (1.0 - g["out_s_dt"][k])) < 1e-13
        assert scale_c.exact and abs(scale_c.overconfidence.ess_to_excitation - g["out_fs_ess_to_excitation"][k]) <= 1e-13 * g["out_fs_ess_to_excitation"][k]
        assert fuse_c.approximation_triggers == [str(x) for x in g["out_fusion_triggers"][k]]
        assert isinstance(fuse_c.conditioning.near_null_count, int) and isinstance(fuse_c.influence.trust_alpha, float)


def test_single_operators_vs_reference_golden(F):
    """compute_excitation_scales / apply_excitation_prior_scaling / fusion_scale_from_certificates / info_fusion_additive one at
    a time, chained as the pipeline chains them, reproduce the golden of the batched path."""
    from gc_slam_b200.certs import CertBundle, ConditioningCert
    g = golden("fusion_k8_alpha_range.npz")
    cfg = cfg_of(g)
    for k in (0, 3):
        L_ev, h_ev = g["out_L_evidence"][k], g["out_h_evidence"][k]
        s_dt, s_ex = F.compute_excitation_scales(L_ev, g["L_prior"][k])
        assert abs(s_dt - g["out_s_dt"][k]) < 1e-13 and abs(s_ex - g["out_s_ex"][k]) < 1e-13
        L_ps, h_ps = F.apply_excitation_prior_scaling(g["L_prior"][k], g["h_prior"][k], s_dt, s_ex)
        assert rel_err(L_ps.cpu().numpy(), g["out_L_prior_scaled"][k]) < 1e-14 and rel_err(h_ps.cpu().numpy(), g["out_h_prior_scaled"][k]) < 1e-14
        ce = CertBundle.create_exact("GC-RIGHT-01", "t")
        ce.support.ess_total = float(g["ess_total"][k])
        ce.excitation.dt_effect, ce.excitation.extrinsic_effect = float(g["dt_effect"][k]), float(g["extrinsic_effect"][k])
        ce.mismatch.nll_per_ess = float(g["nll_per_ess"][k])
        ce.overconfidence.dt_asymmetry, ce.overconfidence.z_to_xy_ratio = float(g["out_dt_asymmetry"][k]), float(g["out_z_to_xy_ratio"][k])
        ce.influence.power_beta = float(g["out_beta"][k])
        ce.conditioning = ConditioningCert(eig_min=float(g["out_pose_eig_min"][k]), eig_max=float(g["out_pose_eig_max"][k]),
                                           cond=float(g["out_pose_cond"][k]), near_null_count=int(g["out_pose_near_null"][k]))
        fs, fs_cert, fs_eff = F.fusion_scale_from_certificates(ce, CertBundle.create_exact("GC-RIGHT-01", "t"), alpha_min=cfg["alpha_min"],
                                                               alpha_max=cfg["alpha_max"], c0_cond=cfg["c0_cond"])
        assert abs(fs.alpha - g["out_alpha"][k]) < 1e-13 and fs_cert.exact and fs_eff.predicted == fs.alpha

        class _B:
            L, h, X_anchor, stamp_sec, z_lin = L_ps, h_ps, None, 0.0, None
        post, cert, eff = F.info_fusion_additive(_B(), L_ev, h_ev, fs.alpha, eps_psd=cfg["eps_psd"])
        assert rel_err(post.L.cpu().numpy(), g["out_L_post"][k]) < 1e-10 and rel_err(post.h.cpu().numpy(), g["out_h_post"][k]) < 1e-13
        scale = float(g["out_post_eig_max"][k])
        assert abs(cert.influence.psd_projection_delta - g["out_psd_projection_delta"][k]) <= 1e-9 * scale
        assert abs(eff.predicted - g["out_trace_increase"][k]) <= 1e-9 * scale and cert.approximation_triggers == ["InfoFusionAdditive"]


def test_fusion_64_hypotheses_vs_oracle_feeds_the_combine(F):
    """BASELINE.json config 4 (64 hypotheses): one launch, against the oracle per hypothesis; the posterior stack goes into the
    hypothesis combine without leaving the device; rerun is bit-identical."""
    from gc_slam_b200 import sharding, synth
    from oracle import fusion as of
    K = 64
    ins = synth.fusion_inputs(K, 101)
    cfg = dict(of.DEFAULT_CFG, alpha_min=0.3, alpha_max=1.0)
    exc = ins["dt_effect"] + ins["extrinsic_effect"]
    fc = F.FusionConfig(**{k: v for k, v in cfg.items()})
    res, certs, _ = F.evidence_fusion_batched(ins["L_lidar"], ins["h_lidar"], ins["L_other"], ins["h_other"], ins["L_prior"], ins["h_prior"],
                                              ins["ess_total"], exc, ins["nll_per_ess"], config=fc)
    for k in range(K):
        o = of.evidence_fusion(ins["L_lidar"][k], ins["h_lidar"][k], ins["L_other"][k], ins["h_other"][k], ins["L_prior"][k],
                               ins["h_prior"][k], ins["ess_total"][k], exc[k], ins["nll_per_ess"][k], cfg)
        d = _as_oracle_layout(res, k)
        for name in ("beta", "dt_asymmetry", "z_to_xy_ratio", "s_dt", "s_ex", "alpha"):
            assert abs(d[name] - o[name]) <= 1e-13 * max(1.0, abs(o[name])), (k, name)
        assert rel_err(d["L_post"], o["L_post"]) < 1e-10 and rel_err(d["h_post"], o["h_post"]) < 1e-13
        assert rel_err(d["L_evidence"], o["L_evidence"]) < 1e-14 and rel_err(d["L_prior_scaled"], o["L_prior_scaled"]) < 1e-14
    res2, _, _ = F.evidence_fusion_batched(ins["L_lidar"], ins["h_lidar"], ins["L_other"], ins["h_other"], ins["L_prior"], ins["h_prior"],
                                           ins["ess_total"], exc, ins["nll_per_ess"], config=fc)
    assert torch.equal(res.L_post, res2.L_post) and torch.equal(res.h_post, res2.h_post) and np.array_equal(res.rec, res2.rec)
    w = np.full(K, 1.0 / K)
    comb, cert, _ = sharding.hypothesis_barycenter_projection(res.L_post, res.h_post, w)
    assert comb.L.is_cuda and np.isfinite(comb.L.cpu().numpy()).all() and cert.conditioning.eig_min > 0.0
