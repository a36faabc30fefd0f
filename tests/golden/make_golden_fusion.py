#!/usr/bin/env python
"""
Golden vectors for the evidence-fusion step (pipeline steps 9-11), produced by the REFERENCE's own functions on top of
oracle/jax_shim:
  compute_excitation_scales_jax, apply_excitation_prior_scaling_jax   fl/backend/operators/excitation.py:15-64
  fusion_scale_from_certificates, info_fusion_additive                fl/backend/operators/fusion.py:46-230
  aggregate_certificates, CertBundle                                  fl/common/certificates.py
The glue between them (raw evidence, sentinels, the beta law, the pose-block conditioning) is inline code of
process_scan_single_hypothesis (fl/backend/pipeline.py:1038-1193) and cannot be called; it is restated below with the
reference's own jnp expressions, statement by statement, as tests/golden/make_golden_prim.py does for step 12b.

Run in the build container only (needs /root/reference):   python tests/golden/make_golden_fusion.py
Outputs tests/golden/fusion_*.npz.  Nothing under /root/reference is written or copied.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path.insert(0, os.path.join(ROOT, "oracle", "jax_shim"))
sys.path.insert(0, os.path.join(REF, "fl_ws", "src", "fl_slam_poc"))
sys.path.insert(0, ROOT)


def main():
    import jax.numpy as jnp
    from fl_slam_poc.backend.operators.excitation import apply_excitation_prior_scaling_jax, compute_excitation_scales_jax
    from fl_slam_poc.backend.operators.fusion import fusion_scale_from_certificates, info_fusion_additive
    from fl_slam_poc.common import constants
    from fl_slam_poc.common.belief import BeliefGaussianInfo
    from fl_slam_poc.common.certificates import CertBundle, ConditioningCert

    from gc_slam_b200 import synth

    cfg = dict(power_beta_min=0.25, power_beta_z_c=1.0, power_beta_exc_c=50.0, alpha_min=1.0, alpha_max=1.0, c0_cond=1e6,
               eps_mass=1e-12, eps_psd=1e-12)          # PipelineConfig defaults (pipeline.py:104-125)

    def one(L_lidar, h_lidar, L_io, h_io, L_prior, h_prior, ess_total, dt_effect, extrinsic_effect, nll_per_ess, cfg):
        # ---- pipeline.py:1038-1039
        L_raw = jnp.asarray(L_io) + jnp.asarray(L_lidar)
        h_raw = jnp.asarray(h_io) + jnp.asarray(h_lidar)
        cert = CertBundle.create_exact(chart_id="GC-RIGHT-01", anchor_id="golden")
        cert.support.ess_total = float(ess_total)
        cert.excitation.dt_effect = float(dt_effect)
        cert.excitation.extrinsic_effect = float(extrinsic_effect)
        cert.mismatch.nll_per_ess = float(nll_per_ess)
        # ---- :1069-1090
        eps = float(cfg["eps_mass"])
        dt_pose = jnp.linalg.norm(L_raw[constants.GC_IDX_DT, constants.GC_IDX_POSE]) + jnp.linalg.norm(L_raw[constants.GC_IDX_POSE, constants.GC_IDX_DT])
        dt_vel = jnp.linalg.norm(L_raw[constants.GC_IDX_DT, constants.GC_IDX_VEL]) + jnp.linalg.norm(L_raw[constants.GC_IDX_VEL, constants.GC_IDX_DT])
        dt_asym = jnp.clip(jnp.abs(dt_vel - dt_pose) / (dt_vel + dt_pose + eps), 0.0, 1.0)
        z_to_xy = jnp.abs(L_raw[2, 2]) / (0.5 * (jnp.abs(L_raw[0, 0]) + jnp.abs(L_raw[1, 1])) + eps)
        cert.overconfidence.dt_asymmetry = float(dt_asym)
        cert.overconfidence.z_to_xy_ratio = float(z_to_xy)
        # ---- :1093-1107
        exc_total = cert.excitation.dt_effect + cert.excitation.extrinsic_effect
        ess_to_exc = float(cert.support.ess_total) / (float(exc_total) + float(cfg["eps_mass"]))
        s_z = float(z_to_xy) / (float(z_to_xy) + float(cfg["power_beta_z_c"]))
        s_exc = 1.0 / (1.0 + (ess_to_exc / float(cfg["power_beta_exc_c"])))
        s = jnp.clip(jnp.asarray(dt_asym * s_z * s_exc, dtype=jnp.float64), 0.0, 1.0)
        beta = float(cfg["power_beta_min"] + (1.0 - cfg["power_beta_min"]) * float(s))
        beta = float(jnp.clip(jnp.asarray(beta, dtype=jnp.float64), cfg["power_beta_min"], 1.0))
        L_ev = beta * L_raw
        h_ev = beta * h_raw
        cert.influence.power_beta = beta
        # ---- :1119-1126
        s_dt, s_ex = compute_excitation_scales_jax(L_evidence=L_ev, L_prior=jnp.asarray(L_prior))
        L_ps, h_ps = apply_excitation_prior_scaling_jax(L_prior=jnp.asarray(L_prior), h_prior=jnp.asarray(h_prior), s_dt=s_dt, s_ex=s_ex)
        # ---- :1155-1177
        eps_cond = float(cfg["eps_psd"])
        L_pose = 0.5 * (L_ev[constants.GC_IDX_POSE, constants.GC_IDX_POSE] + L_ev[constants.GC_IDX_POSE, constants.GC_IDX_POSE].T)
        L_pose = jnp.nan_to_num(L_pose, nan=0.0, posinf=0.0, neginf=0.0)
        ev = jnp.linalg.eigvalsh(L_pose)
        safe = jnp.nan_to_num(ev, nan=eps_cond, posinf=eps_cond, neginf=eps_cond)
        cl = jnp.maximum(safe, eps_cond)
        cert.conditioning = ConditioningCert(eig_min=float(cl[0]), eig_max=float(cl[-1]), cond=float(cl[-1] / cl[0]),
                                             near_null_count=int(jnp.sum(ev <= eps_cond)))
        # ---- :1181-1193
        fs, fs_cert, _ = fusion_scale_from_certificates(cert_evidence=cert, cert_belief=CertBundle.create_exact("GC-RIGHT-01", "golden"),
                                                        alpha_min=cfg["alpha_min"], alpha_max=cfg["alpha_max"], kappa_scale=1.0,
                                                        c0_cond=cfg["c0_cond"], chart_id="GC-RIGHT-01", anchor_id="golden")
        alpha = float(fs.alpha)
        # ---- :1198-1207
        belief = BeliefGaussianInfo(chart_id="GC-RIGHT-01", anchor_id="golden", X_anchor=jnp.zeros(6), stamp_sec=0.0,
                                    z_lin=jnp.zeros(22), L=L_ps, h=h_ps, cert=CertBundle.create_exact("GC-RIGHT-01", "golden"))
        post, f_cert, f_eff = info_fusion_additive(belief_pred=belief, L_evidence=L_ev, h_evidence=h_ev, alpha=alpha,
                                                   eps_psd=cfg["eps_psd"], chart_id="GC-RIGHT-01", anchor_id="golden")
        return dict(L_post=np.asarray(post.L), h_post=np.asarray(post.h), L_evidence=np.asarray(L_ev), h_evidence=np.asarray(h_ev),
                    L_prior_scaled=np.asarray(L_ps), h_prior_scaled=np.asarray(h_ps), beta=beta, dt_asymmetry=float(dt_asym),
                    z_to_xy_ratio=float(z_to_xy), s_dt=float(s_dt), s_ex=float(s_ex), pose_eig_min=cert.conditioning.eig_min,
                    pose_eig_max=cert.conditioning.eig_max, pose_cond=cert.conditioning.cond,
                    pose_near_null=cert.conditioning.near_null_count, alpha=alpha,
                    fs_ess_to_excitation=fs_cert.overconfidence.ess_to_excitation,
                    post_eig_min=f_cert.conditioning.eig_min, post_eig_max=f_cert.conditioning.eig_max,
                    post_cond=f_cert.conditioning.cond, post_near_null=f_cert.conditioning.near_null_count,
                    psd_projection_delta=f_cert.influence.psd_projection_delta, trace_increase=f_eff.predicted,
                    fusion_triggers=np.array(f_cert.approximation_triggers, dtype="U64"))

    cases = {"fusion_k4_defaults": (4, 91, cfg, False),
             "fusion_k8_alpha_range": (8, 92, dict(cfg, alpha_min=0.2, alpha_max=0.9, power_beta_min=0.1, power_beta_exc_c=5.0), False),
             "fusion_k3_indefinite_prior": (3, 93, dict(cfg, alpha_min=0.5, alpha_max=1.0), True)}
    for name, (K, seed, c, nasty) in cases.items():
        ins = synth.fusion_inputs(K, seed, indefinite=nasty)
        outs = [one(ins["L_lidar"][k], ins["h_lidar"][k], ins["L_other"][k], ins["h_other"][k], ins["L_prior"][k], ins["h_prior"][k],
                    ins["ess_total"][k], ins["dt_effect"][k], ins["extrinsic_effect"][k], ins["nll_per_ess"][k], c) for k in range(K)]
        d = {("out_" + key): np.stack([np.asarray(o[key]) for o in outs]) for key in outs[0]}
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **ins, **d, **{"cfg_" + k: v for k, v in c.items()})
        print(name, "beta", d["out_beta"], "alpha", d["out_alpha"], "post near-null", d["out_post_near_null"])


if __name__ == "__main__":
    main()
