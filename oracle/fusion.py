"""
Evidence fusion (pipeline steps 9-11) for one hypothesis, NumPy float64.  TEST INFRASTRUCTURE (see oracle/__init__.py).

Restates, line for line:
  raw evidence, observability sentinels, power tempering      fl/backend/pipeline.py:1038-1117
  compute_excitation_scales_jax, apply_excitation_prior_scaling_jax
                                                              fl/backend/operators/excitation.py:15-64 (pipeline.py:1119-1146)
  pose-block conditioning                                     fl/backend/pipeline.py:1155-1177
  fusion_scale_from_certificates (the alpha law)              fl/backend/operators/fusion.py:76-113
  info_fusion_additive                                        fl/backend/operators/fusion.py:186-225
  domain_projection_psd_core                                  fl/common/primitives.py:80-123 (oracle/hypothesis.py)
Pinned to the reference's own functions by tests/golden/make_golden_fusion.py (tests/golden/fusion_*.npz).
"""
from __future__ import annotations

import numpy as np

from .hypothesis import domain_projection_psd

IDX_POSE = slice(0, 6)
IDX_VEL = slice(6, 9)
IDX_DT = 15
IDX_EX = slice(16, 22)

DEFAULT_CFG = dict(power_beta_min=0.25, power_beta_z_c=1.0, power_beta_exc_c=50.0, alpha_min=1.0, alpha_max=1.0, c0_cond=1e6,
                   eps_mass=1e-12, eps_psd=1e-12, exc_eps=1e-12)


def sentinels(L_raw, eps):
    """pipeline.py:1070-1087."""
    dt_pose = np.linalg.norm(L_raw[IDX_DT, IDX_POSE]) + np.linalg.norm(L_raw[IDX_POSE, IDX_DT])
    dt_vel = np.linalg.norm(L_raw[IDX_DT, IDX_VEL]) + np.linalg.norm(L_raw[IDX_VEL, IDX_DT])
    dt_asym = float(np.clip(np.abs(dt_vel - dt_pose) / (dt_vel + dt_pose + eps), 0.0, 1.0))
    z_to_xy = float(np.abs(L_raw[2, 2]) / (0.5 * (np.abs(L_raw[0, 0]) + np.abs(L_raw[1, 1])) + eps))
    return dt_asym, z_to_xy


def power_beta(dt_asym, z_to_xy, ess_total, exc_total, cfg):
    """pipeline.py:1091-1102."""
    ess_to_exc = float(ess_total) / (float(exc_total) + float(cfg["eps_mass"]))
    s_z = float(z_to_xy) / (float(z_to_xy) + float(cfg["power_beta_z_c"]))
    s_exc = 1.0 / (1.0 + (ess_to_exc / float(cfg["power_beta_exc_c"])))
    s = float(np.clip(dt_asym * s_z * s_exc, 0.0, 1.0))
    beta = float(cfg["power_beta_min"] + (1.0 - cfg["power_beta_min"]) * s)
    return float(np.clip(beta, cfg["power_beta_min"], 1.0)), ess_to_exc


def excitation_scales(L_evidence, L_prior, eps):
    """excitation.py:15-32."""
    e_dt = L_evidence[IDX_DT, IDX_DT]
    e_ex = np.trace(L_evidence[IDX_EX, IDX_EX])
    pi_dt = L_prior[IDX_DT, IDX_DT]
    pi_ex = np.trace(L_prior[IDX_EX, IDX_EX])
    return e_dt / (e_dt + pi_dt + eps), e_ex / (e_ex + pi_ex + eps)


def apply_prior_scaling(L_prior, h_prior, s_dt, s_ex):
    """excitation.py:35-64."""
    Lp = np.array(L_prior, dtype=np.float64)
    hp = np.array(h_prior, dtype=np.float64)
    a_dt, a_ex = 1.0 - s_dt, 1.0 - s_ex
    Lp[IDX_DT, :] = a_dt * Lp[IDX_DT, :]
    Lp[:, IDX_DT] = a_dt * Lp[:, IDX_DT]
    hp[IDX_DT] = a_dt * hp[IDX_DT]
    Lp[IDX_EX, :] = a_ex * Lp[IDX_EX, :]
    Lp[:, IDX_EX] = a_ex * Lp[:, IDX_EX]
    hp[IDX_EX] = a_ex * hp[IDX_EX]
    return Lp, hp


def pose_conditioning(L_evidence, eps_cond):
    """pipeline.py:1155-1177 -> eig_min, eig_max, cond, near_null_count."""
    Lp = 0.5 * (L_evidence[IDX_POSE, IDX_POSE] + L_evidence[IDX_POSE, IDX_POSE].T)
    Lp = np.nan_to_num(Lp, nan=0.0, posinf=0.0, neginf=0.0)
    ev = np.linalg.eigvalsh(Lp)
    safe = np.nan_to_num(ev, nan=eps_cond, posinf=eps_cond, neginf=eps_cond)
    cl = np.maximum(safe, eps_cond)
    return float(cl[0]), float(cl[-1]), float(cl[-1] / cl[0]), int(np.sum(ev <= eps_cond))


def fusion_alpha(cond, ess, exc_total, dt_asym, z_to_xy, beta, nll_per_ess, cfg):
    """fusion.py:76-113."""
    cond_q = cfg["c0_cond"] / (cond + cfg["c0_cond"])
    supp_q = ess / (ess + 1.0)
    mis_q = np.exp(-np.float64(nll_per_ess))
    dt_q = np.clip(np.float64(dt_asym), 0.0, 1.0)
    z_q = np.clip(np.float64(z_to_xy) / (np.float64(z_to_xy) + 1.0), 0.0, 1.0)
    exc_q = np.clip(np.float64(exc_total) / (np.float64(exc_total) + 1.0), 0.0, 1.0)
    base = np.sqrt(cond_q * supp_q)
    quality = base * mis_q * dt_q * z_q * exc_q * np.clip(np.float64(beta), 0.0, 1.0)
    alpha_raw = cfg["alpha_min"] + (cfg["alpha_max"] - cfg["alpha_min"]) * quality
    return float(np.clip(alpha_raw, cfg["alpha_min"], cfg["alpha_max"])), float(quality)


def evidence_fusion(L_lidar, h_lidar, L_other, h_other, L_prior, h_prior, ess_total, exc_total, nll_per_ess, cfg=None):
    """Steps 9-11 for one hypothesis; returns a dict of arrays and scalars."""
    cfg = dict(DEFAULT_CFG, **(cfg or {}))
    L_raw = np.asarray(L_other, np.float64) + np.asarray(L_lidar, np.float64)
    h_raw = np.asarray(h_other, np.float64) + np.asarray(h_lidar, np.float64)
    dt_asym, z_to_xy = sentinels(L_raw, float(cfg["eps_mass"]))
    beta, ess_to_exc = power_beta(dt_asym, z_to_xy, ess_total, exc_total, cfg)
    L_ev, h_ev = beta * L_raw, beta * h_raw
    s_dt, s_ex = excitation_scales(L_ev, np.asarray(L_prior, np.float64), cfg["exc_eps"])
    L_ps, h_ps = apply_prior_scaling(L_prior, h_prior, s_dt, s_ex)
    emin, emax, cond, nn = pose_conditioning(L_ev, float(cfg["eps_psd"]))
    alpha, quality = fusion_alpha(cond, float(ess_total), float(exc_total), dt_asym, z_to_xy, beta, nll_per_ess, cfg)
    L_post, pc = domain_projection_psd(L_ps + alpha * L_ev, cfg["eps_psd"])
    h_post = h_ps + alpha * h_ev
    return dict(L_post=L_post, h_post=h_post, L_evidence=L_ev, h_evidence=h_ev, L_prior_scaled=L_ps, h_prior_scaled=h_ps,
                beta=beta, dt_asymmetry=dt_asym, z_to_xy_ratio=z_to_xy, ess_to_excitation=ess_to_exc, s_dt=float(s_dt),
                s_ex=float(s_ex), pose_eig_min=emin, pose_eig_max=emax, pose_cond=cond, pose_near_null=nn, alpha=alpha,
                quality=quality, psd_cert=pc, trace_increase=float(np.trace(L_post) - np.trace(L_ps)))
