"""
Evidence fusion behind the per-hypothesis LiDAR evidence (fusion half): pipeline steps 9-11
of the reference for all K hypotheses of a scan in ONE launch (C entry gcs_evidence_fusion, csrc/gcs_fusion.cu), plus the
reference's operators of those steps one at a time:

  evidence_fusion_batched            fl/backend/pipeline.py:1038-1207 (raw evidence, sentinels, power tempering, excitation
                                     scaling of the prior, pose-block conditioning, fusion scale, additive fusion)
  compute_excitation_scales, apply_excitation_prior_scaling
                                     fl/backend/operators/excitation.py:15-64
  fusion_scale_from_certificates     fl/backend/operators/fusion.py:46-143   (host scalars in, host scalar out -- as there)
  info_fusion_additive               fl/backend/operators/fusion.py:150-230

Stacks are torch CUDA float64 tensors (NumPy arrays are uploaded); the arithmetic runs in libgcs_b200.so (the stand-alone
apply_excitation_prior_scaling is plain row / column scaling by torch indexing; the batched entry does it in the kernel).
"""
from __future__ import annotations

import ctypes as C
import math
from dataclasses import dataclass
from typing import Optional, Tuple

import numpy as np
import torch

from . import _lib as L
from . import constants
from .certs import CertBundle, ConditioningCert, ExpectedEffect, InfluenceCert, OverconfidenceCert
from .operators import _IO

_vp, _dbl, _int = C.c_void_p, C.c_double, C.c_int

SKIP_TEMPERING, SKIP_PRIOR_SCALING, ALPHA_GIVEN = 1, 2, 4
FU = dict(BETA=0, DT_ASYMMETRY=1, Z_TO_XY=2, ESS_TO_EXC=3, S_DT=4, S_EX=5, POSE_EIG_MIN=6, POSE_EIG_MAX=7, POSE_COND=8,
          POSE_NEAR_NULL=9, ALPHA=10, QUALITY=11, PSD_PROJECTION_DELTA=12, PSD_SYM_DELTA=13, POST_EIG_MIN=14, POST_EIG_MAX=15,
          POST_COND=16, POST_NEAR_NULL=17, TRACE_INCREASE=18, NREC=24)


class CFusionCfg(C.Structure):
    _fields_ = [(n, _dbl) for n in ("power_beta_min", "power_beta_z_c", "power_beta_exc_c", "alpha_min", "alpha_max", "c0_cond",
                                    "eps_mass", "eps_psd", "exc_eps", "alpha_override")] + [("flags", C.c_int32), ("reserved", C.c_int32)]


L.register_prototypes({
    "gcs_evidence_fusion": (_int, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _int, _int, C.POINTER(CFusionCfg), _vp, _vp, _vp, _vp,
                                   _vp, _vp, _vp]),
})


@dataclass
class FusionConfig:
    """The PipelineConfig fields steps 9-11 read (fl/backend/pipeline.py:104-125)."""
    power_beta_min: float = constants.GC_POWER_BETA_MIN
    power_beta_z_c: float = constants.GC_POWER_BETA_Z_C
    power_beta_exc_c: float = constants.GC_POWER_BETA_EXC_C
    alpha_min: float = constants.GC_ALPHA_MIN
    alpha_max: float = constants.GC_ALPHA_MAX
    kappa_scale: float = constants.GC_KAPPA_SCALE
    c0_cond: float = constants.GC_C0_COND
    eps_mass: float = constants.GC_EPS_MASS
    eps_psd: float = constants.GC_EPS_PSD
    exc_eps: float = constants.GC_EXC_EPS

    def _c(self, flags=0, alpha_override=0.0) -> CFusionCfg:
        return CFusionCfg(self.power_beta_min, self.power_beta_z_c, self.power_beta_exc_c, self.alpha_min, self.alpha_max,
                          self.c0_cond, self.eps_mass, self.eps_psd, self.exc_eps, float(alpha_override), int(flags), 0)


@dataclass
class EvidenceFusionResult:
    """Device stacks (K, 22, 22) / (K, 22) and the per-hypothesis record (K, NREC) on the host (index names: fusion.FU)."""
    L_post: torch.Tensor
    h_post: torch.Tensor
    L_evidence: torch.Tensor
    h_evidence: torch.Tensor
    L_prior_scaled: torch.Tensor
    h_prior_scaled: torch.Tensor
    rec: np.ndarray

    def column(self, name: str) -> np.ndarray:
        return self.rec[:, FU[name]]


def _launch(io, L_lidar, h_lidar, L_other, h_other, L_prior, h_prior, scal, cfg_c, want_aux=True):
    Ll = io.dev_in(L_lidar)
    if Ll.dim() == 2:
        Ll = Ll.unsqueeze(0)
    K, D = int(Ll.shape[0]), int(Ll.shape[-1])
    if tuple(Ll.shape) != (K, constants.GC_D_Z, constants.GC_D_Z):
        raise ValueError(f"evidence stacks must be (K, 22, 22), got {tuple(Ll.shape)}")
    def vec(x):
        return io.dev_in(x).reshape(K, D).contiguous()
    def mat(x):
        return io.dev_in(x).reshape(K, D, D).contiguous()
    Ll, hl, Lp, hp = mat(Ll), vec(h_lidar), mat(L_prior), vec(h_prior)
    Lo = mat(L_other) if L_other is not None else None
    ho = vec(h_other) if h_other is not None else None
    sc = io.dev_in(scal).reshape(K, 4).contiguous() if scal is not None else None
    L_post, h_post, rec = io.empty(K, D, D), io.empty(K, D), io.zeros(K, FU["NREC"])
    aux = (io.empty(K, D, D), io.empty(K, D), io.empty(K, D, D), io.empty(K, D)) if want_aux else (None, None, None, None)
    opt = lambda t: L.ptr(t) if t is not None else None
    io.ctx.check(io.ctx.lib.gcs_evidence_fusion(io.ctx.handle, io.stream(), L.ptr(Ll), L.ptr(hl), opt(Lo), opt(ho), L.ptr(Lp), L.ptr(hp),
                                                opt(sc), K, D, C.byref(cfg_c), L.ptr(L_post), L.ptr(h_post), opt(aux[0]), opt(aux[1]),
                                                opt(aux[2]), opt(aux[3]), L.ptr(rec)))
    return L_post, h_post, aux, rec


def evidence_fusion_batched(L_lidar, h_lidar, L_imu_odom, h_imu_odom, L_prior, h_prior, ess_total, excitation_total,
                            nll_per_ess=None, config: Optional[FusionConfig] = None, chart_id: str = constants.GC_CHART_ID,
                            anchor_id: str = "evidence_fusion", support_frac=1.0
                            ) -> Tuple[EvidenceFusionResult, "_PerHypothesis", "_PerHypothesis"]:
    """
    Steps 9-11 of process_scan_single_hypothesis (fl/backend/pipeline.py:1038-1207) for K hypotheses at once: the stacks may
    come straight from a K-hypothesis BinPathPlan (L22, h22) and go straight into hypothesis_barycenter_projection; one
    launch, one read-back of the (K, 24) record.  ess_total / excitation_total / nll_per_ess: per-hypothesis scalars of the
    aggregated evidence certificate (support.ess_total, excitation.dt_effect + extrinsic_effect, mismatch.nll_per_ess;
    support_frac = its support.support_frac, which only enters the fusion-scale certificate's cond_to_support).
    Returns the result, and per hypothesis the certificates the reference appends at these steps
    ([PowerTempering, ExcitationPriorScaling, fusion scale (exact), InfoFusionAdditive]) and the fusion's ExpectedEffect --
    as sequences that build entry k when it is read (len(), indexing, iteration).
    """
    cfg = config or FusionConfig()
    io = _IO()
    ess = np.asarray(ess_total, dtype=np.float64).reshape(-1)
    K = ess.shape[0]
    exc = np.broadcast_to(np.asarray(excitation_total, dtype=np.float64).reshape(-1), (K,))
    nll = np.zeros(K) if nll_per_ess is None else np.broadcast_to(np.asarray(nll_per_ess, dtype=np.float64).reshape(-1), (K,))
    sup = np.broadcast_to(np.asarray(support_frac, dtype=np.float64).reshape(-1), (K,))
    scal = np.stack([ess, exc, nll, np.zeros(K)], axis=1)
    L_post, h_post, aux, rec_d = _launch(io, L_lidar, h_lidar, L_imu_odom, h_imu_odom, L_prior, h_prior, scal, cfg._c())
    rec = io.host(rec_d)
    certs = _PerHypothesis(K, lambda k: _step_certs(rec[k], float(exc[k]), chart_id, anchor_id, io if k == 0 else None, float(sup[k])))
    effects = _PerHypothesis(K, lambda k: ExpectedEffect("predicted_info_trace_increase", float(rec[k][FU["TRACE_INCREASE"]]), None))
    return EvidenceFusionResult(L_post, h_post, aux[0], aux[1], aux[2], aux[3], rec), certs, effects


class _PerHypothesis:
    """Sequence of per-hypothesis objects assembled on access: 64 hypotheses x 4 certificates are ~2 ms of Python object
    construction, more than the launch and the read-back together, and most callers read a few of them."""

    def __init__(self, n, make):
        self._n, self._make, self._cache = int(n), make, {}

    def __len__(self):
        return self._n

    def __getitem__(self, k):
        if isinstance(k, slice):
            return [self[i] for i in range(*k.indices(self._n))]
        k = int(k)
        if k < 0:
            k += self._n
        if not 0 <= k < self._n:
            raise IndexError(k)
        if k not in self._cache:
            self._cache[k] = self._make(k)
        return self._cache[k]

    def __iter__(self):
        return (self[k] for k in range(self._n))


def _step_certs(r, exc_total, chart_id, anchor_id, io=None, support_frac=1.0):
    """[PowerTempering, ExcitationPriorScaling, fusion scale, InfoFusionAdditive] of one hypothesis (pipeline.py:1108-1207)."""
    beta, alpha = float(r[FU["BETA"]]), float(r[FU["ALPHA"]])
    temper = CertBundle.create_approx(chart_id=chart_id, anchor_id=anchor_id, triggers=["PowerTempering"],
                                      frobenius_applied=abs(1.0 - beta) > 0.0,
                                      influence=InfluenceCert.identity().with_overrides(power_beta=beta))
    exc_c = CertBundle.create_approx(chart_id=chart_id, anchor_id=anchor_id, triggers=["ExcitationPriorScaling"],
                                     influence=InfluenceCert.identity().with_overrides(dt_scale=float(1.0 - r[FU["S_DT"]]),
                                                                                       extrinsic_scale=float(1.0 - r[FU["S_EX"]])))
    scale_c = CertBundle.create_exact(chart_id=chart_id, anchor_id=anchor_id,
                                      overconfidence=OverconfidenceCert(excitation_total=exc_total,
                                                                        ess_to_excitation=float(r[FU["ESS_TO_EXC"]]),
                                                                        cond_to_support=float(r[FU["POSE_COND"]]) / (float(support_frac) + constants.GC_EPS_MASS),
                                                                        dt_asymmetry=float(r[FU["DT_ASYMMETRY"]]),
                                                                        z_to_xy_ratio=float(r[FU["Z_TO_XY"]])),
                                      conditioning=ConditioningCert(eig_min=float(r[FU["POSE_EIG_MIN"]]), eig_max=float(r[FU["POSE_EIG_MAX"]]),
                                                                    cond=float(r[FU["POSE_COND"]]), near_null_count=int(r[FU["POSE_NEAR_NULL"]])),
                                      influence=InfluenceCert.identity().with_overrides(trust_alpha=alpha))
    return [temper, exc_c, scale_c, _fusion_cert(r, alpha, chart_id, anchor_id, io)]


def _fusion_cert(r, alpha, chart_id, anchor_id, io=None):
    kw = dict(compute=io.compute(psd_projection_count=1)) if io is not None else {}
    return CertBundle.create_approx(chart_id=chart_id, anchor_id=anchor_id, triggers=["InfoFusionAdditive"],
                                    conditioning=ConditioningCert(eig_min=float(r[FU["POST_EIG_MIN"]]), eig_max=float(r[FU["POST_EIG_MAX"]]),
                                                                  cond=float(r[FU["POST_COND"]]), near_null_count=int(r[FU["POST_NEAR_NULL"]])),
                                    influence=InfluenceCert.identity().with_overrides(psd_projection_delta=float(r[FU["PSD_PROJECTION_DELTA"]]),
                                                                                      trust_alpha=alpha), **kw)


# ---- the operators of these steps one at a time ------------------------------------------------------------------
def compute_excitation_scales(L_evidence, L_prior, eps: float = constants.GC_EXC_EPS):
    """compute_excitation_scales_jax (excitation.py:15-32) -> (s_dt, s_ex) as Python floats."""
    io = _IO()
    cfg = FusionConfig(exc_eps=eps)._c(flags=SKIP_TEMPERING | ALPHA_GIVEN, alpha_override=0.0)
    D = constants.GC_D_Z
    _, _, _, rec = _launch(io, L_evidence, torch.zeros(D, dtype=torch.float64), None, None, L_prior,
                           torch.zeros(D, dtype=torch.float64), None, cfg, want_aux=False)
    r = io.host(rec)[0]
    return float(r[FU["S_DT"]]), float(r[FU["S_EX"]])


def apply_excitation_prior_scaling(L_prior, h_prior, s_dt, s_ex):
    """apply_excitation_prior_scaling_jax (excitation.py:35-64) -> (L_prior_scaled, h_prior_scaled) on the device.
    Pure row / column scaling by 1 - s on the dt row and the extrinsic block."""
    io = _IO()
    Lp, hp = io.dev_in(L_prior).clone(), io.dev_in(h_prior).clone()
    a_dt, a_ex = 1.0 - float(s_dt), 1.0 - float(s_ex)
    dt, ex = 15, slice(16, 22)
    Lp[dt, :] = a_dt * Lp[dt, :]; Lp[:, dt] = a_dt * Lp[:, dt]; hp[dt] = a_dt * hp[dt]
    Lp[ex, :] = a_ex * Lp[ex, :]; Lp[:, ex] = a_ex * Lp[:, ex]; hp[ex] = a_ex * hp[ex]
    return Lp, hp


@dataclass
class FusionScaleResult:
    alpha: float


def fusion_scale_from_certificates(cert_evidence: CertBundle, cert_belief: CertBundle, alpha_min: float = constants.GC_ALPHA_MIN,
                                   alpha_max: float = constants.GC_ALPHA_MAX, kappa_scale: float = constants.GC_KAPPA_SCALE,
                                   c0_cond: float = constants.GC_C0_COND, chart_id: str = constants.GC_CHART_ID,
                                   anchor_id: str = "initial") -> Tuple[FusionScaleResult, CertBundle, ExpectedEffect]:
    """fusion_scale_from_certificates (fusion.py:46-143): a closed-form law on certificate scalars (host floats in the
    reference as well); the batched entry evaluates the same law on the device for all hypotheses."""
    del cert_belief, kappa_scale
    ce = cert_evidence
    cond, ess, support_frac = ce.conditioning.cond, ce.support.ess_total, ce.support.support_frac
    exc = ce.excitation.dt_effect + ce.excitation.extrinsic_effect
    dt_asym, z = ce.overconfidence.dt_asymmetry, ce.overconfidence.z_to_xy_ratio
    clip01 = lambda x: min(max(x, 0.0), 1.0)
    quality = (math.sqrt((c0_cond / (cond + c0_cond)) * (ess / (ess + 1.0))) * math.exp(-ce.mismatch.nll_per_ess) * clip01(dt_asym)
               * clip01(z / (z + 1.0)) * clip01(exc / (exc + 1.0)) * clip01(ce.influence.power_beta))
    alpha = float(min(max(alpha_min + (alpha_max - alpha_min) * quality, alpha_min), alpha_max))
    cert = CertBundle.create_exact(chart_id=chart_id, anchor_id=anchor_id,
                                   overconfidence=OverconfidenceCert(excitation_total=float(exc),
                                                                     ess_to_excitation=float(ess) / (float(exc) + constants.GC_EPS_MASS),
                                                                     cond_to_support=float(cond) / (float(support_frac) + constants.GC_EPS_MASS),
                                                                     dt_asymmetry=float(dt_asym), z_to_xy_ratio=float(z)),
                                   influence=InfluenceCert.identity().with_overrides(trust_alpha=alpha))
    return FusionScaleResult(alpha=alpha), cert, ExpectedEffect("fusion_alpha", alpha, None)


@dataclass
class FusedBelief:
    """The fields of BeliefGaussianInfo that info_fusion_additive changes or passes through (fl/common/belief.py)."""
    chart_id: str
    anchor_id: str
    X_anchor: object
    stamp_sec: float
    z_lin: object
    L: torch.Tensor
    h: torch.Tensor
    cert: CertBundle


def info_fusion_additive(belief_pred, L_evidence, h_evidence, alpha: float, eps_psd: float = constants.GC_EPS_PSD,
                         chart_id: str = constants.GC_CHART_ID, anchor_id: str = "initial"
                         ) -> Tuple[FusedBelief, CertBundle, ExpectedEffect]:
    """info_fusion_additive (fusion.py:150-230): L_post = DomainProjectionPSD(L_pred + alpha L_evidence), h_post likewise;
    belief_pred is any object with attributes L, h (and X_anchor, stamp_sec, z_lin, passed through)."""
    io = _IO()
    alpha = float(alpha)
    cfg = FusionConfig(eps_psd=eps_psd)._c(flags=SKIP_TEMPERING | SKIP_PRIOR_SCALING | ALPHA_GIVEN, alpha_override=alpha)
    L_post, h_post, _, rec = _launch(io, L_evidence, h_evidence, None, None, belief_pred.L, belief_pred.h, None, cfg, want_aux=False)
    r = io.host(rec)[0]
    cert = _fusion_cert(r, alpha, chart_id, anchor_id, io)
    post = FusedBelief(chart_id=chart_id, anchor_id=anchor_id, X_anchor=getattr(belief_pred, "X_anchor", None),
                       stamp_sec=getattr(belief_pred, "stamp_sec", 0.0), z_lin=getattr(belief_pred, "z_lin", None),
                       L=L_post[0], h=h_post[0], cert=cert)
    return post, cert, ExpectedEffect("predicted_info_trace_increase", float(r[FU["TRACE_INCREASE"]]), None)
