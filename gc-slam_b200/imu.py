"""
IMU scan twist on the device: the step in front of DeskewConstantTwist.

Mirrors, with the same names and argument meaning,
  smooth_window_weights               fl/backend/operators/imu_preintegration.py:19-43
  preintegrate_imu_relative_pose_jax  fl/backend/operators/imu_preintegration.py:46-146
  se3_log                             fl/common/geometry/se3_jax.py:220-256
and adds the fused form the pipeline's glue amounts to (fl/backend/pipeline.py:436-483): `imu_scan_twist` runs window
weights -> preintegration -> se3_log -> rotation-only scale for all hypotheses of a scan in ONE launch and leaves
xi_body on the device, where `BinPathPlan.set_twist_from_imu` / `deskew_constant_twist` consume it.

Everything goes through the C-ABI entry gcs_imu_scan_twist (include/gcs_b200.h); no CPU fallback.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional

import numpy as np
import torch

from . import _lib as L
from . import constants
from .operators import _IO, F64


@dataclass
class ImuTwistResult:
    """Device tensors, leading axis = hypothesis.  Field names follow the reference's return tuple."""
    delta_pose: torch.Tensor        # (H, 6) [trans, rotvec], start body frame
    delta_R: torch.Tensor           # (H, 3, 3)
    delta_p: torch.Tensor           # (H, 3)
    delta_v: torch.Tensor           # (H, 3)
    ess: torch.Tensor               # (H,)  sum of the membership weights
    a_body_mean: torch.Tensor       # (H, 3)
    a_world_nog_mean: torch.Tensor  # (H, 3)
    a_world_mean: torch.Tensor      # (H, 3)
    dt_eff_sum: torch.Tensor        # (H,)
    xi_body: torch.Tensor           # (H, 6) se3_log(delta_pose), translation scaled (contiguous)
    weights: Optional[torch.Tensor]  # (H, M) membership weights, if requested


def _params(io, n_hyp, rotvec_start_WB, gyro_bias, accel_bias, gravity_W, sigma, t0, t1, trans_scale):
    p = np.zeros((n_hyp, L.IMU_NPARAM), dtype=np.float64)
    for lo, v in ((0, rotvec_start_WB), (3, gyro_bias), (6, accel_bias), (9, gravity_W)):
        if v is not None:
            v = v.detach().cpu().numpy() if isinstance(v, torch.Tensor) else np.asarray(v, dtype=np.float64)
            p[:, lo:lo + 3] = np.broadcast_to(v.reshape(-1, 3), (n_hyp, 3))
    for col, v in ((12, sigma), (13, t0), (14, t1), (15, trans_scale)):
        v = v.detach().cpu().numpy() if isinstance(v, torch.Tensor) else np.asarray(v, dtype=np.float64)
        p[:, col] = np.broadcast_to(v.reshape(-1), (n_hyp,))
    return io.dev_in(p)


def _split(out, H, weights, xi):
    o = L.IMU_OFF
    g = lambda k: out[:, o[k][0]:o[k][1]]
    return ImuTwistResult(delta_pose=g("delta_pose"), delta_R=g("delta_R").reshape(H, 3, 3), delta_p=g("delta_p"),
                          delta_v=g("delta_v"), ess=g("ess").reshape(H), a_body_mean=g("a_body_mean"),
                          a_world_nog_mean=g("a_world_nog_mean"), a_world_mean=g("a_world_mean"),
                          dt_eff_sum=g("dt_eff_sum").reshape(H), xi_body=xi, weights=weights)


def _n_hyp(*vs):
    n = 1
    for v in vs:
        if v is None:
            continue
        a = v if isinstance(v, torch.Tensor) else np.asarray(v)
        if a.ndim == 2:
            n = max(n, int(a.shape[0]))
        elif a.ndim == 1 and a.shape[0] not in (1, 3):
            n = max(n, int(a.shape[0]))
    return n


def smooth_window_weights(imu_stamps, scan_start_time: float, scan_end_time: float, sigma: float) -> torch.Tensor:
    """w(t) = sigmoid((t - start)/sigma) sigmoid((end - t)/sigma) (1 - floor) + floor   -> (M,) device tensor."""
    io = _IO()
    st = io.dev_in(imu_stamps).reshape(-1)
    M = int(st.shape[0])
    if M < 1:
        raise ValueError("smooth_window_weights: empty stamp array")
    w = io.empty(1, M)
    prm = _params(io, 1, None, None, None, None, sigma, scan_start_time, scan_end_time, 1.0)
    io.ctx.check(io.ctx.lib.gcs_imu_scan_twist(io.ctx.handle, io.stream(), L.ptr(st), None, None, M, L.ptr(prm), None, 1,
                                               None, None, L.ptr(w)))
    return w.reshape(M)


def _run(io, st, gy, ac, prm, H, weights_in, want_weights, xi_out=None):
    M = int(st.shape[0])
    if M < 1:
        raise ValueError("imu preintegration: empty stamp array")
    if tuple(gy.shape) != (M, 3) or tuple(ac.shape) != (M, 3):
        raise ValueError(f"imu preintegration: gyro / accel must be ({M}, 3), got {tuple(gy.shape)} / {tuple(ac.shape)}")
    out = io.empty(H, L.IMU_NOUT)
    xi = xi_out if xi_out is not None else io.empty(H, 6)
    if tuple(xi.shape) != (H, 6) or not xi.is_contiguous():
        raise ValueError(f"xi_out must be a contiguous ({H}, 6) device tensor")
    w = io.empty(H, M) if want_weights else None
    io.ctx.check(io.ctx.lib.gcs_imu_scan_twist(io.ctx.handle, io.stream(), L.ptr(st), L.ptr(gy), L.ptr(ac), M, L.ptr(prm),
                                               L.ptr(weights_in), H, L.ptr(out), L.ptr(xi), L.ptr(w)))
    return _split(out, H, w, xi)


def preintegrate_imu_relative_pose(imu_stamps, imu_gyro, imu_accel, weights, rotvec_start_WB, gyro_bias, accel_bias,
                                   gravity_W) -> ImuTwistResult:
    """
    Reference signature (weights supplied by the caller).  rotvec_start_WB / biases may be (3,) or (H, 3) and weights
    (M,) or (H, M): one launch integrates all hypotheses.  xi_body = se3_log(delta_pose) comes for free.
    """
    io = _IO()
    st = io.dev_in(imu_stamps).reshape(-1)
    M = int(st.shape[0])
    H = _n_hyp(rotvec_start_WB, gyro_bias, accel_bias)
    w = io.dev_in(weights)
    if w.numel() == M and H > 1:
        w = w.reshape(1, M).expand(H, M).contiguous()
    if w.numel() != H * M:
        raise ValueError(f"weights must have {M} or {H}x{M} entries, got {w.numel()}")
    prm = _params(io, H, rotvec_start_WB, gyro_bias, accel_bias, gravity_W, 1.0, 0.0, 0.0, 1.0)
    return _run(io, st, io.dev_in(imu_gyro).reshape(-1, 3), io.dev_in(imu_accel).reshape(-1, 3), prm, H,
                w.reshape(H, M).contiguous(), False)


def imu_scan_twist(imu_stamps, imu_gyro, imu_accel, scan_start_time, scan_end_time, sigma_warp, rotvec_start_WB,
                   gyro_bias, accel_bias, gravity_W=None, deskew_rotation_only: bool = False, want_weights: bool = False,
                   xi_out: Optional[torch.Tensor] = None) -> ImuTwistResult:
    """
    pipeline.py:436-483 in one launch: within-scan membership weights (sigma_warp), preintegration, xi_body =
    se3_log(delta_pose), translation zeroed for rotation-only deskew.  Per-hypothesis arguments may carry a leading
    hypothesis axis.  xi_out: a contiguous (H, 6) device tensor to receive xi_body in place (e.g. a BinPathPlan's).
    """
    io = _IO()
    if gravity_W is None:
        gravity_W = np.asarray(constants.GC_GRAVITY_W, dtype=np.float64)
    H = _n_hyp(rotvec_start_WB, gyro_bias, accel_bias, np.atleast_1d(np.asarray(sigma_warp, dtype=np.float64))
               if not isinstance(sigma_warp, torch.Tensor) else sigma_warp)
    prm = _params(io, H, rotvec_start_WB, gyro_bias, accel_bias, gravity_W, sigma_warp, scan_start_time, scan_end_time,
                  0.0 if deskew_rotation_only else 1.0)
    return _run(io, io.dev_in(imu_stamps).reshape(-1), io.dev_in(imu_gyro).reshape(-1, 3),
                io.dev_in(imu_accel).reshape(-1, 3), prm, H, None, want_weights, xi_out)
