"""
IMU window weights + preintegration -> constant scan twist xi_body, NumPy float64.
TEST INFRASTRUCTURE (see oracle/__init__.py): only tests/, __graft_entry__.smoke() and bench.py's CPU legs import it.

Restates, line for line:
  smooth_window_weights               fl/backend/operators/imu_preintegration.py:19-43
  preintegrate_imu_relative_pose_jax  fl/backend/operators/imu_preintegration.py:46-146
  _se3_V_inv, se3_log                 fl/common/geometry/se3_jax.py:178-256
  the deskew-twist glue of the pipeline (sigma_warp, within-scan window, se3_log, rotation-only scale)
                                      fl/backend/pipeline.py:436-483
Pinned to the reference's own sources by tests/golden/make_golden_imu.py (tests/golden/imu_*.npz).
"""
from __future__ import annotations

import numpy as np

from . import lie

GC_WEIGHT_FLOOR = 1e-12  # fl/common/constants.py (strictly positive continuous floor)


def _sigmoid(x):
    # jax.nn.sigmoid = logistic; evaluated as 1/(1+exp(-x)) like XLA's expansion
    with np.errstate(over="ignore"):
        return 1.0 / (1.0 + np.exp(-x))


def smooth_window_weights(imu_stamps, scan_start_time, scan_end_time, sigma):
    """imu_preintegration.py:19-43."""
    t = np.asarray(imu_stamps, dtype=np.float64)
    sig = max(float(sigma), 1e-6)
    a = (t - float(scan_start_time)) / sig
    b = (float(scan_end_time) - t) / sig
    w_raw = _sigmoid(a) * _sigmoid(b)
    return w_raw * (1.0 - GC_WEIGHT_FLOOR) + GC_WEIGHT_FLOOR


def preintegrate_imu_relative_pose(imu_stamps, imu_gyro, imu_accel, weights, rotvec_start_WB, gyro_bias, accel_bias,
                                   gravity_W):
    """imu_preintegration.py:46-146: sequential fixed-cost integration, relative pose in the start body frame."""
    imu_stamps = np.asarray(imu_stamps, dtype=np.float64).reshape(-1)
    imu_gyro = np.asarray(imu_gyro, dtype=np.float64)
    imu_accel = np.asarray(imu_accel, dtype=np.float64)
    w = np.asarray(weights, dtype=np.float64).reshape(-1)
    gyro_bias = np.asarray(gyro_bias, dtype=np.float64)
    accel_bias = np.asarray(accel_bias, dtype=np.float64)
    gravity_W = np.asarray(gravity_W, dtype=np.float64)
    ess = np.sum(w)
    dt = np.concatenate([imu_stamps[1:] - imu_stamps[:-1], np.zeros(1)])
    dt = np.maximum(dt, 0.0)
    R = lie.so3_exp(np.asarray(rotvec_start_WB, dtype=np.float64))
    v = np.zeros(3)
    p = np.zeros(3)
    sum_wdt = 0.0
    sum_a_body = np.zeros(3)
    sum_a_world_nog = np.zeros(3)
    sum_a_world = np.zeros(3)
    for i in range(imu_stamps.shape[0]):
        dt_eff = w[i] * dt[i]
        omega = imu_gyro[i] - gyro_bias
        dR = lie.so3_exp(omega * dt_eff)
        R_next = R @ dR
        a_body = imu_accel[i] - accel_bias
        a_world_nog = R @ a_body
        a_world = a_world_nog + gravity_W
        sum_wdt = sum_wdt + dt_eff
        sum_a_body = sum_a_body + a_body * dt_eff
        sum_a_world_nog = sum_a_world_nog + a_world_nog * dt_eff
        sum_a_world = sum_a_world + a_world * dt_eff
        v_next = v + a_world * dt_eff
        p = p + v * dt_eff + 0.5 * a_world * (dt_eff * dt_eff)
        v = v_next
        R = R_next
    R_start = lie.so3_exp(np.asarray(rotvec_start_WB, dtype=np.float64))
    delta_R = R_start.T @ R
    rotvec_delta = lie.so3_log(delta_R)
    p_body = R_start.T @ p
    v_body = R_start.T @ v
    delta_pose = np.concatenate([p_body, rotvec_delta])
    denom = max(sum_wdt, 1e-12)
    return dict(delta_pose=delta_pose, delta_R=delta_R, delta_p=p_body, delta_v=v_body, ess=ess,
                a_body_mean=sum_a_body / denom, a_world_nog_mean=sum_a_world_nog / denom,
                a_world_mean=sum_a_world / denom, dt_eff_sum=sum_wdt)


def se3_V_inv(phi):
    """se3_jax.py:178-218."""
    phi = np.asarray(phi, dtype=np.float64).reshape(-1)
    theta_sq = float(phi @ phi)
    theta = np.sqrt(theta_sq)
    K = lie.skew(phi)
    K_sq = K @ K
    small = theta < lie.SMALL_ANGLE_THRESHOLD
    safe_theta = 1.0 if small else theta
    safe_theta_sq = 1.0 if theta_sq < lie.SMALL_ANGLE_THRESHOLD ** 2 else theta_sq
    denom = 2.0 * safe_theta * np.sin(safe_theta) + 1e-12
    D = (1.0 / 12.0 + theta_sq / 720.0) if small else (1.0 / safe_theta_sq) - (1.0 + np.cos(safe_theta)) / denom
    return np.eye(3) - 0.5 * K + D * K_sq


def se3_log(T):
    """se3_jax.py:220-256: [t, rotvec] -> [V(phi)^-1 t, phi] with phi = Log(Exp(rotvec))."""
    T = np.asarray(T, dtype=np.float64).reshape(-1)
    phi = lie.so3_log(lie.so3_exp(T[3:6]))
    rho = se3_V_inv(phi) @ T[:3]
    return np.concatenate([rho, phi])


def imu_scan_twist(imu_stamps, imu_gyro, imu_accel, scan_start_time, scan_end_time, sigma_warp, rotvec_start_WB,
                   gyro_bias, accel_bias, gravity_W, deskew_rotation_only=False):
    """pipeline.py:436-483: within-scan window -> preintegration -> xi_body = se3_log(delta_pose), translation scaled."""
    w = smooth_window_weights(imu_stamps, scan_start_time, scan_end_time, sigma_warp)
    pre = preintegrate_imu_relative_pose(imu_stamps, imu_gyro, imu_accel, w, rotvec_start_WB, gyro_bias, accel_bias,
                                         gravity_W)
    xi = se3_log(pre["delta_pose"])
    xi[:3] = xi[:3] * (0.0 if deskew_rotation_only else 1.0)
    pre["weights"] = w
    pre["xi_body"] = xi
    return pre
