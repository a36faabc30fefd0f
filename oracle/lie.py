"""
SO(3)/SE(3) maps, NumPy float64.  TEST INFRASTRUCTURE (see oracle/__init__.py).

Follows fl/common/geometry/se3_jax.py: skew :43-54, se3_V :137-175,
so3_exp :259-301, so3_log :304-366, se3_exp :473-504.
"""

from __future__ import annotations

import numpy as np

SMALL_ANGLE_THRESHOLD = 1e-7  # se3_jax.py:30
NEAR_PI_THRESHOLD = 1e-7  # se3_jax.py:34


def skew(v):
    """[v]x for v of shape (...,3) -> (...,3,3)   (se3_jax.py:43-54)."""
    v = np.asarray(v, dtype=np.float64)
    z = np.zeros_like(v[..., 0])
    return np.stack(
        [
            np.stack([z, -v[..., 2], v[..., 1]], axis=-1),
            np.stack([v[..., 2], z, -v[..., 0]], axis=-1),
            np.stack([-v[..., 1], v[..., 0], z], axis=-1),
        ],
        axis=-2,
    )


def _theta_terms(phi):
    theta_sq = np.sum(phi * phi, axis=-1)
    theta = np.sqrt(theta_sq)
    small = theta < SMALL_ANGLE_THRESHOLD
    safe_theta = np.where(small, 1.0, theta)
    safe_theta_sq = np.where(theta_sq < SMALL_ANGLE_THRESHOLD**2, 1.0, theta_sq)
    return theta_sq, theta, small, safe_theta, safe_theta_sq


def so3_exp(omega):
    """Rodrigues, batched over leading axes (se3_jax.py:259-301)."""
    omega = np.asarray(omega, dtype=np.float64)
    theta_sq, theta, small, st, stsq = _theta_terms(omega)
    K = skew(omega)
    K_sq = K @ K
    sin_coeff = np.where(small, 1.0, np.sin(st) / st)
    cos_coeff = np.where(small, 0.5, (1.0 - np.cos(st)) / stsq)
    I = np.eye(3)
    return I + sin_coeff[..., None, None] * K + cos_coeff[..., None, None] * K_sq


def se3_V(phi):
    """V(phi) with Taylor switch (se3_jax.py:137-175)."""
    phi = np.asarray(phi, dtype=np.float64)
    theta_sq, theta, small, st, stsq = _theta_terms(phi)
    stcu = stsq * st
    K = skew(phi)
    K_sq = K @ K
    B = np.where(small, 0.5 - theta_sq / 24.0, (1.0 - np.cos(st)) / stsq)
    C = np.where(small, 1.0 / 6.0 - theta_sq / 120.0, (st - np.sin(st)) / stcu)
    return np.eye(3) + B[..., None, None] * K + C[..., None, None] * K_sq


def se3_exp(xi):
    """[rho,phi] -> [V(phi) rho, phi], batched (se3_jax.py:473-504)."""
    xi = np.asarray(xi, dtype=np.float64)
    rho = xi[..., :3]
    phi = xi[..., 3:6]
    V = se3_V(phi)
    t = np.einsum("...ij,...j->...i", V, rho)
    return np.concatenate([t, phi], axis=-1)


def _softmax(x):
    un = np.exp(x - np.max(x))
    return un / np.sum(un)


def so3_log(R):
    """Single-matrix log with softmax-blended near-pi axis (se3_jax.py:304-366)."""
    R = np.asarray(R, dtype=np.float64)
    cos_theta = np.clip(0.5 * (np.trace(R) - 1.0), -1.0, 1.0)
    theta = np.arccos(cos_theta)
    skew_part = 0.5 * (R - R.T)
    vex = np.array([skew_part[2, 1], skew_part[0, 2], skew_part[1, 0]])
    omega_small = vex
    sin_theta = np.sin(theta)
    safe_sin = 1.0 if abs(sin_theta) < SMALL_ANGLE_THRESHOLD else sin_theta
    omega_general = (theta / (2.0 * safe_sin)) * (2.0 * vex)
    w = _softmax(50.0 * (np.diag(R) + 1.0))
    I = np.eye(3)
    axis_col = w[0] * (R[:, 0] + I[:, 0]) + w[1] * (R[:, 1] + I[:, 1]) + w[2] * (R[:, 2] + I[:, 2])
    axis_norm = np.linalg.norm(axis_col)
    safe_axis_norm = 1.0 if axis_norm < SMALL_ANGLE_THRESHOLD else axis_norm
    omega_pi = (axis_col / safe_axis_norm) * theta
    if theta < SMALL_ANGLE_THRESHOLD:
        return omega_small
    if abs(theta - np.pi) < NEAR_PI_THRESHOLD:
        return omega_pi
    return omega_general
