"""
HypothesisBarycenterProjection core, NumPy float64.  TEST INFRASTRUCTURE (see oracle/__init__.py).

Restates, line for line:
  _hypothesis_barycenter_core      fl/backend/operators/hypothesis.py:51-115
  domain_projection_psd_core       fl/common/primitives.py:80-123      (any dimension)
  spd_cholesky_solve_lifted_core   fl/common/primitives.py:141-165
Pinned to the reference's own sources by tests/golden/make_golden_hyp.py (tests/golden/hyp_*.npz).
"""
from __future__ import annotations

import numpy as np


def domain_projection_psd(M, eps_psd=1e-12):
    """primitives.py:80-123 -> (M_psd, [projection_delta, sym_delta, eig_min, eig_max, cond, near_null_count])."""
    M = np.asarray(M, dtype=np.float64)
    M_sym = 0.5 * (M + M.T)
    sym_delta = np.linalg.norm(M_sym - M, ord="fro")
    eigvals, eigvecs = np.linalg.eigh(M_sym)
    vals = np.maximum(eigvals, eps_psd)
    M_psd = eigvecs @ np.diag(vals) @ eigvecs.T
    projection_delta = np.linalg.norm(M_psd - M_sym, ord="fro")
    near_null = float(np.sum(vals < 10.0 * eps_psd))
    return M_psd, np.array([projection_delta, sym_delta, np.min(vals), np.max(vals), np.max(vals) / np.min(vals), near_null])


def spd_cholesky_solve_lifted(L, b, eps_lift=1e-9):
    """primitives.py:141-165."""
    L = np.asarray(L, dtype=np.float64)
    d = L.shape[0]
    from scipy.linalg import solve_triangular
    A = L + eps_lift * np.eye(d)
    C = np.linalg.cholesky(0.5 * (A + A.T))      # jnp.linalg.cholesky: symmetrize_input=True
    y = solve_triangular(C, np.asarray(b, dtype=np.float64), lower=True)
    return solve_triangular(C.T, y, lower=False)


def hypothesis_barycenter(L_stack, h_stack, z_lin_stack, weights, weight_floor=0.0025, eps_psd=1e-12, eps_lift=1e-9):
    """hypothesis.py:51-115."""
    L_stack = np.asarray(L_stack, dtype=np.float64)
    h_stack = np.asarray(h_stack, dtype=np.float64)
    z_lin_stack = np.asarray(z_lin_stack, dtype=np.float64)
    weights = np.asarray(weights, dtype=np.float64)
    w_fl = np.maximum(weights, weight_floor)
    floor_adjustment = np.sum(np.abs(w_fl - weights))
    wn = w_fl / np.sum(w_fl)
    L_raw = np.einsum("k,kij->ij", wn, L_stack)
    h_out = np.einsum("k,ki->i", wn, h_stack)
    z_out = np.einsum("k,ki->i", wn, z_lin_stack)
    L_out, cert = domain_projection_psd(L_raw, eps_psd)
    means = np.stack([spd_cholesky_solve_lifted(L_stack[k], h_stack[k], eps_lift) for k in range(L_stack.shape[0])])
    mom = np.einsum("k,ki->i", wn, means)
    d = means - mom[None, :]
    spread = np.sum(wn * np.sum(d * d, axis=1))
    return dict(L=L_out, h=h_out, z_lin=z_out, floor_adjustment=floor_adjustment, weights_normalized=wn, psd_cert=cert,
                spread_proxy=spread, means=means)
