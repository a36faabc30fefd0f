"""
Bin-family LiDAR evidence operators, NumPy float64.  TEST INFRASTRUCTURE
(see oracle/__init__.py).  Each function returns ``(result: dict, cert: dict)``
where ``cert`` holds the scalars the reference puts into its CertBundle /
ExpectedEffect.

Reference anchors (``fl/`` = fl_ws/src/fl_slam_poc/fl_slam_poc/):
  point_budget_resample    fl/backend/operators/point_budget.py:50-109,117-221
  smooth_window_weights    fl/backend/operators/imu_preintegration.py:19-43
  deskew_constant_twist    fl/backend/operators/deskew_constant_twist.py:31-117
  ray_directions           fl/backend/pipeline.py:589-593
  fibonacci_atlas          archive/bin_atlas.py:40-75
  bin_soft_assign          archive/legacy_operators/binning.py:56-131
  scan_bin_moment_match    archive/legacy_operators/binning.py:139-324
  kappa_*                  fl/backend/operators/kappa.py:84-234
  psd_project / inv_mass   fl/common/primitives.py:80-123,195-212
  map bin stats            archive/bin_atlas.py:83-257
  matrix_fisher_rotation   archive/legacy_operators/matrix_fisher_evidence.py:83-394
  planar_translation       archive/legacy_operators/matrix_fisher_evidence.py:413-671
  combined 22-D evidence   archive/legacy_operators/matrix_fisher_evidence.py:729-756
"""

from __future__ import annotations

import math

import numpy as np

from . import lie

EPS_PSD = 1e-12  # fl/common/constants.py:70
EPS_LIFT = 1e-9  # :71
EPS_MASS = 1e-12  # :72
EPS_R = 1e-6  # :73
KAPPA_BLEND_R0 = 0.8  # :95
KAPPA_BLEND_TAU = 0.03  # :96
TIME_WARP_SIGMA_FRAC = 0.1  # :141
WEIGHT_FLOOR = 1e-12  # :237
N_POINTS_CAP = 8192  # :64
D_Z = 22  # :58
B_BINS = 48  # prose only: CHANGELOG.md:492, docs/PIPELINE_DESIGN_GAPS.md:257


def _sigmoid(x):
    with np.errstate(over="ignore"):
        return 1.0 / (1.0 + np.exp(-x))


# --------------------------------------------------------------------------- a1
def point_budget_resample(points, timestamps, weights, ring=None, tag=None,
                          n_points_cap=N_POINTS_CAP, eps_mass=EPS_MASS):
    points = np.asarray(points, dtype=np.float64).reshape(-1, 3)
    timestamps = np.asarray(timestamps, dtype=np.float64).reshape(-1)
    weights = np.asarray(weights, dtype=np.float64).reshape(-1)
    n_input = points.shape[0]
    ring = np.zeros(n_input, np.uint8) if ring is None else np.asarray(ring, np.uint8).reshape(-1)
    tag = np.zeros(n_input, np.uint8) if tag is None else np.asarray(tag, np.uint8).reshape(-1)
    stride = max(1, int(math.ceil(n_input / n_points_cap)))  # point_budget.py:160

    total_mass_in = np.sum(weights)
    idx = np.arange(0, n_input, stride)
    n_sel = idx.shape[0]
    pts_out = np.zeros((n_points_cap, 3))
    pts_out[:n_sel] = points[idx]
    t_out = np.zeros(n_points_cap)
    t_out[:n_sel] = timestamps[idx]
    w_raw = weights[idx]
    mass_scale = total_mass_in / (np.sum(w_raw) + eps_mass)
    w_out = np.zeros(n_points_cap)
    w_out[:n_sel] = w_raw * mass_scale
    ring_out = np.zeros(n_points_cap, np.uint8)
    ring_out[:n_sel] = ring[idx]
    tag_out = np.zeros(n_points_cap, np.uint8)
    tag_out[:n_sel] = tag[idx]
    w_norm = w_out / (total_mass_in + eps_mass)
    ess = 1.0 / np.sum(w_norm**2 + eps_mass)  # eps inside the sum (point_budget.py:96)

    res = dict(points=pts_out, timestamps=t_out, weights=w_out, ring=ring_out, tag=tag_out,
               indices=idx.astype(np.int64), stride=stride, n_input=int(n_input), n_output=int(n_sel),
               total_mass_in=float(total_mass_in), total_mass_out=float(total_mass_in))
    cert = dict(exact=False, triggers=["PointBudgetResample"], ess_total=float(ess),
                support_frac=float(min(1.0, n_points_cap / (n_input + EPS_MASS))),
                mass_epsilon_ratio=EPS_MASS / (float(total_mass_in) + EPS_MASS),
                effect_name="predicted_ess", effect_predicted=float(ess))
    return res, cert


# --------------------------------------------------------------------------- a2
def smooth_window_weights(stamps, scan_start_time, scan_end_time, sigma):
    t = np.asarray(stamps, dtype=np.float64)
    sig = max(float(sigma), 1e-6)
    a = (t - scan_start_time) / sig
    b = (scan_end_time - t) / sig
    return _sigmoid(a) * _sigmoid(b) * (1.0 - WEIGHT_FLOOR) + WEIGHT_FLOOR


def deskew_constant_twist(points, timestamps, weights, scan_start_time, scan_end_time, xi_body,
                          ess_imu=1.0):
    points = np.asarray(points, dtype=np.float64).reshape(-1, 3)
    timestamps = np.asarray(timestamps, dtype=np.float64).reshape(-1)
    weights = np.asarray(weights, dtype=np.float64).reshape(-1)
    xi = np.asarray(xi_body, dtype=np.float64).reshape(-1)
    t0 = np.float64(scan_start_time)
    t1 = np.float64(scan_end_time)
    denom = max(t1 - t0, 1e-12)
    alpha = (timestamps - t0) / denom  # no clipping
    T = lie.se3_exp(alpha[:, None] * xi[None, :])  # (N,6) [t, rotvec]
    R = lie.so3_exp(T[:, 3:6])
    # p0 = R^T (p - t)
    p0 = np.einsum("nji,nj->ni", R, points - T[:, :3])
    sigma = TIME_WARP_SIGMA_FRAC * denom
    w_time = smooth_window_weights(timestamps, t0, t1, sigma)
    w_out = weights * w_time
    retained = float(np.sum(w_out) / (np.sum(weights) + EPS_MASS))
    res = dict(points=p0, timestamps=timestamps, weights=w_out, ess_imu=float(ess_imu))
    cert = dict(exact=True, triggers=[], ess_total=float(ess_imu), support_frac=retained,
                effect_name="deskew_variance_reduction_proxy", effect_predicted=0.0)
    return res, cert


# --------------------------------------------------------------------------- a3
def ray_directions(points, origin, eps=EPS_MASS):
    rays = np.asarray(points, np.float64) - np.asarray(origin, np.float64)[None, :]
    norms = np.linalg.norm(rays, axis=1, keepdims=True)
    return rays / (norms + eps)


# --------------------------------------------------------------------------- a4
def fibonacci_atlas(n_bins=B_BINS):
    i = np.arange(n_bins, dtype=np.float64) + 0.5
    phi = np.arccos(1 - 2 * i / n_bins)
    theta = np.pi * (1 + np.sqrt(5)) * i
    dirs = np.stack([np.sin(phi) * np.cos(theta), np.sin(phi) * np.sin(theta), np.cos(phi)], axis=1)
    return dirs / (np.linalg.norm(dirs, axis=1, keepdims=True) + EPS_MASS)


def bin_soft_assign(point_directions, bin_directions, tau, eps_mass=EPS_MASS):
    d = np.asarray(point_directions, np.float64)
    b = np.asarray(bin_directions, np.float64)
    n_points = d.shape[0]
    sim = d @ b.T
    x = sim / tau
    un = np.exp(x - np.max(x, axis=1, keepdims=True))
    resp = un / np.sum(un, axis=1, keepdims=True)
    max_resp = np.max(resp)
    ent = -np.sum(resp * np.log(resp + eps_mass), axis=1)
    avg_entropy = np.sum(ent) / (float(n_points) + eps_mass)  # N counts padded rows
    res = dict(responsibilities=resp)
    cert = dict(exact=True, triggers=[], ess_total=float(np.exp(avg_entropy)),
                support_frac=float(max_resp), effect_name="predicted_assignment_entropy",
                effect_predicted=float(avg_entropy))
    return res, cert


# ---------------------------------------------------------------- L1 numerics
def inv_mass(m, eps_mass=EPS_MASS):
    denom = np.asarray(m, np.float64) + eps_mass + np.finfo(np.float64).eps
    return 1.0 / denom, eps_mass / denom


def psd_project(M, eps_psd=EPS_PSD):
    """-> (M_psd, [projection_delta, sym_delta, eig_min, eig_max, cond, near_null_count])."""
    M = np.asarray(M, np.float64)
    M_sym = 0.5 * (M + M.T)
    sym_delta = np.linalg.norm(M_sym - M, ord="fro")
    vals, vecs = np.linalg.eigh(M_sym)
    vc = np.maximum(vals, eps_psd)
    M_psd = vecs @ np.diag(vc) @ vecs.T
    delta = np.linalg.norm(M_psd - M_sym, ord="fro")
    return M_psd, np.array([delta, sym_delta, vc.min(), vc.max(), vc.max() / vc.min(),
                            float(np.sum(vc < 10.0 * eps_psd))])


# --------------------------------------------------------------------------- a6
def kappa_from_resultant_batch(R_bar, eps_r=EPS_R, d=3, r0=KAPPA_BLEND_R0, tau=KAPPA_BLEND_TAU):
    R = np.clip(np.asarray(R_bar, np.float64), 0.0, 1.0 - eps_r)
    R2 = R * R
    k_low = (R * (d - R2)) / (1.0 - R2 + eps_r)
    k_high = -np.log(np.maximum(1.0 - R2, eps_r))
    s = _sigmoid((R - r0) / max(tau, 1e-6))
    return (1.0 - s) * k_low + s * k_high


def kappa_from_resultant_v2(R_bar, eps_r=EPS_R):
    """Scalar variant (kappa.py:172-234, :84-127); cert approx ["KappaLowRApproximation"]."""
    x = float(R_bar)
    Rc = min(max(x, 0.0), 1.0 - eps_r)
    R2 = Rc * Rc
    k_low = (Rc * (3.0 - R2)) / (1.0 - R2 + eps_r)
    k_high = -math.log(max(1.0 - R2, eps_r))
    s = 1.0 / (1.0 + math.exp(-(Rc - KAPPA_BLEND_R0) / max(KAPPA_BLEND_TAU, 1e-6)))
    kappa = (1.0 - s) * k_low + s * k_high
    res = dict(kappa=kappa, R_clamped=Rc, clamp_delta=abs(Rc - x))
    cert = dict(exact=False, triggers=["KappaLowRApproximation"], effect_name="kappa",
                effect_predicted=kappa)
    return res, cert


# --------------------------------------------------------------------------- a5
def scan_bin_moment_match(points, point_covariances, weights, responsibilities, point_lambda=None,
                          direction_origin=None, eps_psd=EPS_PSD, eps_mass=EPS_MASS):
    p = np.asarray(points, np.float64).reshape(-1, 3)
    n = p.shape[0]
    w = np.asarray(weights, np.float64).reshape(-1)
    r = np.asarray(responsibilities, np.float64)
    lam = np.ones(n) if point_lambda is None else np.asarray(point_lambda, np.float64).reshape(-1)
    o = np.zeros(3) if direction_origin is None else np.asarray(direction_origin, np.float64).reshape(-1)
    w_r = (w * lam)[:, None] * r
    rays = p - o[None, :]
    d = rays / (np.linalg.norm(rays, axis=1, keepdims=True) + eps_mass)

    N = np.sum(w_r, axis=0)
    s_dir = w_r.T @ d
    S_scatter = np.einsum("nb,ni,nj->bij", w_r, d, d, optimize=True)
    sum_p = w_r.T @ p
    sum_ppT = np.einsum("nb,ni,nj->bij", w_r, p, p, optimize=True)
    if point_covariances is None:
        sum_cov = np.zeros_like(sum_ppT)
    else:
        sum_cov = np.einsum("nb,nij->bij", w_r, np.asarray(point_covariances, np.float64), optimize=True)

    inv_N, eps_ratio = inv_mass(N, eps_mass)
    p_bar = sum_p * inv_N[:, None]
    scatter = sum_ppT * inv_N[:, None, None] - np.einsum("bi,bj->bij", p_bar, p_bar)
    Sigma_raw = scatter + sum_cov * inv_N[:, None, None]
    B = N.shape[0]
    Sigma_p = np.zeros((B, 3, 3))
    delta_total = 0.0
    for b in range(B):
        Sigma_p[b], cv = psd_project(Sigma_raw[b], eps_psd)
        delta_total += cv[0]
    Rbar = np.linalg.norm(s_dir, axis=1) * inv_N
    kappa = kappa_from_resultant_batch(Rbar, eps_r=EPS_R)
    total = np.sum(N)
    ess = total**2 / (np.sum(N**2) + eps_mass)
    support_frac = np.mean(N / (N + eps_mass))
    res = dict(N=N, s_dir=s_dir, S_dir_scatter=S_scatter, p_bar=p_bar, Sigma_p=Sigma_p,
               kappa_scan=kappa, sum_p=sum_p, sum_ppT=sum_ppT, Sigma_raw=Sigma_raw)
    cert = dict(exact=False, triggers=["ScanBinMomentMatch"], ess_total=float(ess),
                support_frac=float(support_frac), psd_projection_delta=float(delta_total),
                mass_epsilon_ratio=float(np.max(eps_ratio)), effect_name="predicted_ess",
                effect_predicted=float(ess))
    return res, cert


# --------------------------------------------------------------------------- a9
def empty_map_stats(n_bins=B_BINS):
    return dict(S_dir=np.zeros((n_bins, 3)), S_dir_scatter=np.zeros((n_bins, 3, 3)),
                N_dir=np.zeros(n_bins), N_pos=np.zeros(n_bins), sum_p=np.zeros((n_bins, 3)),
                sum_ppT=np.zeros((n_bins, 3, 3)))


def update_map_stats(ms, inc_S_dir, inc_S_scatter, inc_N_dir, inc_N_pos, inc_sum_p, inc_sum_ppT):
    return dict(S_dir=ms["S_dir"] + inc_S_dir, S_dir_scatter=ms["S_dir_scatter"] + inc_S_scatter,
                N_dir=ms["N_dir"] + inc_N_dir, N_pos=ms["N_pos"] + inc_N_pos,
                sum_p=ms["sum_p"] + inc_sum_p, sum_ppT=ms["sum_ppT"] + inc_sum_ppT)


def apply_forgetting(ms, forgetting_factor=0.99):
    g = float(forgetting_factor)
    return {k: g * v for k, v in ms.items()}


def map_derived_stats(ms, eps_mass=EPS_MASS, eps_psd=EPS_PSD):
    """(mu_dir, kappa, centroid, Sigma_c)   archive/bin_atlas.py:159-198; safe_normalize = v/(|v|+eps)."""
    S = ms["S_dir"]
    norms = np.linalg.norm(S, axis=1)
    mu_dir = S / (norms[:, None] + eps_mass)
    inv_Nd, _ = inv_mass(ms["N_dir"], eps_mass)
    kappa = kappa_from_resultant_batch(norms * inv_Nd, eps_r=EPS_R)
    inv_Np, _ = inv_mass(ms["N_pos"], eps_mass)
    centroid = ms["sum_p"] * inv_Np[:, None]
    Sigma_raw = ms["sum_ppT"] * inv_Np[:, None, None] - np.einsum("bi,bj->bij", centroid, centroid)
    Sigma_c = np.stack([psd_project(Sigma_raw[b], eps_psd)[0] for b in range(S.shape[0])])
    return mu_dir, kappa, centroid, Sigma_c


def pushforward_scan_stats_to_map(scan, R, t, planar_z=True):
    """
    Rigid pushforward of additive scan-bin statistics into the world frame with the
    START-of-scan pose, ``t[2]`` forced to 0 (CHANGELOG.md:575-578, :684-721).  The operator
    that did this (PoseCovInflationPushforward, fl/backend/operators/map_update.py) was deleted
    from the reference: PARITY UNPINNED -- derived from prose and from the rigid-transform
    identities used by its successor (fl/backend/pipeline.py:1248-1256).
    Returns increments (S_dir, S_dir_scatter, N_dir, N_pos, sum_p, sum_ppT).
    """
    R = np.asarray(R, np.float64)
    t = np.array(t, dtype=np.float64).reshape(3)
    if planar_z:
        t[2] = 0.0
    N = scan["N"]
    Rs = scan["sum_p"] @ R.T
    inc_S_dir = scan["s_dir"] @ R.T
    inc_S_scatter = np.einsum("ij,bjk,lk->bil", R, scan["S_dir_scatter"], R)
    inc_sum_p = Rs + N[:, None] * t[None, :]
    inc_sum_ppT = (np.einsum("ij,bjk,lk->bil", R, scan["sum_ppT"], R)
                   + np.einsum("bi,j->bij", Rs, t) + np.einsum("i,bj->bij", t, Rs)
                   + N[:, None, None] * np.outer(t, t)[None])
    return inc_S_dir, inc_S_scatter, N.copy(), N.copy(), inc_sum_p, inc_sum_ppT


# --------------------------------------------------------------------------- a7
def scatter_metrics(S_scatter, N_total, eps=EPS_MASS):
    T = np.asarray(S_scatter, np.float64) * (1.0 / (N_total + eps))
    vals_asc, vecs = np.linalg.eigh(T)
    idx = np.argsort(vals_asc, kind="stable")[::-1]
    vals = np.maximum(vals_asc[idx], 0.0)
    vecs = vecs[:, idx]
    l1, l2, l3 = vals
    inv1 = 1.0 / (l1 + eps)
    total = l1 + l2 + l3 + eps
    p1, p2, p3 = l1 / total, l2 / total, l3 / total
    ent = -(p1 * np.log(p1 + eps) + p2 * np.log(p2 + eps) + p3 * np.log(p3 + eps))
    return dict(eigenvalues=vals, eigenvectors=vecs, linearity=float((l1 - l2) * inv1),
                planarity=float((l2 - l3) * inv1), sphericity=float(l3 * inv1),
                anisotropy=float(1.0 - l3 * inv1), effective_rank=float(np.exp(ent)))


def matrix_fisher_rotation(R_pred, scan_s_dir, scan_S_scatter, scan_N, map_S_dir, map_S_scatter,
                           map_N_dir, eps_psd=EPS_PSD, eps_mass=EPS_MASS):
    eps = eps_mass
    scan_s_dir = np.asarray(scan_s_dir, np.float64)
    map_S_dir = np.asarray(map_S_dir, np.float64)
    scan_N = np.asarray(scan_N, np.float64)
    map_N = np.asarray(map_N_dir, np.float64)
    w_b = np.sqrt(scan_N * map_N + eps)
    sn = np.linalg.norm(scan_s_dir, axis=1, keepdims=True)
    mn = np.linalg.norm(map_S_dir, axis=1, keepdims=True)
    u_scan = scan_s_dir / (sn + eps)
    u_map = map_S_dir / (mn + eps)
    Rbar_s = sn[:, 0] * (1.0 / (scan_N + eps))
    Rbar_m = mn[:, 0] * (1.0 / (map_N + eps))
    w_final = w_b * (Rbar_s * Rbar_m)
    H = np.einsum("b,bi,bj->ij", w_final, u_map, u_scan)
    U, s, Vt = np.linalg.svd(H, full_matrices=True)
    det_sign = np.linalg.det(U @ Vt)
    Uc = U.copy()
    Uc[:, 2] = U[:, 2] * np.sign(det_sign)
    R_mf = Uc @ Vt
    V = Vt.T
    L_raw = V @ np.diag([s[1] + s[2], s[0] + s[2], s[0] + s[1]]) @ V.T
    N_eff = np.sum(w_final)
    scan_tot = np.sum(np.asarray(scan_S_scatter, np.float64), axis=0)
    map_tot = np.sum(np.asarray(map_S_scatter, np.float64), axis=0)

    R_pred = np.asarray(R_pred, np.float64)
    delta_rot = lie.so3_log(R_pred.T @ R_mf)
    L_rot, psd_cert = psd_project(L_raw, eps_psd)
    h_rot = L_rot @ delta_rot
    rot_nll = 0.5 * float(delta_rot @ L_rot @ delta_rot)
    eig_min, eig_max = float(np.min(s)), float(np.max(s))
    res = dict(R_mf=R_mf, L_rot=L_rot, h_rot=h_rot, delta_rot=delta_rot, svd_singular_values=s, H=H,
               scan_scatter_metrics=scatter_metrics(scan_tot, float(np.sum(scan_N)), eps),
               map_scatter_metrics=scatter_metrics(map_tot, float(np.sum(map_N)), eps))
    cert = dict(exact=False, triggers=["MatrixFisherRotationEvidence"], eig_min=eig_min, eig_max=eig_max,
                cond=eig_max / (eig_min + eps), near_null_count=int(np.sum(s < eps)),
                nll_per_ess=rot_nll / (N_eff + eps), directional_score=float(np.sum(s)),
                psd_projection_delta=float(psd_cert[0]), mass_epsilon_ratio=float(eps / (N_eff + eps)),
                effect_name="predicted_rotation_nll", effect_predicted=rot_nll, N_eff=float(N_eff))
    return res, cert


# --------------------------------------------------------------------------- a8
def planar_translation(t_pred, scan_p_bar, scan_Sigma_p, scan_N, map_centroid, map_Sigma_c, map_N_pos,
                       map_S_scatter, map_N_dir, R_hat, eps_psd=EPS_PSD, eps_mass=EPS_MASS):
    eps = eps_mass
    R_hat = np.asarray(R_hat, np.float64)
    T_map = np.sum(np.asarray(map_S_scatter, np.float64), axis=0) / (np.sum(map_N_dir) + eps)
    ev = np.sort(np.linalg.eigvalsh(T_map))[::-1]
    z_scale = max(ev[2], 0.0) / max(ev[0], eps)

    p_rot = np.einsum("ij,bj->bi", R_hat, scan_p_bar)
    t_b = np.asarray(map_centroid, np.float64) - p_rot
    Sig = np.asarray(map_Sigma_c, np.float64) + np.einsum("ij,bjk,lk->bil", R_hat, scan_Sigma_p, R_hat)
    w_b = np.sqrt(np.asarray(scan_N, np.float64) * np.asarray(map_N_pos, np.float64) + eps)
    Winv = np.stack([w * np.linalg.inv(S + eps * np.eye(3)) for S, w in zip(Sig, w_b)])
    L_full = np.sum(Winv, axis=0)
    h_full = np.sum(np.einsum("bij,bj->bi", Winv, t_b), axis=0)
    t_wls = np.linalg.solve(L_full + eps * np.eye(3), h_full)
    mask = np.array([1.0, 1.0, z_scale])
    L_raw = L_full * mask[:, None] * mask[None, :]
    N_eff = np.sum(w_b)

    delta = t_wls - np.asarray(t_pred, np.float64)
    L_trans, psd_cert = psd_project(L_raw, eps_psd)
    h_trans = L_trans @ delta
    nll = 0.5 * float(delta @ L_trans @ delta)
    eigs = np.linalg.eigvalsh(L_trans)
    eig_min, eig_max = float(np.min(eigs)), float(np.max(eigs))
    res = dict(t_wls=t_wls, L_trans=L_trans, h_trans=h_trans, delta_trans=delta,
               xy_info_scale=float(0.5 * (L_trans[0, 0] + L_trans[1, 1])),
               z_info_scale=float(L_trans[2, 2]), z_precision_scale=float(z_scale))
    cert = dict(exact=False, triggers=["PlanarTranslationEvidence"], eig_min=eig_min, eig_max=eig_max,
                cond=eig_max / (eig_min + eps), near_null_count=int(np.sum(eigs < eps)),
                nll_per_ess=nll / (N_eff + eps), directional_score=0.0,
                psd_projection_delta=float(psd_cert[0]), mass_epsilon_ratio=float(eps / (N_eff + eps)),
                effect_name="predicted_translation_nll", effect_predicted=nll, N_eff=float(N_eff))
    return res, cert


def combined_lidar_evidence_22d(L_trans, h_trans, L_rot, h_rot):
    L = np.zeros((D_Z, D_Z))
    h = np.zeros(D_Z)
    L[0:3, 0:3] = L_trans
    h[0:3] = h_trans
    L[3:6, 3:6] = L_rot
    h[3:6] = h_rot
    return L, h


# ------------------------------------------------------- whole bin path, one scan x one hypothesis
def lidar_evidence_bins(raw_points, raw_t, raw_w, raw_ring, raw_tag, n_points_cap, xi_body, t0, t1,
                        origin, bin_dirs, tau, map_stats, R_pred, t_pred):
    """README.md:105-120 steps 1,3,4,5,6,7,8 + LiDAR term of step 9, chained as the legacy pipeline did."""
    rs, c_rs = point_budget_resample(raw_points, raw_t, raw_w, raw_ring, raw_tag, n_points_cap)
    dk, c_dk = deskew_constant_twist(rs["points"], rs["timestamps"], rs["weights"], t0, t1, xi_body)
    dirs = ray_directions(dk["points"], origin)
    sa, c_sa = bin_soft_assign(dirs, bin_dirs, tau)
    st, c_st = scan_bin_moment_match(dk["points"], None, dk["weights"], sa["responsibilities"],
                                     direction_origin=origin)
    mf, c_mf = matrix_fisher_rotation(R_pred, st["s_dir"], st["S_dir_scatter"], st["N"],
                                      map_stats["S_dir"], map_stats["S_dir_scatter"], map_stats["N_dir"])
    mu_dir, kappa_map, centroid, Sigma_c = map_derived_stats(map_stats)
    pt, c_pt = planar_translation(t_pred, st["p_bar"], st["Sigma_p"], st["N"], centroid, Sigma_c,
                                  map_stats["N_pos"], map_stats["S_dir_scatter"], map_stats["N_dir"],
                                  mf["R_mf"])
    L, h = combined_lidar_evidence_22d(pt["L_trans"], pt["h_trans"], mf["L_rot"], mf["h_rot"])
    return dict(resample=rs, deskew=dk, directions=dirs, soft_assign=sa, stats=st, mf=mf, trans=pt, L=L, h=h,
                certs=dict(resample=c_rs, deskew=c_dk, soft_assign=c_sa, stats=c_st, mf=c_mf, trans=c_pt))
