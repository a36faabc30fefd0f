"""
Primitive-family LiDAR evidence operators, NumPy float64.  TEST INFRASTRUCTURE (see oracle/__init__.py).

Reference anchors (``fl/`` = fl_ws/src/fl_slam_poc/fl_slam_poc/):
  hex_cell_3d / bin_points_3d       fl/common/ma_hex_web.py:221-303
  extract_lidar_surfels             fl/backend/operators/lidar_surfel_extraction.py:84-431
  measurement batch                 fl/backend/structures/measurement_batch.py:68-428
  tiling                            fl/common/tiling.py:32-209
  tiles / view / fuse / insert ...  fl/backend/structures/primitive_map.py:98-1484
  associate_primitives_ot           fl/backend/operators/primitive_association.py:105-553
  block_associations_for_fuse       fl/backend/operators/primitive_association.py:561-588
  visual_pose_evidence              fl/backend/operators/visual_pose_evidence.py:74-436
  map update (pipeline step 12b)    fl/backend/pipeline.py:1233-1447
JAX semantics mirrored: argsort / lax.sort are stable and lax.sort uses only its FIRST operand as key (quirk Q3);
.at[].add accumulates duplicates; boolean masks not gates.
"""

from __future__ import annotations

import math

import numpy as np

from . import lie

EPS_LIFT = 1e-9
EPS_MASS = 1e-12
N_FEAT, N_SURFEL, K_ASSOC, K_SINKHORN = 512, 1024, 8, 50
M_TILE, M_TILE_VIEW, H_TILE, K_INSERT_TILE = 50000, 1024, 2.0, 64
VMF_N_LOBES = 3
RECENCY_DECAY_LAMBDA, RECENCY_MIN_SCALE = 0.02, 0.05
FORGETTING, CULL_THRESHOLD = 0.995, 1e-4
ASSOC_BLOCK_SIZE = 256
NONFINITE_SENTINEL = 1e6
BITS_PER_AXIS, BIAS = 21, 1 << 20
SQRT3_2 = np.sqrt(np.float64(3.0)) * 0.5


# ----------------------------------------------------------------------------------------- tiling
def cells_from_xyz(XYZ, h):
    XYZ = np.asarray(XYZ, np.float64).reshape(-1, 3)
    h = max(float(h), 1e-12)
    s2 = XYZ[:, 0] * 0.5 + XYZ[:, 1] * SQRT3_2
    return (np.floor(XYZ[:, 0] / h).astype(np.int64), np.floor(s2 / h).astype(np.int64),
            np.floor(XYZ[:, 2] / h).astype(np.int64))


def pack_tile_ids(c1, c2, cz):
    m = (1 << BITS_PER_AXIS) - 1
    u1 = (np.asarray(c1, np.int64) + BIAS) & m
    u2 = (np.asarray(c2, np.int64) + BIAS) & m
    uz = (np.asarray(cz, np.int64) + BIAS) & m
    return (u1 << (2 * BITS_PER_AXIS)) | (u2 << BITS_PER_AXIS) | uz


def tile_ids_from_xyz(XYZ, h=H_TILE):
    return pack_tile_ids(*cells_from_xyz(XYZ, h))


def hex_disk_axial(radius):
    r = int(radius)
    out = []
    for q in range(-r, r + 1):
        for rr in range(max(-r, -q - r), min(r, -q + r) + 1):
            out.append((q, rr))
    out.sort()
    return out


def stencil_tile_ids(center_xyz, h=H_TILE, radius_xy=1, radius_z=0):
    """tiling.py:189-209.  Note: this host helper uses a1 = (1,0), a2 = (0.5, 0.5*sqrt(3)) via a dot product."""
    xyz = np.asarray(center_xyz, np.float64).ravel()
    hh = max(float(h), 1e-12)
    a2 = np.array([0.5, 0.5 * np.sqrt(3.0)])
    c1 = int(np.floor(float(np.array([1.0, 0.0]) @ xyz[:2]) / hh))
    c2 = int(np.floor(float(a2 @ xyz[:2]) / hh))
    cz = int(np.floor(float(xyz[2]) / hh))
    ids = []
    for dz in range(-int(radius_z), int(radius_z) + 1):
        for dq, dr in hex_disk_axial(radius_xy):
            ids.append(int(pack_tile_ids(c1 + dq, c2 + dr, cz + dz)))
    return ids


# ------------------------------------------------------------------------------ surfel extraction
def hex_cell_3d(points, h):
    h = max(float(h), 1e-12)
    p = np.asarray(points, np.float64).reshape(-1, 3)
    s2 = p[:, 0] * 0.5 + p[:, 1] * SQRT3_2
    with np.errstate(invalid="ignore"):
        return np.stack([np.floor(p[:, 0] / h).astype(np.int32), np.floor(s2 / h).astype(np.int32),
                         np.floor(p[:, 2] / h).astype(np.int32)], axis=1)


def bin_points_3d(points, point_mask, nc=(32, 32, 8), max_occ=32, voxel=0.1):
    """-> (bucket (n_cells,max_occ) int32 with -1 = empty, count (n_cells,) clipped, linear (N,) cell key)."""
    N = points.shape[0]
    n_cells = nc[0] * nc[1] * nc[2]
    cells = np.mod(hex_cell_3d(points, voxel), np.asarray(nc, np.int32)[None, :])
    linear = (cells[:, 0] * (nc[1] * nc[2]) + cells[:, 1] * nc[2] + cells[:, 2]).astype(np.int32)
    mask = np.asarray(point_mask).astype(np.int32).reshape(-1)
    linear = np.where(mask > 0, linear, 0).astype(np.int32)
    key = linear + (1 - mask) * n_cells
    order = np.argsort(key, kind="stable")
    linear_s, mask_s, idx_s = linear[order], mask[order], np.arange(N, dtype=np.int32)[order]
    pos = np.arange(N, dtype=np.int32)
    count = np.zeros(n_cells, np.int32)
    np.add.at(count, linear_s, mask_s)
    start = np.full(n_cells, N, np.int32)
    np.minimum.at(start, linear_s, pos)
    start = np.where(count > 0, start, 0)
    rank = pos - start[linear_s]
    keep = (mask_s == 1) & (rank < max_occ)
    bucket = np.full((n_cells + 1, max_occ + 1), -1, np.int32)
    bucket[np.where(keep, linear_s, n_cells), np.where(keep, rank, max_occ)] = np.where(keep, idx_s, -1)
    return bucket[:n_cells, :max_occ], np.minimum(count, max_occ), linear


def _normalize_rows(v, eps=1e-12):
    return v / (np.linalg.norm(v, axis=-1, keepdims=True) + eps)


def fit_cells(points_c, timestamps, weights, bucket, count, min_points=3, sensor_var=1e-6, wishart_nu=5.0,
              wishart_psi=0.1, kappa_scale=10.0, kappa_min=0.1, kappa_max=100.0, eig_min=1e-12, eps=1e-12):
    """Vectorised _fit_one_cell over all cells (lidar_surfel_extraction.py:84-163)."""
    idx_safe = np.maximum(bucket, 0)
    present = (bucket >= 0).astype(np.float64)
    pts = points_c[idx_safe]  # (C, occ, 3)
    w = weights[idx_safe] * present
    t = timestamps[idx_safe] * present
    w_sum = np.sum(w, axis=1) + eps
    centroid = np.sum(pts * w[:, :, None], axis=1) / w_sum[:, None]
    centered = pts - centroid[:, None, :]
    cov = np.einsum("coi,coj->cij", centered * w[:, :, None], centered) / w_sum[:, None, None]
    I = np.eye(3)
    cov = 0.5 * (cov + np.swapaxes(cov, 1, 2)) + eig_min * I
    eigvals, eigvecs = np.linalg.eigh(cov)
    normal = eigvecs[:, :, 0]
    normal = normal * np.where(normal[:, 2:3] < 0.0, -1.0, 1.0)
    normal = _normalize_rows(normal, eps)
    n = _normalize_rows(normal, eps)  # _orthonormal_basis_from_normal re-normalises
    z = np.zeros(n.shape[0])
    e1_a = np.stack([-n[:, 1], n[:, 0], z], axis=1)
    e1_b = np.stack([-n[:, 2], z, n[:, 0]], axis=1)
    e1 = _normalize_rows(np.where((np.abs(n[:, 2]) < 0.9)[:, None], e1_a, e1_b), eps)
    e2 = _normalize_rows(np.cross(n, e1), eps)
    proj1 = np.einsum("coi,ci->co", centered, e1)
    proj2 = np.einsum("coi,ci->co", centered, e2)
    var_e1 = np.sum(w * proj1 * proj1, axis=1) / w_sum + sensor_var
    var_e2 = np.sum(w * proj2 * proj2, axis=1) / w_sum + sensor_var
    sig_perp = np.maximum(eigvals[:, 0], eig_min)
    var_perp = sig_perp + sensor_var
    V = np.stack([e1, e2, normal], axis=2)
    D = np.stack([np.maximum(var_e1, eig_min), np.maximum(var_e2, eig_min), np.maximum(var_perp, eig_min)], axis=1)
    Sigma = np.einsum("cik,ck,cjk->cij", V, D, V)
    Sigma = 0.5 * (Sigma + np.swapaxes(Sigma, 1, 2)) + eig_min * I
    Lam = np.linalg.inv(Sigma + eig_min * I)
    Lam = 0.5 * (Lam + np.swapaxes(Lam, 1, 2))
    Lam_reg = Lam + (wishart_nu / max(wishart_psi, eps)) * I
    Lam_reg = 0.5 * (Lam_reg + np.swapaxes(Lam_reg, 1, 2)) + eig_min * I
    Sigma_reg = np.linalg.inv(Lam_reg)
    Sigma_reg = 0.5 * (Sigma_reg + np.swapaxes(Sigma_reg, 1, 2)) + eig_min * I
    kappa = np.clip(kappa_scale / np.sqrt(np.maximum(sig_perp, eig_min)), kappa_min, kappa_max)
    w_surfel = np.sum(w, axis=1)
    t_surfel = np.sum(t, axis=1) / w_sum
    valid = (count >= min_points) & (w_surfel > 0.0)
    return centroid, Sigma_reg, normal, kappa, w_surfel, t_surfel, valid


def empty_measurement_batch(n_feat=N_FEAT, n_surfel=N_SURFEL):
    n = n_feat + n_surfel
    return dict(Lambdas=np.zeros((n, 3, 3)), thetas=np.zeros((n, 3)), etas=np.zeros((n, VMF_N_LOBES, 3)),
                weights=np.zeros(n), sources=np.zeros(n, np.int32), source_indices=np.zeros(n, np.int32),
                valid_mask=np.zeros(n, bool), timestamps=np.zeros(n), colors=np.zeros((n, 3)), n_feat=n_feat,
                n_surfel=n_surfel, n_camera_valid=0, n_lidar_valid=0)


def batch_from_camera_splats(positions, covariances, directions, kappas, weights, timestamps, colors=None,
                             n_feat=N_FEAT, n_surfel=N_SURFEL, eps_lift=EPS_LIFT):
    """measurement_batch.py:174-259"""
    b = empty_measurement_batch(n_feat, n_surfel)
    nv = min(positions.shape[0], n_feat)
    Lam = np.linalg.inv(covariances[:nv] + eps_lift * np.eye(3)[None])
    b["Lambdas"][:nv] = Lam
    b["thetas"][:nv] = np.einsum("nij,nj->ni", Lam, positions[:nv])
    b["etas"][:nv, 0, :] = kappas[:nv, None] * directions[:nv]
    b["weights"][:nv] = weights[:nv]
    b["source_indices"][:nv] = np.arange(nv)
    b["valid_mask"][:nv] = True
    b["timestamps"][:nv] = timestamps[:nv]
    b["colors"][:nv] = np.clip(colors[:nv], 0.0, 1.0) if colors is not None else 0.5
    b["n_camera_valid"] = nv
    return b


def batch_add_lidar_surfels(batch, positions, covariances, normals, kappas, weights, timestamps, n_valid,
                            eps_lift=EPS_LIFT):
    """measurement_batch.py:272-347 (ring_indices=None, colors from normal z)."""
    b = {k: (v.copy() if isinstance(v, np.ndarray) else v) for k, v in batch.items()}
    nv = min(int(n_valid), positions.shape[0], b["n_surfel"])
    s, e = b["n_feat"], b["n_feat"] + nv
    Lam = np.linalg.inv(covariances[:nv] + eps_lift * np.eye(3)[None]) if nv else np.zeros((0, 3, 3))
    b["Lambdas"][s:e] = Lam
    b["thetas"][s:e] = np.einsum("nij,nj->ni", Lam, positions[:nv])
    et = np.zeros((nv, VMF_N_LOBES, 3))
    et[:, 0, :] = kappas[:nv, None] * normals[:nv]
    b["etas"][s:e] = et
    b["weights"][s:e] = weights[:nv]
    b["sources"][s:e] = 1
    b["source_indices"][s:e] = np.arange(nv)
    b["valid_mask"][s:e] = True
    b["timestamps"][s:e] = timestamps[:nv]
    g = 0.25 + 0.5 * (np.clip(normals[:nv, 2:3], -1.0, 1.0) + 1.0) / 2.0
    b["colors"][s:e] = np.broadcast_to(g, (nv, 3))
    b["n_lidar_valid"] = nv
    return b


def extract_lidar_surfels(points, timestamps, weights, base_batch=None, n_surfel=N_SURFEL, n_feat=N_FEAT,
                          voxel=0.1, nc=(32, 32, 8), max_occ=32, min_points=3, eps_lift=EPS_LIFT, eig_min=1e-12):
    points = np.asarray(points, np.float64).reshape(-1, 3)
    timestamps = np.asarray(timestamps, np.float64).reshape(-1)
    weights = np.asarray(weights, np.float64).reshape(-1)
    point_mask = np.all(np.abs(points) < 0.1 * NONFINITE_SENTINEL, axis=1)
    w_eff = weights * point_mask.astype(np.float64)
    w_sum = np.sum(w_eff) + eig_min
    center = np.sum(points * w_eff[:, None], axis=0) / w_sum
    pc = points - center[None, :]
    bucket, count, linear = bin_points_3d(pc, point_mask, nc, max_occ, voxel)
    cen_c, covs, normals, kappas, sw, st, valid = fit_cells(pc, timestamps, w_eff, bucket, count, min_points,
                                                            eig_min=eig_min)
    centroids = cen_c + center[None, :]
    n_cells = nc[0] * nc[1] * nc[2]
    key = np.arange(n_cells, dtype=np.int32) + (1 - valid.astype(np.int32)) * n_cells
    take = np.argsort(key, kind="stable")[:n_surfel]
    n_valid = int(np.sum(valid[take]))
    m = (np.arange(n_surfel) < n_valid).astype(np.float64)
    pos_sel = centroids[take] * m[:, None]
    cov_sel = covs[take] * m[:, None, None] + (1.0 - m)[:, None, None] * np.eye(3)[None]
    nrm_sel = normals[take] * m[:, None]
    base = empty_measurement_batch(n_feat, n_surfel) if base_batch is None else base_batch
    batch = batch_add_lidar_surfels(base, pos_sel, cov_sel, nrm_sel, kappas[take] * m, sw[take] * m, st[take] * m,
                                    n_valid, eps_lift)
    aux = dict(bucket=bucket, count=count, linear=linear, point_mask=point_mask, center=center, take=take,
               positions=pos_sel, covariances=cov_sel, normals=nrm_sel, kappas=kappas[take] * m,
               weights=sw[take] * m, timestamps=st[take] * m, cell_valid=valid)
    cert = dict(exact=False, triggers=["ma_hex3d_binning", "plane_fit_batched", "wishart_regularization"],
                ess_total=float(n_valid), support_frac=float(n_valid) / float(max(n_surfel, 1)),
                effect_name="surfel_extraction", effect_predicted=float(n_valid))
    return batch, aux, cert


def batch_mean_positions(b, eps_lift=EPS_LIFT):
    return np.linalg.solve(b["Lambdas"] + eps_lift * np.eye(3)[None], b["thetas"][:, :, None])[:, :, 0]


def batch_mean_directions(b, eps_mass=EPS_MASS):
    es = np.sum(b["etas"], axis=1)
    return es / (np.linalg.norm(es, axis=1, keepdims=True) + eps_mass)


def batch_kappas(b):
    return np.linalg.norm(np.sum(b["etas"], axis=1), axis=1)


# ----------------------------------------------------------------------------------------- the map
TILE_FIELDS = ("Lambdas", "thetas", "etas", "weights", "timestamps", "created_timestamps", "last_supported_scan_seq",
               "last_update_scan_seq", "primitive_ids", "valid_mask", "colors", "cam_mass", "lidar_mass",
               "rgb_cam_accum", "rgb_cam_denom", "rgb")


def create_empty_tile(tile_id, m_tile=M_TILE):
    return dict(tile_id=int(tile_id), Lambdas=np.zeros((m_tile, 3, 3)), thetas=np.zeros((m_tile, 3)),
                etas=np.zeros((m_tile, VMF_N_LOBES, 3)), weights=np.zeros(m_tile), timestamps=np.zeros(m_tile),
                created_timestamps=np.zeros(m_tile), last_supported_scan_seq=np.zeros(m_tile, np.int64),
                last_update_scan_seq=np.zeros(m_tile, np.int64), primitive_ids=np.zeros(m_tile, np.int64),
                valid_mask=np.zeros(m_tile, bool), colors=np.zeros((m_tile, 3)), cam_mass=np.zeros(m_tile),
                lidar_mass=np.zeros(m_tile), rgb_cam_accum=np.zeros((m_tile, 3)), rgb_cam_denom=np.zeros(m_tile),
                rgb=np.full((m_tile, 3), 0.5), next_local_id=0, count=0)


def create_empty_atlas(m_tile=M_TILE):
    return dict(tiles={}, next_global_id=0, total_count=0, m_tile=int(m_tile))


def copy_atlas(a):
    return dict(tiles={k: {f: (v.copy() if isinstance(v, np.ndarray) else v) for f, v in t.items()}
                       for k, t in a["tiles"].items()},
                next_global_id=a["next_global_id"], total_count=a["total_count"], m_tile=a["m_tile"])


def select_topk_slots(weights, valid_mask, k):
    """primitive_map.py:303-322: stable sort on -score ONLY (primitive_id is not a key: quirk Q3)."""
    score = np.where(valid_mask, weights, -1e30)
    return np.argsort(-score, kind="stable")[:k].astype(np.int32)


def select_lowest_retention_slots(weights, valid_mask, last_supported, scan_seq, lam, k):
    """primitive_map.py:325-353"""
    dt = np.maximum(0, np.int64(scan_seq) - last_supported.astype(np.int64))
    retention = weights * np.exp(-lam * dt.astype(np.float64))
    key = np.where(valid_mask, retention, -np.inf)
    return np.argsort(key, kind="stable")[:k].astype(np.int32)


def recency_inflate(atlas, tile_ids, scan_seq, lam=RECENCY_DECAY_LAMBDA, min_scale=RECENCY_MIN_SCALE):
    """primitive_map.py:1400-1484.  In place on a copy; returns (atlas, stats dict)."""
    atlas = copy_atlas(atlas)
    tot_down = tot_tr = n_valid_total = 0.0
    for tid in tile_ids:
        t = atlas["tiles"].get(int(tid))
        if t is None:
            continue
        valid = t["valid_mask"].astype(np.float64)
        dt = np.maximum(0, np.int64(scan_seq) - t["last_supported_scan_seq"])
        decay = np.clip(np.exp(-lam * dt.astype(np.float64)), float(min_scale), 1.0)
        decay = np.where(t["valid_mask"], decay, 1.0)
        t["Lambdas"] = t["Lambdas"] * decay[:, None, None]
        t["thetas"] = t["thetas"] * decay[:, None]
        n_valid_total += float(np.sum(valid))
        tot_down += float(np.sum((1.0 - decay) * valid))
        tot_tr += float(np.sum(((1.0 / decay) - 1.0) * valid))
    return atlas, dict(staleness_inflation_strength=tot_down / max(n_valid_total, 1.0),
                       staleness_cov_inflation_trace=tot_tr, stale_precision_downscale_total=tot_down,
                       n_valid_total=n_valid_total)


def extract_atlas_map_view(atlas, tile_ids, m_tile_view=M_TILE_VIEW, eps_lift=EPS_LIFT, eps_mass=EPS_MASS):
    """primitive_map.py:356-450, :474-498"""
    k = int(m_tile_view)
    parts = {n: [] for n in ("slots", "tids", "valid", "Lam", "th", "et", "w", "ids", "last", "col")}
    for tid in tile_ids:
        t = atlas["tiles"].get(int(tid)) or create_empty_tile(int(tid), atlas["m_tile"])
        s = select_topk_slots(t["weights"], t["valid_mask"], k)
        parts["slots"].append(s)
        parts["tids"].append(np.full(k, int(tid), np.int64))
        parts["valid"].append(t["valid_mask"][s])
        parts["Lam"].append(t["Lambdas"][s]); parts["th"].append(t["thetas"][s]); parts["et"].append(t["etas"][s])
        parts["w"].append(t["weights"][s]); parts["ids"].append(t["primitive_ids"][s])
        parts["last"].append(t["last_supported_scan_seq"][s]); parts["col"].append(t["rgb"][s])
    c = {n: np.concatenate(v, axis=0) for n, v in parts.items()}
    Lreg = c["Lam"] + eps_lift * np.eye(3)[None]
    positions = np.linalg.solve(Lreg, c["th"][:, :, None])[:, :, 0]
    covariances = np.linalg.inv(Lreg)
    eta_sum = np.sum(c["et"], axis=1)
    kappas = np.linalg.norm(eta_sum, axis=1)
    directions = eta_sum / (kappas[:, None] + eps_mass)
    return dict(candidate_tile_ids=c["tids"], candidate_slots=c["slots"], valid_mask=c["valid"].astype(bool),
                tile_ids=np.asarray(tile_ids, np.int64), m_tile_view=k, positions=positions, covariances=covariances,
                directions=directions, kappas=kappas, weights=c["w"], primitive_ids=c["ids"],
                last_supported_scan_seq=c["last"], etas=c["et"], colors=c["col"])


# ------------------------------------------------------------------------------------ association
def A_vmf(k, eps=1e-12):
    k = np.maximum(np.asarray(k, np.float64), eps)
    with np.errstate(over="ignore"):
        ls = np.where(k > 20.0, k - np.log(2.0), np.where(k >= 1e-2, np.log(np.sinh(np.minimum(k, 700.0))),
                                                         np.log(k + (k**3) / 6.0)))
    return np.log(4.0 * np.pi) + ls - np.log(k)


def sparse_cost(meas_pos, meas_dir, meas_kappa, map_pos, map_dir, map_kappa, cand, beta=0.5, eig_min=1e-12):
    """primitive_association.py:152-197"""
    mp, md, mk = map_pos[cand], map_dir[cand], map_kappa[cand]
    diff = meas_pos[:, None, :] - mp
    d_pos = np.sum(diff * diff, axis=-1)
    km = 0.5 * np.linalg.norm(meas_kappa[:, None, None] * meas_dir[:, None, :] + mk[:, :, None] * md, axis=-1)
    A_km = A_vmf(np.maximum(km, eig_min), eig_min)
    A_k1 = A_vmf(np.maximum(meas_kappa[:, None], eig_min), eig_min)
    A_k2 = A_vmf(np.maximum(mk, eig_min), eig_min)
    d_dir = np.maximum(0.0, 1.0 - np.exp(A_km - 0.5 * (A_k1 + A_k2)))
    d_dir = np.where((meas_kappa[:, None] > 0.0) & (mk > 0.0), d_dir, 0.0)
    return d_pos + float(beta) * d_dir


def sinkhorn_unbalanced(Cm, a, b, epsilon, tau_a, tau_b, K):
    """primitive_association.py:105-138: one (M,) vector v shared by all rows."""
    eps = max(float(epsilon), 1e-12)
    Kmat = np.exp(-Cm / eps)
    u = np.ones(Cm.shape[0])
    v = np.ones(Cm.shape[1])
    ua, vb = 1.0 / (1.0 + tau_a / eps), 1.0 / (1.0 + tau_b / eps)
    for _ in range(int(K)):
        u = (a / (Kmat @ v + 1e-12)) ** ua
        v = (b / (Kmat.T @ u + 1e-12)) ** vb
    return u[:, None] * Kmat * v[None, :]


def associate_primitives_ot(batch, view, scan_seq=0, k_assoc=K_ASSOC, k_sinkhorn=K_SINKHORN, beta=0.5, epsilon=0.1,
                            tau_a=0.5, tau_b=0.5, eps_mass=EPS_MASS, h_tile=H_TILE, lam=RECENCY_DECAY_LAMBDA,
                            eps_lift=EPS_LIFT, chunk=128, a_policy="uniform"):
    """primitive_association.py:239-553.  a_policy: "uniform" | "weight_proportional" (:412-424)."""
    N = batch["n_feat"] + batch["n_surfel"]
    n_valid = batch["n_camera_valid"] + batch["n_lidar_valid"]
    M_valid = int(np.sum(view["valid_mask"]))
    if n_valid == 0 or M_valid == 0:
        z = np.zeros((N, k_assoc))
        return dict(responsibilities=z, candidate_pool_indices=z.astype(np.int32), candidate_tile_ids=z.astype(np.int64),
                    candidate_slots=z.astype(np.int64), row_masses=np.zeros(N), cost_matrix=z.copy()), dict(exact=True, triggers=[])
    mpos = batch_mean_positions(batch, eps_lift)
    mdir = batch_mean_directions(batch, eps_mass)
    mkap = batch_kappas(batch)
    vm = batch["valid_mask"].astype(np.float64)
    c1, c2, cz = cells_from_xyz(mpos, h_tile)
    disk = hex_disk_axial(1)
    st_ids = np.stack([pack_tile_ids(c1 + dq, c2 + dr, cz) for dq, dr in disk], axis=1)  # (N, 7)
    tiles_pool = view["tile_ids"]
    mtv = int(view["m_tile_view"])
    eq = st_ids[:, :, None] == tiles_pool[None, None, :]
    has = np.any(eq, axis=2)
    tidx = np.where(has, np.argmax(eq, axis=2), 0)
    P = st_ids.shape[1] * mtv
    cand = np.zeros((N, k_assoc), np.int32)
    off = np.arange(mtv, dtype=np.int32)
    for s in range(0, N, chunk):
        e = min(N, s + chunk)
        pool_idx = (tidx[s:e, :, None].astype(np.int32) * mtv + off[None, None, :]).reshape(e - s, P)
        cp = sparse_cost(mpos[s:e], mdir[s:e], mkap[s:e], view["positions"], view["directions"], view["kappas"],
                         pool_idx, beta)
        pv = view["valid_mask"][pool_idx] & np.repeat(has[s:e], mtv, axis=1)
        cp = np.where(pv, cp, 1e12)
        order = np.argsort(cp, axis=1, kind="stable")[:, :k_assoc]   # key = cost only (quirk Q3)
        cand[s:e] = np.take_along_axis(pool_idx, order, axis=1)
    cand = np.where(vm[:, None] > 0.0, cand, 0).astype(np.int32)
    cslots = view["candidate_slots"][cand].astype(np.int64)
    ctids = view["candidate_tile_ids"][cand].astype(np.int64)
    Cm = sparse_cost(mpos, mdir, mkap, view["positions"], view["directions"], view["kappas"], cand, beta)
    cdt = np.maximum(0, np.int64(scan_seq) - view["last_supported_scan_seq"][cand]).astype(np.float64)
    Cm = Cm + float(epsilon) * float(lam) * cdt
    Cm = Cm - np.min(Cm, axis=1, keepdims=True)
    if a_policy == "uniform":
        sum_a = max(np.sum(vm), eps_mass)
        a = vm / sum_a
    elif a_policy == "weight_proportional":
        weighted = vm * batch["weights"].astype(np.float64)
        sum_a = max(np.sum(weighted), eps_mass)
        a = weighted / sum_a
    else:
        raise ValueError(f"Unsupported measurement mass policy: {a_policy}. Only UNIFORM and WEIGHT_PROPORTIONAL are implemented.")
    b = np.ones(k_assoc) / float(k_assoc)
    pi = sinkhorn_unbalanced(Cm, a, b, epsilon, tau_a, tau_b, k_sinkhorn)
    row = np.sum(pi, axis=1)
    resp = pi * (vm[:, None] > 0.0)
    rd = np.exp(-lam * cdt)
    brow = rd / np.maximum(np.sum(rd, axis=1, keepdims=True), eps_mass)
    bs = np.sort(brow.reshape(-1))
    tm = float(np.sum(pi))
    a_s = np.sort(a)
    cert = dict(exact=False, triggers=["sinkhorn_fixed_iter", "sinkhorn_unbalanced_kl_relax"],
                ess_total=float(np.sum(row) ** 2 / (np.sum(row**2) + eps_mass)),
                support_frac=float(int(np.sum(a > eps_mass))) / float(max(N, 1)),
                mass_epsilon_ratio=float(eps_mass) / (tm + eps_mass),
                marginal_defect_a=float(np.linalg.norm(row - a)), marginal_defect_b=float(np.linalg.norm(np.sum(pi, 0) - b)),
                transport_mass_total=tm, sum_a=float(sum_a), sum_b=float(np.sum(b)), sum_m=float(np.sum(row)),
                sum_novel=float(np.sum(np.maximum(a - row, 0.0))), p95_a=float(a_s[min(int(0.95 * N), N - 1)]),
                p95_b=float(b[min(int(0.95 * k_assoc), k_assoc - 1)]), nonzero_a=int(np.sum(a > eps_mass)),
                nonzero_b=int(np.sum(b > eps_mass)), b_recency_p95=float(bs[min(int(0.95 * bs.shape[0]), bs.shape[0] - 1)]),
                total_cost=float(np.sum(pi * Cm)))
    return dict(responsibilities=resp, candidate_pool_indices=cand, candidate_tile_ids=ctids, candidate_slots=cslots,
                row_masses=row, cost_matrix=Cm), cert


# ---------------------------------------------------------------------------------- pose evidence
def visual_pose_evidence(assoc, batch, view, pose6, eps_lift=EPS_LIFT, eps_mass=EPS_MASS):
    """visual_pose_evidence.py:261-436 with z_lin_pose = pose6."""
    N_meas = batch["n_camera_valid"] + batch["n_lidar_valid"]
    if N_meas == 0 or int(np.sum(view["valid_mask"])) == 0:
        return dict(L_pose=eps_lift * np.eye(22), h_pose=np.zeros(22), L_trans=np.zeros((3, 3)), h_trans=np.zeros(3),
                    L_rot=np.zeros((3, 3)), h_rot=np.zeros(3), total_weighted_cost=0.0, n_associations=0,
                    mean_transported_mass=0.0), dict(exact=True, triggers=[])
    pose6 = np.asarray(pose6, np.float64).ravel()[:6]
    t_pred, R_pred = pose6[:3], lie.so3_exp(pose6[3:6])
    vi = np.where(batch["valid_mask"])[0][:N_meas]
    mpos = batch_mean_positions(batch, eps_lift)[vi]
    mdir = batch_mean_directions(batch, eps_mass)[vi]
    mkap = batch_kappas(batch)[vi]
    Lam = (batch["Lambdas"] + eps_lift * np.eye(3)[None])[vi]
    pi = assoc["responsibilities"][vi]
    cand = assoc["candidate_pool_indices"][vi].astype(np.int32)
    row = assoc["row_masses"][vi]
    mw = mpos @ R_pred.T
    mp_all = view["positions"][cand]
    resid = mp_all - mw[:, None, :] - t_pred[None, None, :]
    L_t = np.einsum("n,nij->ij", np.sum(pi, axis=1), Lam)
    target = mp_all - mw[:, None, :]
    h_t = np.einsum("nij,nj->i", Lam, np.einsum("nk,nkj->nj", pi, target))
    Lr = np.einsum("nij,nkj->nki", Lam, resid)
    tcost = float(np.sum(pi * np.einsum("nki,nki->nk", resid, Lr)))
    L_t = L_t + eps_lift * np.eye(3)
    md_all = view["directions"][cand]
    mk_all = view["kappas"][cand]
    wts = pi * np.sqrt(mkap[:, None] * mk_all + 1e-12)
    S = np.einsum("nk,nki,nj->ij", wts, md_all, mdir)
    rcost = float(np.sum(wts * (1.0 - np.einsum("ni,nki->nk", mdir @ R_pred.T, md_all))))
    U, s, Vt = np.linalg.svd(S)
    L_r = np.diag(s + eps_lift)
    R_sc = U @ Vt
    if np.linalg.det(R_sc) < 0:
        R_sc = U @ np.diag([1.0, 1.0, -1.0]) @ Vt
    dr = lie.so3_log(R_sc @ R_pred.T)
    h_r = L_r @ dr
    L = eps_lift * np.eye(22)
    h = np.zeros(22)
    L[0:3, 0:3] = L_t; h[0:3] = h_t; L[3:6, 3:6] = L_r; h[3:6] = h_r
    res = dict(L_pose=L, h_pose=h, L_trans=L_t, h_trans=h_t, L_rot=L_r, h_rot=h_r, total_weighted_cost=tcost + rcost,
               n_associations=int(pi.shape[0] * pi.shape[1]), mean_transported_mass=float(np.mean(row)), R_scatter=R_sc,
               S=S, svd_s=s, delta_rot=dr)
    cert = dict(exact=False, triggers=["linearization", "ot_soft_correspondence"], frobenius_applied=True,
                ess_total=float(np.sum(row)), support_frac=float(pi.shape[0]) / float(max(N_meas, 1)),
                lift_strength=eps_lift)
    return res, cert


# ------------------------------------------------------------------------------------- map update
def block_associations_for_fuse(assoc, valid_mask, block=ASSOC_BLOCK_SIZE):
    N, K = assoc["responsibilities"].shape
    nb = (N + block - 1) // block
    meas_idx = np.arange(nb * block, dtype=np.int32).reshape(nb, block)
    clipped = np.minimum(meas_idx, N - 1)
    valid_rows = (meas_idx < N) & np.asarray(valid_mask, bool)[clipped]
    return (clipped, assoc["candidate_tile_ids"][clipped], assoc["candidate_slots"][clipped],
            assoc["responsibilities"][clipped] * valid_rows[:, :, None], valid_rows)


def to_world(Lam_b, th_b, eta_b, R, t, eps_lift=EPS_LIFT):
    """pipeline.py:1248-1256 (batched)."""
    Lw = np.einsum("ij,njk,lk->nil", R, Lam_b, R)
    mu_b = np.linalg.solve(Lam_b + eps_lift * np.eye(3)[None], th_b[:, :, None])[:, :, 0]
    mu_w = mu_b @ R.T + t[None, :]
    return Lw, np.einsum("nij,nj->ni", Lw, mu_w), np.einsum("ij,nbj->nbi", R, eta_b)


def primitive_map_fuse(atlas, tile_id, target_slots, Lam, th, eta, w_meas, resp, timestamp, scan_seq, valid_mask,
                       colors, sources, eps_mass=EPS_MASS):
    """primitive_map.py:992-1163 (one call = one (block, tile)); mutates atlas in place."""
    tid = int(tile_id)
    t = atlas["tiles"].get(tid)
    if t is None:
        t = create_empty_tile(tid, atlas["m_tile"])
    if target_slots.shape[0] == 0:
        return 0
    M = t["Lambdas"].shape[0]
    r = (resp * valid_mask.astype(np.float64)).astype(np.float64) if valid_mask is not None else np.asarray(resp, np.float64)
    dL = np.zeros((M, 3, 3)); dth = np.zeros((M, 3)); det = np.zeros((M, VMF_N_LOBES, 3)); dw = np.zeros(M)
    drs = np.zeros(M); dcam = np.zeros(M); dlid = np.zeros(M); dacc = np.zeros((M, 3)); dden = np.zeros(M)
    idx = target_slots
    np.add.at(dL, idx, r[:, None, None] * Lam)
    np.add.at(dth, idx, r[:, None] * th)
    np.add.at(det, idx, r[:, None, None] * eta)
    np.add.at(dw, idx, r * w_meas)
    np.add.at(drs, idx, r)
    if sources is not None:                       # :1084-1095: masses need sources, colour accumulators need both
        w_cam = r * w_meas * (sources == 0).astype(np.float64)
        w_lid = r * w_meas * (sources == 1).astype(np.float64)
        np.add.at(dcam, idx, w_cam); np.add.at(dlid, idx, w_lid)
        if colors is not None:
            cc = np.clip(colors, 0.0, 1.0)
            np.add.at(dacc, idx, cc * w_cam[:, None]); np.add.at(dden, idx, w_cam)
    t["cam_mass"] = t["cam_mass"] + dcam
    t["lidar_mass"] = t["lidar_mass"] + dlid
    t["rgb_cam_accum"] = t["rgb_cam_accum"] + dacc
    t["rgb_cam_denom"] = t["rgb_cam_denom"] + dden
    est = np.clip(t["rgb_cam_accum"] / np.maximum(t["rgb_cam_denom"][:, None], eps_mass), 0.0, 1.0)
    t["rgb"] = np.where((t["cam_mass"] > 0.0)[:, None], est, 0.5)
    t["colors"] = t["rgb"].copy()
    t["Lambdas"] = t["Lambdas"] + dL
    t["thetas"] = t["thetas"] + dth
    t["etas"] = t["etas"] + det
    t["weights"] = t["weights"] + dw
    uniq = np.unique(target_slots)
    t["timestamps"] = t["timestamps"].copy()
    t["timestamps"][uniq] = timestamp            # UNMASKED (quirk Q7)
    upd = drs > 0.0
    t["last_supported_scan_seq"] = np.where(upd, np.int64(scan_seq), t["last_supported_scan_seq"])
    t["last_update_scan_seq"] = np.where(upd, np.int64(scan_seq), t["last_update_scan_seq"])
    atlas["tiles"][tid] = t
    return int(uniq.shape[0])


def primitive_map_insert_masked(atlas, tile_id, Lam, th, eta, w_new, timestamp, valid_new, scan_seq, lam, colors, sources):
    """primitive_map.py:807-981; mutates atlas in place; returns (n_inserted, new_ids, target_slots)."""
    tid = int(tile_id)
    t = atlas["tiles"].get(tid)
    if t is None:
        t = create_empty_tile(tid, atlas["m_tile"])
    K = Lam.shape[0]
    slots = select_lowest_retention_slots(t["weights"], t["valid_mask"], t["last_supported_scan_seq"], scan_seq, lam, K)
    do = np.asarray(valid_new, bool).reshape(-1)
    n_ins = int(np.sum(do))
    prefix = np.cumsum(do.astype(np.int64)) - 1
    new_ids = np.where(do, np.int64(atlas["next_global_id"]) + prefix, np.int64(-1))
    if colors is None:                            # :884-893 defaults: black, all lidar
        colors = np.zeros((K, 3))
    is_cam = (sources == 0).astype(np.float64) if sources is not None else np.zeros(K)
    is_lid = (sources == 1).astype(np.float64) if sources is not None else np.ones(K)
    cam_new, lid_new = w_new * is_cam, w_new * is_lid
    rgb_new = np.where((cam_new > 0.0)[:, None], np.clip(colors, 0.0, 1.0), 0.5)

    def put(name, new, expand):
        cur = t[name]
        sel = do.reshape((-1,) + (1,) * expand)
        out = cur.copy()
        out[slots] = np.where(sel, new, cur[slots])
        t[name] = out

    put("Lambdas", Lam, 2); put("thetas", th, 1); put("etas", eta, 2); put("weights", w_new, 0)
    put("timestamps", float(timestamp), 0); put("created_timestamps", float(timestamp), 0)
    put("last_supported_scan_seq", np.int64(scan_seq), 0); put("last_update_scan_seq", np.int64(scan_seq), 0)
    put("primitive_ids", new_ids, 0)
    vm = t["valid_mask"].copy(); vm[slots] = t["valid_mask"][slots] | do; t["valid_mask"] = vm
    put("colors", rgb_new, 1); put("cam_mass", cam_new, 0); put("lidar_mass", lid_new, 0)
    put("rgb_cam_accum", colors * cam_new[:, None], 1); put("rgb_cam_denom", cam_new, 0); put("rgb", rgb_new, 1)
    t["count"] = int(np.sum(t["valid_mask"]))
    atlas["tiles"][tid] = t
    atlas["next_global_id"] = int(atlas["next_global_id"] + n_ins)
    atlas["total_count"] = int(atlas["total_count"] + n_ins)
    return n_ins, new_ids, slots


def primitive_map_cull(atlas, tile_id, thr=CULL_THRESHOLD, max_primitives=None):
    """primitive_map.py:1175-1304 (max_primitives: :1226-1232, the threshold becomes the weight of descending rank
    max_primitives of weights * valid when more than max_primitives primitives would survive)."""
    t = atlas["tiles"].get(int(tile_id))
    if t is None or t["count"] == 0:
        return 0, 0.0
    below = t["valid_mask"] & (t["weights"] < thr)
    n_keep = t["count"] - int(np.sum(below))
    if max_primitives is not None and n_keep > max_primitives:
        sw = np.sort(t["weights"] * t["valid_mask"].astype(np.float64))[::-1]
        if max_primitives < len(sw):
            below = t["valid_mask"] & (t["weights"] < float(sw[max_primitives]))
    n = int(np.sum(below))
    if n == 0:
        return 0, 0.0
    mass = float(np.sum(t["weights"] * below.astype(np.float64)))
    t["valid_mask"] = t["valid_mask"] & ~below
    t["count"] = t["count"] - n
    atlas["total_count"] -= n
    return n, mass


def primitive_map_forget(atlas, tile_id, gamma=FORGETTING):
    t = atlas["tiles"].get(int(tile_id))
    if t is not None:
        t["weights"] = float(gamma) * t["weights"]


def map_update(atlas, batch, assoc, active_tile_ids, pose6, scan_seq, timestamp, k_insert=K_INSERT_TILE,
               lam=RECENCY_DECAY_LAMBDA, eps_lift=EPS_LIFT, eps_mass=EPS_MASS, h_tile=H_TILE, cull_thr=CULL_THRESHOLD,
               gamma=FORGETTING):
    """pipeline.py:1233-1447 (merge_reduce is a no-op above 2048 slots per tile and is not restated here)."""
    atlas = copy_atlas(atlas)
    pose6 = np.asarray(pose6, np.float64)
    R, t = lie.so3_exp(pose6[3:6]), pose6[:3]
    stats = dict(fused_count=0, fused_mass_total=0.0, insert_count_total=0, insert_mass_total=0.0, insert_mass_p95=0.0,
                 evicted_count=0, evicted_mass_total=0.0, new_ids=[], insert_slots=[])
    meas_idx_b, tile_b, slot_b, resp_b, vrows_b = block_associations_for_fuse(assoc, batch["valid_mask"])
    K = slot_b.shape[2]
    for b in range(meas_idx_b.shape[0]):
        mi = meas_idx_b[b]
        tflat = slot_b[b].reshape(-1).astype(np.int32)
        tids = tile_b[b].reshape(-1).astype(np.int64)
        rf = resp_b[b].reshape(-1)
        vf = np.repeat(vrows_b[b], K)
        Lw, thw, etw = to_world(np.repeat(batch["Lambdas"][mi], K, 0), np.repeat(batch["thetas"][mi], K, 0),
                                np.repeat(batch["etas"][mi], K, 0), R, t, eps_lift)
        wm = np.repeat(batch["weights"][mi], K, 0)
        cm = np.repeat(batch["colors"][mi], K, 0)
        sm = np.repeat(batch["sources"][mi], K, 0)
        for tid in active_tile_ids:
            vt = vf & (tids == int(tid))
            stats["fused_mass_total"] += float(np.sum(wm * rf * vt.astype(np.float64)))
            stats["fused_count"] += primitive_map_fuse(atlas, tid, tflat, Lw, thw, etw, wm, rf, timestamp, scan_seq, vt,
                                                       cm, sm, eps_mass)
    a = batch["valid_mask"].astype(np.float64)
    a = a / max(np.sum(a), eps_mass)
    novelty = np.maximum(a - assoc["row_masses"], 0.0)
    score = novelty * batch["weights"] - (1.0 - batch["valid_mask"].astype(np.float64)) * 1e6
    mu_b = np.linalg.solve(batch["Lambdas"] + eps_lift * np.eye(3)[None], batch["thetas"][:, :, None])[:, :, 0]
    mu_w = mu_b @ R.T + t[None, :]
    mt = tile_ids_from_xyz(mu_w, h_tile)
    for tid in active_tile_ids:
        in_tile = mt == np.int64(tid)
        st = np.where(in_tile, score, -1e30)
        ins = np.argsort(-st, kind="stable")[:k_insert].astype(np.int32)
        vnew = in_tile[ins] & (st[ins] > -1e20)
        vnew = vnew if np.any(vnew) else np.ones_like(vnew)
        w_ins = np.where(in_tile[ins], novelty[ins] * batch["weights"][ins], 0.0)
        stats["insert_mass_total"] += float(np.sum(w_ins))
        ws = np.sort(w_ins)
        stats["insert_mass_p95"] = max(stats["insert_mass_p95"], float(ws[min(int(0.95 * ws.shape[0]), ws.shape[0] - 1)]))
        Lw, thw, etw = to_world(batch["Lambdas"][ins], batch["thetas"][ins], batch["etas"][ins], R, t, eps_lift)
        n_ins, ids, slots = primitive_map_insert_masked(atlas, tid, Lw, thw, etw, w_ins, timestamp, vnew, scan_seq, lam,
                                                        batch["colors"][ins], batch["sources"][ins])
        stats["insert_count_total"] += n_ins
        stats["new_ids"].append(ids); stats["insert_slots"].append(slots)
    for tid in active_tile_ids:
        n, m = primitive_map_cull(atlas, tid, cull_thr)
        stats["evicted_count"] += n; stats["evicted_mass_total"] += m
        primitive_map_forget(atlas, tid, gamma)
    return atlas, stats
