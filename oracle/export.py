"""
Map export: per-tile primitive view -> renderable batch -> /gc/map/points PointCloud2, NumPy float64.
TEST INFRASTRUCTURE (see oracle/__init__.py).

Restates:
  extract_primitive_map_view / _extract_primitive_map_view_core   fl/backend/structures/primitive_map.py:474-576
  renderable_batch_from_view                                       fl/backend/structures/primitive_map.py:580-616
  PrimitiveMapPublisher.publish (concatenate, recency order)       fl/backend/map_publisher.py:131-258
  _build_pointcloud2_from_view                                     fl/backend/map_publisher.py:44-90
Pinned to the reference's own sources by tests/golden/make_golden_export.py (tests/golden/export_*.npz).
max_primitives (the optional per-tile down-selection) is restated too; the device path builds the publisher's default
(None).
"""
from __future__ import annotations

import numpy as np


def extract_primitive_map_view(tile: dict, max_primitives=None, eps_lift=1e-9, eps_mass=1e-12):
    """tile: dict of the PrimitiveMapTile arrays (valid_mask, Lambdas, thetas, etas, weights, primitive_ids, rgb, ...)."""
    valid = np.asarray(tile["valid_mask"]).astype(bool)
    slots = np.nonzero(valid)[0]
    if max_primitives is not None and slots.shape[0] > max_primitives:
        top = np.argsort(-np.asarray(tile["weights"])[slots], kind="stable")[:max_primitives]
        slots = slots[top]
    Lam = np.asarray(tile["Lambdas"], dtype=np.float64)[slots] + eps_lift * np.eye(3)[None]
    th = np.asarray(tile["thetas"], dtype=np.float64)[slots]
    et = np.asarray(tile["etas"], dtype=np.float64)[slots]
    n = slots.shape[0]
    pos = np.linalg.solve(Lam, th[..., None])[..., 0] if n else np.zeros((0, 3))
    cov = np.linalg.inv(Lam) if n else np.zeros((0, 3, 3))
    eta_sum = np.sum(et, axis=1)
    kap = np.linalg.norm(eta_sum, axis=1)
    dirs = eta_sum / (kap[:, None] + eps_mass)
    return dict(slot_indices=slots.astype(np.int32), positions=pos, covariances=cov, directions=dirs, kappas=kap,
                weights=np.asarray(tile["weights"], dtype=np.float64)[slots],
                primitive_ids=np.asarray(tile["primitive_ids"], dtype=np.int64)[slots],
                colors=np.asarray(tile["rgb"], dtype=np.float64)[slots], etas=et,
                last_supported_scan_seq=np.asarray(tile["last_supported_scan_seq"], dtype=np.int64)[slots])


def pointcloud2_xyz_intensity(positions, weights):
    """map_publisher.py:44-90: 16-byte records x, y, z, intensity (float32 little endian), intensity = clip(w, 0, 1e6)."""
    n = positions.shape[0]
    arr = np.zeros((n,), dtype=np.dtype([("x", "<f4"), ("y", "<f4"), ("z", "<f4"), ("intensity", "<f4")]))
    if n:
        arr["x"] = positions[:, 0].astype(np.float32)
        arr["y"] = positions[:, 1].astype(np.float32)
        arr["z"] = positions[:, 2].astype(np.float32)
        arr["intensity"] = np.clip(weights.astype(np.float32), 0.0, 1e6)
    return np.frombuffer(arr.tobytes(), dtype=np.uint8).copy()


def export_map_points(atlas: dict, tile_ids=None, max_primitives=None, eps_lift=1e-9, eps_mass=1e-12):
    """map_publisher.py:131-258: views of the selected tiles (default all, sorted ids), newest first, ties by id."""
    if tile_ids is None:
        tile_ids = sorted(atlas["tiles"].keys())
    views = [extract_primitive_map_view(atlas["tiles"][int(t)], max_primitives, eps_lift, eps_mass)
             for t in tile_ids if int(t) in atlas["tiles"]]
    cat = lambda k, shape: (np.concatenate([v[k] for v in views], axis=0) if views else np.zeros(shape))
    pos, w, col = cat("positions", (0, 3)), cat("weights", (0,)), cat("colors", (0, 3))
    rec, pid = cat("last_supported_scan_seq", (0,)).astype(np.int64), cat("primitive_ids", (0,)).astype(np.int64)
    cov, eta = cat("covariances", (0, 3, 3)), cat("etas", (0, 3, 3))
    order = np.lexsort((pid, -rec))
    pos, w, col, rec, pid, cov, eta = pos[order], w[order], col[order], rec[order], pid[order], cov[order], eta[order]
    lam = np.linalg.inv(cov + eps_lift * np.eye(3)[None]) if cov.size else np.zeros((0, 3, 3))
    return dict(mu_world=pos, Sigma_world=cov, Lambda_world=lam, eta=eta, mass=w, color=col, primitive_ids=pid,
                last_supported_scan_seq=rec, cloud=pointcloud2_xyz_intensity(pos, w), order=order)
