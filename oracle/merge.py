"""
primitive_map_merge_reduce, NumPy float64.  TEST INFRASTRUCTURE (see oracle/__init__.py).

Restates fl/backend/structures/primitive_map.py:1501-2031: all-pairs Bhattacharyya distance of the tile's Gaussians,
greedy disjoint selection of at most max_pairs pairs in stable ascending order of distance (< merge_threshold), moment-
matched merge into the first slot of each pair.  No-op above max_tile_size slots (the reference's budget cap; M_TILE is
50,000 in the reference configuration, so the operator only acts on small tiles).
Pinned by tests/golden/make_golden_merge.py and by the reference's own known-answer test
(test/test_primitive_map_merge_reduce.py:77-99: three primitives, the close pair merges into weight 2).
"""
from __future__ import annotations

import numpy as np


def merge_reduce_tile(tile: dict, merge_threshold=0.1, max_pairs=4, max_tile_size=2048, eps_psd=1e-12, eps_lift=1e-9):
    """-> (new tile dict (copy), n_merged, status) with status in {"merged", "noop", "budget_cap"}."""
    t = {k: (np.array(v, copy=True) if isinstance(v, np.ndarray) else v) for k, v in tile.items()}
    w = np.asarray(t["weights"], dtype=np.float64).reshape(-1)
    valid = np.asarray(t["valid_mask"]).astype(bool).reshape(-1)
    M = w.shape[0]
    if M < 2 or int(valid.sum()) < 2 or int(max_pairs) <= 0:
        return t, 0, "noop"
    if max_tile_size > 0 and M > max_tile_size:
        return t, 0, "budget_cap"
    Lam_reg = np.asarray(t["Lambdas"], dtype=np.float64) + eps_lift * np.eye(3)[None]
    mu = np.linalg.solve(Lam_reg, np.asarray(t["thetas"], dtype=np.float64)[..., None])[..., 0]
    Sigma = np.linalg.inv(Lam_reg)
    det_S = np.linalg.det(Sigma)
    i_idx, j_idx = np.triu_indices(M, k=1)
    S = 0.5 * (Sigma[i_idx] + Sigma[j_idx])
    S_inv = np.linalg.inv(S + eps_lift * np.eye(3)[None])
    dmu = (mu[i_idx] - mu[j_idx])[:, :, None]
    quad = 0.125 * np.squeeze(np.matmul(np.matmul(dmu.transpose(0, 2, 1), S_inv), dmu), axis=(1, 2))
    with np.errstate(invalid="ignore", divide="ignore"):
        log_term = 0.5 * np.log(np.linalg.det(S) / np.sqrt(det_S[i_idx] * det_S[j_idx] + 1e-24))
    dist = np.where(valid[i_idx] & valid[j_idx], quad + log_term, np.inf)
    order = np.argsort(dist, kind="stable")
    used = np.zeros(M, dtype=bool)
    sel = []
    for k in order:
        if len(sel) >= max_pairs:
            break
        d = dist[k]
        if not (np.isfinite(d) and d < merge_threshold):
            break                # ascending order (inf / NaN last): nothing below the threshold is left
        i, j = int(i_idx[k]), int(j_idx[k])
        if used[i] or used[j]:
            continue
        used[i] = used[j] = True
        sel.append((i, j))
    for i, j in sel:
        w1, w2 = t["weights"][i], t["weights"][j]
        wsum = w1 + w2
        if not wsum > 0.0:
            continue
        mu_m = (w1 * mu[i] + w2 * mu[j]) / wsum
        d1, d2 = (mu[i] - mu_m).reshape(3, 1), (mu[j] - mu_m).reshape(3, 1)
        Sig_m = (w1 * (Sigma[i] + d1 @ d1.T) + w2 * (Sigma[j] + d2 @ d2.T)) / wsum + eps_psd * np.eye(3)
        Lam_m = np.linalg.inv(Sig_m)
        t["Lambdas"][i] = Lam_m
        t["thetas"][i] = Lam_m @ mu_m
        t["etas"][i] = (w1 * t["etas"][i] + w2 * t["etas"][j]) / wsum
        cam = t["cam_mass"][i] + t["cam_mass"][j]
        acc = t["rgb_cam_accum"][i] + t["rgb_cam_accum"][j]
        den = t["rgb_cam_denom"][i] + t["rgb_cam_denom"][j]
        rgb = np.where(cam > 0.0, np.clip(acc / np.maximum(den, eps_psd), 0.0, 1.0), np.array([0.5, 0.5, 0.5]))
        t["weights"][i] = wsum
        t["colors"][i] = rgb
        t["rgb"][i] = rgb
        t["cam_mass"][i] = cam
        t["lidar_mass"][i] = t["lidar_mass"][i] + t["lidar_mass"][j]
        t["rgb_cam_accum"][i] = acc
        t["rgb_cam_denom"][i] = den
        t["timestamps"][i] = max(t["timestamps"][i], t["timestamps"][j])
        t["created_timestamps"][i] = min(t["created_timestamps"][i], t["created_timestamps"][j])
        t["last_supported_scan_seq"][i] = max(t["last_supported_scan_seq"][i], t["last_supported_scan_seq"][j])
        t["last_update_scan_seq"][i] = max(t["last_update_scan_seq"][i], t["last_update_scan_seq"][j])
        t["weights"][j] = 0.0
        t["valid_mask"][j] = False
    n = len(sel)
    if n:
        t["count"] = int(np.sum(t["valid_mask"]))
    return t, n, ("merged" if n else "noop")
