"""
Seeded synthetic inputs for the LiDAR evidence path.  Host-side NumPy only;
used by tests/, bench.py and __graft_entry__.smoke().  No reference code is involved: the scene is a
20 x 8 x 3 m room with two pillars, sampled in VLP-16 firing order.

Sensor conventions mirrored from the reference:
  * range weights: fl/backend/backend_node.py:449-459 (sigmoid window 0.5 m .. 50 m, sigma 0.25, floor 1e-12)
  * extrinsic T_base_lidar: config/gc_unified.yaml:18-24 ([t, rotvec])
  * per-point stamps over a 0.1 s sweep: docs/KIMERA_DATASET_AND_PIPELINE.md:49,212
  * epoch time base 1.6657729e9: config/time_alignment/kimera_10_14_acl_jackal_005.yaml:4
"""

from __future__ import annotations

import numpy as np

from . import constants as C

T_BASE_LIDAR = np.array([-0.065447, -0.100474, 0.108987, -0.002723, -0.069383, 0.028979])
EPOCH_T0 = 1.6657729e9
SCAN_PERIOD = 0.1


def rotvec_to_matrix(r):
    r = np.asarray(r, dtype=np.float64)
    th = np.linalg.norm(r)
    K = np.array([[0, -r[2], r[1]], [r[2], 0, -r[0]], [-r[1], r[0], 0]])
    if th < 1e-12:
        return np.eye(3) + K
    return np.eye(3) + np.sin(th) / th * K + (1 - np.cos(th)) / th**2 * (K @ K)


def _ray_room(u, o):
    """Distance along unit rays u (N,3) from o (3,) to the room box + two pillars."""
    lo = np.array([-10.0, -4.0, -0.6])
    hi = np.array([10.0, 4.0, 2.4])
    with np.errstate(divide="ignore", invalid="ignore"):
        t_lo = (lo[None, :] - o[None, :]) / u
        t_hi = (hi[None, :] - o[None, :]) / u
    t_exit = np.where(u > 0, t_hi, t_lo)
    t_exit = np.where(np.abs(u) < 1e-12, np.inf, t_exit)
    r = np.min(t_exit, axis=1)
    for cx, cy, rad in ((3.0, 1.5, 0.3), (-4.0, -2.0, 0.4)):
        ox, oy = o[0] - cx, o[1] - cy
        a = u[:, 0] ** 2 + u[:, 1] ** 2
        b = 2 * (ox * u[:, 0] + oy * u[:, 1])
        c = ox * ox + oy * oy - rad * rad
        disc = b * b - 4 * a * c
        ok = (disc > 0) & (a > 1e-12)
        sq = np.sqrt(np.where(ok, disc, 0.0))
        t_c = np.where(ok, (-b - sq) / (2 * np.where(ok, a, 1.0)), np.inf)
        t_c = np.where(t_c > 0, t_c, np.inf)
        r = np.minimum(r, t_c)
    return r


def vlp16_scan(n_points: int, seed: int, t0: float = EPOCH_T0, sensor_xy=(0.0, 0.0)):
    """
    One synthetic VLP-16-shaped scan, already in the base frame.
    Returns (points (N,3) f64, stamps (N,) f64, weights (N,) f64, ring (N,) u8, tag (N,) u8).
    """
    rng = np.random.default_rng(seed)
    i = np.arange(n_points)
    ring = (i % 16).astype(np.uint8)
    n_cols = max(1, (n_points + 15) // 16)
    az = 2.0 * np.pi * (i // 16) / n_cols
    el = np.deg2rad(-15.0 + 2.0 * ring.astype(np.float64))
    u = np.stack([np.cos(el) * np.cos(az), np.cos(el) * np.sin(az), np.sin(el)], axis=1)
    o = np.array([sensor_xy[0], sensor_xy[1], 0.0])
    rng_m = _ray_room(u, o) + rng.normal(0.0, 0.02, size=n_points)
    rng_m = np.clip(rng_m, 0.5, 50.0)
    p_lidar = u * rng_m[:, None]
    dist = np.linalg.norm(p_lidar, axis=1)
    a = (dist - C.GC_RANGE_WEIGHT_MIN_R) / C.GC_RANGE_WEIGHT_SIGMA
    b = (C.GC_RANGE_WEIGHT_MAX_R - dist) / C.GC_RANGE_WEIGHT_SIGMA
    w = (1.0 / (1.0 + np.exp(-a))) * (1.0 / (1.0 + np.exp(-b)))
    w = w * (1.0 - C.GC_WEIGHT_FLOOR) + C.GC_WEIGHT_FLOOR
    R = rotvec_to_matrix(T_BASE_LIDAR[3:])
    p_base = p_lidar @ R.T + T_BASE_LIDAR[None, :3]
    t = t0 + SCAN_PERIOD * i / float(n_points)
    return (np.ascontiguousarray(p_base), t.astype(np.float64), w.astype(np.float64), ring,
            np.zeros(n_points, np.uint8))


# VLP-16 driver layout used on the wire (docs/KIMERA_DATASET_AND_PIPELINE.md section 6): x, y, z, intensity float32,
# ring uint16, time float32 -> 22 bytes per point.  (offset, sensor_msgs/PointField datatype)
PC2_VLP16_FIELDS = {"x": (0, 7), "y": (4, 7), "z": (8, 7), "intensity": (12, 7), "ring": (16, 4), "time": (18, 7)}
PC2_VLP16_POINT_STEP = 22


def vlp16_pointcloud2(n_points: int, seed: int, t0: float = 0.0, sensor_xy=(0.0, 0.0), time_unit: str = "s"):
    """
    The same synthetic sweep as vlp16_scan, as PointCloud2 wire bytes in the LIDAR frame (float32 coordinates, what the
    driver publishes).  Per-point time is the offset into the sweep in seconds ("s") or nanoseconds ("ns").
    Returns (data uint8 (N * 22,), fields, point_step).
    """
    rng = np.random.default_rng(seed)
    i = np.arange(n_points)
    ring = (i % 16).astype(np.uint16)
    n_cols = max(1, (n_points + 15) // 16)
    az = 2.0 * np.pi * (i // 16) / n_cols
    el = np.deg2rad(-15.0 + 2.0 * ring.astype(np.float64))
    u = np.stack([np.cos(el) * np.cos(az), np.cos(el) * np.sin(az), np.sin(el)], axis=1)
    o = np.array([sensor_xy[0], sensor_xy[1], 0.0])
    rng_m = np.clip(_ray_room(u, o) + rng.normal(0.0, 0.02, size=n_points), 0.5, 50.0)
    p = (u * rng_m[:, None]).astype(np.float32)
    toff = t0 + SCAN_PERIOD * i / float(n_points)
    if time_unit == "ns":
        toff = toff * 1e9
    rec = np.zeros(n_points, dtype=np.dtype({"names": ["x", "y", "z", "intensity", "ring", "time"],
                                             "formats": ["<f4", "<f4", "<f4", "<f4", "<u2", "<f4"],
                                             "offsets": [0, 4, 8, 12, 16, 18], "itemsize": PC2_VLP16_POINT_STEP}))
    rec["x"], rec["y"], rec["z"] = p[:, 0], p[:, 1], p[:, 2]
    rec["intensity"] = rng.uniform(0, 255, n_points).astype(np.float32)
    rec["ring"] = ring
    rec["time"] = toff.astype(np.float32)
    return np.frombuffer(rec.tobytes(), dtype=np.uint8).copy(), dict(PC2_VLP16_FIELDS), PC2_VLP16_POINT_STEP


def base_lidar_extrinsics():
    """(R_base_lidar (3,3), t_base_lidar (3,)) of config/gc_unified.yaml:18-24."""
    return rotvec_to_matrix(T_BASE_LIDAR[3:]), T_BASE_LIDAR[:3].copy()


def scan_twist(seed: int):
    """xi_body = [rho ~ U(-0.1,0.1)^3 m, phi ~ U(-0.05,0.05)^3 rad] (IMU-twist stand-in)."""
    rng = np.random.default_rng(seed)
    return np.concatenate([rng.uniform(-0.1, 0.1, 3), rng.uniform(-0.05, 0.05, 3)])


def hypothesis_poses(n_hyp: int, seed: int = 42):
    """H poses [t(3), rotvec(3)] = N(0, diag(0.05 m, 0.02 rad)) about the origin."""
    rng = np.random.default_rng(seed)
    return np.concatenate([rng.normal(0, 0.05, (n_hyp, 3)), rng.normal(0, 0.02, (n_hyp, 3))], axis=1)


def lidar_origin_base():
    return T_BASE_LIDAR[:3].copy()


def fibonacci_atlas(n_bins: int = C.GC_B_BINS):
    """48 quasi-uniform unit directions; same closed form as archive/bin_atlas.py:40-75."""
    i = np.arange(n_bins, dtype=np.float64) + 0.5
    phi = np.arccos(1 - 2 * i / n_bins)
    theta = np.pi * (1 + np.sqrt(5)) * i
    d = np.stack([np.sin(phi) * np.cos(theta), np.sin(phi) * np.sin(theta), np.cos(phi)], axis=1)
    return d / (np.linalg.norm(d, axis=1, keepdims=True) + C.GC_EPS_MASS)


def random_map_bin_stats(n_bins: int, seed: int, bin_dirs=None):
    """
    Plausible additive map-side bin statistics (archive/bin_atlas.py:83-105 layout) without running any
    operator: per bin a vMF-ish cloud of directions about the bin axis and a Gaussian cloud of positions.
    """
    rng = np.random.default_rng(seed)
    dirs = fibonacci_atlas(n_bins) if bin_dirs is None else np.asarray(bin_dirs)
    S_dir = np.zeros((n_bins, 3)); S_sc = np.zeros((n_bins, 3, 3)); N = np.zeros(n_bins)
    sum_p = np.zeros((n_bins, 3)); sum_pp = np.zeros((n_bins, 3, 3))
    for b in range(n_bins):
        m = 64
        w = rng.gamma(2.0, 1.0, m)
        u = dirs[b][None, :] + 0.25 * rng.normal(size=(m, 3))
        u /= np.linalg.norm(u, axis=1, keepdims=True)
        p = u * rng.uniform(2.0, 12.0, (m, 1))
        N[b] = w.sum()
        S_dir[b] = (w[:, None] * u).sum(0)
        S_sc[b] = np.einsum("n,ni,nj->ij", w, u, u)
        sum_p[b] = (w[:, None] * p).sum(0)
        sum_pp[b] = np.einsum("n,ni,nj->ij", w, p, p)
    return dict(S_dir=S_dir, S_dir_scatter=S_sc, N_dir=N, N_pos=N.copy(), sum_p=sum_p, sum_ppT=sum_pp)


# --------------------------------------------------------------------------------------------------
# primitive family: synthetic surfel map, camera splats
# --------------------------------------------------------------------------------------------------
def _tile_ids_from_xyz(XYZ, h):
    """Packed MA-hex tile ids (closed form of fl/common/tiling.py:126-145; 21 bits per axis, bias 2^20)."""
    XYZ = np.asarray(XYZ, np.float64).reshape(-1, 3)
    s2 = XYZ[:, 0] * 0.5 + XYZ[:, 1] * (np.sqrt(np.float64(3.0)) * 0.5)
    c1 = np.floor(XYZ[:, 0] / h).astype(np.int64)
    c2 = np.floor(s2 / h).astype(np.int64)
    cz = np.floor(XYZ[:, 2] / h).astype(np.int64)
    m = (1 << 21) - 1
    return (((c1 + (1 << 20)) & m) << 42) | (((c2 + (1 << 20)) & m) << 21) | ((cz + (1 << 20)) & m)


def room_surface_points(n, seed):
    """n points + outward-ish normals on the surfaces of the synthetic room (base frame)."""
    rng = np.random.default_rng(seed)
    i = np.arange(n)
    az = rng.uniform(0, 2 * np.pi, n)
    el = rng.uniform(-0.6, 0.6, n)
    u = np.stack([np.cos(el) * np.cos(az), np.cos(el) * np.sin(az), np.sin(el)], axis=1)
    r = np.clip(_ray_room(u, np.zeros(3)), 0.5, 50.0)
    p = u * r[:, None]
    # surface normal = axis of the wall that was hit (largest |coordinate| relative to the box), else radial
    lo = np.array([-10.0, -4.0, -0.6]); hi = np.array([10.0, 4.0, 2.4])
    d_lo = np.abs(p - lo[None]); d_hi = np.abs(p - hi[None])
    d = np.minimum(d_lo, d_hi)
    ax = np.argmin(d, axis=1)
    nrm = np.zeros((n, 3))
    nrm[i, ax] = np.where(d_lo[i, ax] < d_hi[i, ax], 1.0, -1.0)
    far = d[i, ax] > 0.05  # pillars
    nrm[far] = -u[far] * np.array([1.0, 1.0, 0.0])
    nrm[far] /= np.linalg.norm(nrm[far], axis=1, keepdims=True) + 1e-12
    R = rotvec_to_matrix(T_BASE_LIDAR[3:])
    return p @ R.T + T_BASE_LIDAR[None, :3], nrm @ R.T


def synthetic_atlas(n_surfels, m_tile, seed, scan_seq=10, h_tile=C.GC_H_TILE, n_lobes=C.GC_VMF_N_LOBES,
                    camera_fraction=0.2):
    """
    Synthetic surfel atlas as plain NumPy tile dicts (field names of fl/backend/structures/primitive_map.py:98-145).
    Surfels lie on the room surfaces, tiles are MA-hex cells (h = 2 m) keyed by position, at most m_tile per tile,
    slots are scattered (holes stay invalid), Lambda from plane-like covariances, eta = kappa * n, kappa ~ U(1,100),
    weights ~ Gamma(2,1), last_supported_scan_seq ~ U{0..scan_seq}, ids sequential.
    Returns dict(tiles={tile_id: tile}, next_global_id, total_count, m_tile).
    """
    rng = np.random.default_rng(seed)
    pos, nrm = room_surface_points(n_surfels, seed + 1)
    tids = _tile_ids_from_xyz(pos, h_tile)
    order = np.argsort(tids, kind="stable")
    pos, nrm, tids = pos[order], nrm[order], tids[order]
    uniq, start, counts = np.unique(tids, return_index=True, return_counts=True)
    tiles = {}
    next_id = 0
    for tid, s0, cnt in zip(uniq, start, counts):
        k = int(min(cnt, m_tile))
        slots = np.sort(rng.choice(m_tile, size=k, replace=False))
        mu, n = pos[s0:s0 + k], nrm[s0:s0 + k]
        # clutter: perturb the wall normals so that direction statistics are not rank deficient near the sensor
        n = n + 0.35 * rng.normal(size=n.shape)
        n = n / (np.linalg.norm(n, axis=1, keepdims=True) + 1e-12)
        # orthonormal basis with n as third axis
        a = np.where(np.abs(n[:, 2:3]) < 0.9, np.array([[0.0, 0.0, 1.0]]), np.array([[1.0, 0.0, 0.0]]))
        e1 = np.cross(n, a); e1 /= np.linalg.norm(e1, axis=1, keepdims=True) + 1e-12
        e2 = np.cross(n, e1)
        V = np.stack([e1, e2, n], axis=2)
        var = np.stack([rng.uniform(0.004, 0.02, k), rng.uniform(0.004, 0.02, k), rng.uniform(5e-5, 4e-4, k)], axis=1)
        Lam = np.einsum("kij,kj,klj->kil", V, 1.0 / var, V)
        t = dict(tile_id=int(tid), Lambdas=np.zeros((m_tile, 3, 3)), thetas=np.zeros((m_tile, 3)),
                 etas=np.zeros((m_tile, n_lobes, 3)), weights=np.zeros(m_tile), timestamps=np.zeros(m_tile),
                 created_timestamps=np.zeros(m_tile), last_supported_scan_seq=np.zeros(m_tile, np.int64),
                 last_update_scan_seq=np.zeros(m_tile, np.int64), primitive_ids=np.zeros(m_tile, np.int64),
                 valid_mask=np.zeros(m_tile, bool), colors=np.zeros((m_tile, 3)), cam_mass=np.zeros(m_tile),
                 lidar_mass=np.zeros(m_tile), rgb_cam_accum=np.zeros((m_tile, 3)), rgb_cam_denom=np.zeros(m_tile),
                 rgb=np.full((m_tile, 3), 0.5), next_local_id=int(slots.max()) + 1 if k else 0, count=k)
        t["Lambdas"][slots] = Lam
        t["thetas"][slots] = np.einsum("kij,kj->ki", Lam, mu)
        t["etas"][slots, 0, :] = rng.uniform(1.0, 100.0, (k, 1)) * n
        w = rng.gamma(2.0, 1.0, k)
        w[rng.uniform(size=k) < 0.01] = 5e-5          # a few below the cull threshold
        t["weights"][slots] = w
        t["timestamps"][slots] = EPOCH_T0 - rng.uniform(0, 5, k)
        t["created_timestamps"][slots] = EPOCH_T0 - 10.0
        ls = rng.integers(0, scan_seq + 1, k)
        t["last_supported_scan_seq"][slots] = ls
        t["last_update_scan_seq"][slots] = ls
        t["primitive_ids"][slots] = next_id + np.arange(k)
        t["valid_mask"][slots] = True
        is_cam = rng.uniform(size=k) < camera_fraction
        col = rng.uniform(0, 1, (k, 3))
        cm = np.where(is_cam, w, 0.0)
        t["cam_mass"][slots] = cm
        t["lidar_mass"][slots] = np.where(is_cam, 0.0, w)
        t["rgb_cam_accum"][slots] = col * cm[:, None]
        t["rgb_cam_denom"][slots] = cm
        t["rgb"][slots] = np.where(is_cam[:, None], col, 0.5)
        t["colors"][slots] = t["rgb"][slots]
        tiles[int(tid)] = t
        next_id += k
    return dict(tiles=tiles, next_global_id=next_id, total_count=next_id, m_tile=int(m_tile))


def camera_splats(n, seed):
    """Random camera splats in the body frame (positions 1-6 m ahead): positions, covariances, directions, kappas,
    weights, timestamps, colors -- the inputs of measurement_batch_from_camera_splats."""
    rng = np.random.default_rng(seed)
    pos = np.stack([rng.uniform(1.0, 6.0, n), rng.uniform(-2.5, 2.5, n), rng.uniform(-0.4, 1.5, n)], axis=1)
    A = rng.normal(size=(n, 3, 3)) * 0.03
    cov = A @ np.transpose(A, (0, 2, 1)) + 1e-4 * np.eye(3)[None]
    d = rng.normal(size=(n, 3)); d /= np.linalg.norm(d, axis=1, keepdims=True)
    return dict(positions=pos, covariances=cov, directions=d, kappas=rng.uniform(1, 50, n), weights=rng.uniform(0.2, 1.0, n),
                timestamps=np.full(n, EPOCH_T0 + 0.05), colors=rng.uniform(0, 1, (n, 3)))


def imu_window(n_total: int, n_valid: int, seed: int, t_start: float = EPOCH_T0, rate_hz: float = 200.0,
               lead_s: float = 0.02, gyro_scale: float = 0.4, accel_noise: float = 0.3):
    """
    IMU buffer as the pipeline hands it to the preintegration (fixed length GC_MAX_IMU_PREINT_LEN = 512,
    fl/common/constants.py:65-67): `n_valid` samples at `rate_hz` starting `lead_s` before the scan, zero-padded
    stamps / rates after them.  Gyro: slowly varying body rates; accel: specific force of a body near rest
    (+9.81 on z) plus a smooth manoeuvre and white noise.  Returns (stamps (M,), gyro (M,3), accel (M,3)).
    """
    rng = np.random.default_rng(seed)
    stamps = np.zeros(n_total)
    gyro = np.zeros((n_total, 3))
    accel = np.zeros((n_total, 3))
    k = np.arange(n_valid)
    tt = k / rate_hz
    stamps[:n_valid] = t_start - lead_s + tt
    w0 = rng.uniform(-gyro_scale, gyro_scale, 3)
    w1 = rng.uniform(-gyro_scale, gyro_scale, 3)
    gyro[:n_valid] = w0[None, :] + np.sin(2.0 * np.pi * 1.3 * tt)[:, None] * w1[None, :] + rng.normal(0.0, 0.01, (n_valid, 3))
    a0 = rng.uniform(-1.0, 1.0, 3)
    accel[:n_valid] = np.array([0.0, 0.0, 9.81])[None, :] + np.cos(2.0 * np.pi * 0.7 * tt)[:, None] * a0[None, :] \
        + rng.normal(0.0, accel_noise, (n_valid, 3))
    return stamps, gyro, accel


def imu_hypothesis_params(n_hyp: int, seed: int):
    """Per-hypothesis inputs of the twist prologue: start orientation (rotvec), gyro / accel bias, time-warp sigma."""
    rng = np.random.default_rng(seed)
    return dict(rotvec0=rng.uniform(-0.6, 0.6, (n_hyp, 3)), gyro_bias=rng.normal(0.0, 0.01, (n_hyp, 3)),
                accel_bias=rng.normal(0.0, 0.05, (n_hyp, 3)), sigma=rng.uniform(0.01, 0.03, n_hyp))


def hypothesis_evidence_stack(n_hyp: int, dim: int, seed: int, indefinite: bool = False):
    """
    K posterior information pairs (L_k, h_k), linearisation points and hypothesis weights as the combine receives them
    (backend_node.py:2036-2097): SPD information matrices with eigenvalues over nine decades (pose blocks strong, bias /
    extrinsic blocks weak), means scattered around a common state.  `indefinite`: nearly singular matrices (eigenvalues
    down to 1e-15, below eps_psd, so that the projection clamps), one asymmetric input and weights under the floor.
    """
    rng = np.random.default_rng(seed)
    Ls, hs, zs = [], [], []
    mu0 = rng.normal(0.0, 1.0, dim)
    for k in range(n_hyp):
        Q, _ = np.linalg.qr(rng.normal(size=(dim, dim)))
        ev = 10.0 ** rng.uniform(-3.0, 6.0, dim)
        if indefinite:
            ev[:3] = 10.0 ** rng.uniform(-15.0, -13.0, 3)
            Q = np.linalg.qr(np.random.default_rng(seed + 1000).normal(size=(dim, dim)))[0]   # shared weak directions
        L = (Q * ev) @ Q.T
        L = 0.5 * (L + L.T)
        if indefinite and k == 2:
            L[0, 1] += 1e-11         # asymmetric input: sym_delta > 0, still positive definite once lifted
        mu = mu0 + rng.normal(0.0, 0.05, dim)
        Ls.append(L); hs.append(L @ mu); zs.append(mu0 + rng.normal(0.0, 0.01, dim))
    w = rng.dirichlet(np.ones(n_hyp))
    if indefinite:
        w[0] = 1e-4
        w[-1] = 0.0
    return np.stack(Ls), np.stack(hs), np.stack(zs), w


def fusion_inputs(n_hyp: int, seed: int, indefinite: bool = False):
    """
    Inputs of the evidence-fusion step (pipeline.py:1038-1207) for K hypotheses: LiDAR evidence living on the pose block
    (as build_combined_lidar_evidence_22d embeds it), IMU + odometry evidence coupling pose, velocity, biases and the time
    offset, a predicted belief (L, h) with eigenvalues over eight decades, and the certificate-level scalars of the two
    control laws.  `indefinite`: priors with a few negative eigenvalues and an asymmetric entry, so that the PSD
    projection of the posterior clamps and reports a projection delta.
    """
    rng = np.random.default_rng(seed)
    D = 22
    out = {k: [] for k in ("L_lidar", "h_lidar", "L_other", "h_other", "L_prior", "h_prior")}
    for k in range(n_hyp):
        a = rng.normal(size=(6, 6))
        Ll = np.zeros((D, D))
        Ll[:6, :6] = a @ a.T * 10.0 ** rng.uniform(1.0, 4.0) + np.diag(10.0 ** rng.uniform(0.0, 3.0, 6))
        Ll[2, 2] *= 10.0 ** rng.uniform(-2.0, 1.0)          # weak / strong z: drives z_to_xy
        hl = Ll @ rng.normal(0.0, 0.05, D)
        b = rng.normal(size=(16, 16)) * 10.0 ** rng.uniform(-1.0, 1.5, 16)[None, :]
        Lo = np.zeros((D, D))
        Lo[:16, :16] = b @ b.T
        Lo[15, 6:9] *= 10.0 ** rng.uniform(-1.0, 1.0); Lo[6:9, 15] = Lo[15, 6:9]      # dt-velocity coupling vs dt-pose: dt_asymmetry
        Lo[16:, 16:] = np.diag(10.0 ** rng.uniform(-4.0, 0.0, 6))
        ho = Lo @ rng.normal(0.0, 0.05, D)
        Q, _ = np.linalg.qr(rng.normal(size=(D, D)))
        ev = 10.0 ** rng.uniform(-2.0, 6.0, D)
        if indefinite:
            ev[:2] = -10.0 ** rng.uniform(3.0, 5.0, 2)
        Lp = (Q * ev) @ Q.T
        Lp = 0.5 * (Lp + Lp.T)
        if indefinite:
            Lp[3, 7] += 1e-3
        hp = Lp @ rng.normal(0.0, 0.1, D)
        for key, v in zip(out, (Ll, hl, Lo, ho, Lp, hp)):
            out[key].append(v)
    res = {k: np.stack(v) for k, v in out.items()}
    res.update(ess_total=10.0 ** rng.uniform(0.0, 4.0, n_hyp), dt_effect=10.0 ** rng.uniform(-2.0, 1.0, n_hyp),
               extrinsic_effect=10.0 ** rng.uniform(-2.0, 1.0, n_hyp), nll_per_ess=rng.uniform(0.0, 2.0, n_hyp))
    return res
