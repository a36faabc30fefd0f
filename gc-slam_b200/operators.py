"""
Bin-family LiDAR evidence operators -- host-side mirror of the reference's Python operator interface.

Same names, argument meaning, result field names, certificates and error behaviour as the reference
(``fl/`` = fl_ws/src/fl_slam_poc/fl_slam_poc/ in whabacivch/GC-SLAM):

  point_budget_resample            fl/backend/operators/point_budget.py:117-221
  deskew_constant_twist            fl/backend/operators/deskew_constant_twist.py:72-117
  bin_soft_assign                  archive/legacy_operators/binning.py:79-131
  scan_bin_moment_match            archive/legacy_operators/binning.py:212-324
  kappa_from_resultant_batch / _v2 fl/backend/operators/kappa.py:130-234
  create_fibonacci_atlas, MapBinStats, update_map_stats, apply_forgetting, compute_map_derived_stats
                                   archive/bin_atlas.py:40-257
  matrix_fisher_rotation_evidence  archive/legacy_operators/matrix_fisher_evidence.py:264-394
  planar_translation_evidence      archive/legacy_operators/matrix_fisher_evidence.py:502-671
  build_combined_lidar_evidence_22d  :729-756
plus the fused entry ``lidar_evidence_bins`` (README.md:105-120 steps 1,3-8 and the LiDAR term of 9).

Every function returns ``(result, CertBundle, ExpectedEffect)``.  Arrays in results are torch CUDA float64
tensors; inputs may be NumPy arrays or torch tensors on any device.  All arithmetic happens in
libgcs_b200.so (CUDA, sm_100a) -- there is no CPU path in this module.
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
from .certs import (CertBundle, ComputeCert, ConditioningCert, DeviceRuntimeCert, ExpectedEffect, InfluenceCert,
                    MismatchCert, SupportCert)

F64 = torch.float64


# --------------------------------------------------------------------------------------------------
# marshalling helpers
# --------------------------------------------------------------------------------------------------
class _IO:
    """Counts the bytes this call really moved between host and device (DeviceRuntimeCert)."""

    def __init__(self, device=None, ctx=None):
        self.ctx = ctx if ctx is not None else L.context(device)
        self.dev = torch.device("cuda", self.ctx.device)
        self.h2d = 0
        self.d2h = 0
        self.syncs = 0

    def dev_in(self, x, dtype=F64, shape=None):
        if x is None:
            return None
        if isinstance(x, torch.Tensor):
            t = x
            if not t.is_cuda:
                self.h2d += t.numel() * t.element_size()
            t = t.to(device=self.dev, dtype=dtype)
        else:
            a = np.ascontiguousarray(np.asarray(x), dtype={F64: np.float64, torch.uint8: np.uint8,
                                                           torch.int32: np.int32, torch.int64: np.int64}[dtype])
            self.h2d += a.nbytes
            t = torch.from_numpy(a).to(self.dev)
        t = t.contiguous()
        if shape is not None:
            t = t.reshape(shape)
        return t

    def empty(self, *shape, dtype=F64):
        return torch.empty(shape, dtype=dtype, device=self.dev)

    def zeros(self, *shape, dtype=F64):
        return torch.zeros(shape, dtype=dtype, device=self.dev)

    def host(self, t):
        """One batched device->host read of certificate scalars (a host sync)."""
        self.d2h += t.numel() * t.element_size()
        self.syncs += 1
        return t.detach().cpu().numpy()

    def stream(self):
        return L.stream_ptr(self.dev)

    def runtime_cert(self):
        return DeviceRuntimeCert(host_sync_count_est=self.syncs, device_to_host_bytes_est=self.d2h,
                                 host_to_device_bytes_est=self.h2d, jit_recompile_count=0)

    def compute(self, **kw):
        return ComputeCert(device_runtime=self.runtime_cert(), **kw)


# ---- deferred certificate reads ------------------------------------------------------------------------------------
# An operator body is a generator: it enqueues its kernels, then yields (io, device tensor of certificate scalars,
# provisional result) and is resumed with the host copy of that tensor to assemble (result, CertBundle, ExpectedEffect).
# `_drive` runs one operator the way the reference's call sites expect (one batched read-back, i.e. one host sync per
# operator).  `drive_group` advances several operators to their yield first -- later ones may consume the provisional
# (device-side) results of earlier ones -- and serves all their read-backs with ONE stream synchronisation.
import threading

_PINNED = threading.local()      # per calling thread (contexts are per thread: two threads never share a read-back buffer)


def _pinned_like(t, slot):
    """Pinned read-back buffer for (device, slot, dtype, size), private to the calling thread."""
    cache = getattr(_PINNED, "bufs", None)
    if cache is None:
        cache = _PINNED.bufs = {}
    key = (t.device.index, slot, t.dtype, t.numel())
    buf = cache.get(key)
    if buf is None:
        buf = cache[key] = torch.empty(t.numel(), dtype=t.dtype).pin_memory()
    return buf


def _drive(gen):
    try:
        io, t, _prov = next(gen)
        while True:
            io, t, _prov = gen.send(io.host(t))
    except StopIteration as e:
        return e.value


class _Pending:
    """An operator advanced to its certificate read-back; `provisional` holds its device-side outputs."""

    def __init__(self, gen):
        self.gen, self.done, self.value = gen, False, None
        try:
            self.io, self.tensor, self.provisional = next(gen)
        except StopIteration as e:           # early exit without any device work
            self.done, self.value, self.provisional = True, e.value, (e.value[0] if isinstance(e.value, tuple) else e.value)


def drive_group(pending):
    """Finish a list of _Pending operators with one stream synchronisation; returns their final values in order."""
    live = [p for p in pending if not p.done]
    bufs, events = [], {}
    for k, p in enumerate(live):
        b = _pinned_like(p.tensor, k)
        with torch.cuda.device(p.io.dev):
            b.copy_(p.tensor.detach().reshape(-1), non_blocking=True)
            # an event on the stream that performed this copy (operators of one group may live on different devices)
            ev = torch.cuda.Event()
            ev.record(torch.cuda.current_stream(p.io.dev))
        events[p.io.dev] = ev
        bufs.append(b)
    for ev in events.values():
        ev.synchronize()
    for k, (p, b) in enumerate(zip(live, bufs)):
        p.io.d2h += b.numel() * b.element_size()
        p.io.syncs += 1 if k == 0 else 0
        try:
            p.gen.send(b.numpy().reshape(tuple(p.tensor.shape)).copy())
            raise RuntimeError("operator body yielded twice: not usable in a fused group")
        except StopIteration as e:
            p.done, p.value = True, e.value
    return [p.value for p in pending]


def _dptr(a):
    return (C.c_double * len(a))(*[float(v) for v in a])


def _host_vec(x, n):
    if isinstance(x, torch.Tensor):
        x = x.detach().cpu().numpy()
    v = np.asarray(x, dtype=np.float64).reshape(-1)
    if v.shape[0] != n:
        raise ValueError(f"expected a vector of {n} values, got shape {v.shape}")
    return v


# --------------------------------------------------------------------------------------------------
# result types (field names = the reference's)
# --------------------------------------------------------------------------------------------------
@dataclass
class PointBudgetResult:
    points: torch.Tensor
    timestamps: torch.Tensor
    weights: torch.Tensor
    ring: torch.Tensor
    tag: torch.Tensor
    n_input: int
    n_output: int
    total_mass_in: float
    total_mass_out: float


@dataclass
class DeskewConstantTwistResult:
    points: torch.Tensor
    timestamps: torch.Tensor
    weights: torch.Tensor
    ess_imu: float


@dataclass
class BinSoftAssignResult:
    responsibilities: torch.Tensor


@dataclass
class ScanBinStats:
    N: torch.Tensor
    s_dir: torch.Tensor
    S_dir_scatter: torch.Tensor
    p_bar: torch.Tensor
    Sigma_p: torch.Tensor
    kappa_scan: torch.Tensor
    # additive raw sums (not in the reference dataclass; needed by the bin-map update)
    sum_p: Optional[torch.Tensor] = None
    sum_ppT: Optional[torch.Tensor] = None


@dataclass
class KappaResult:
    kappa: float
    R_clamped: float
    clamp_delta: float


@dataclass
class BinAtlas:
    dirs: torch.Tensor


@dataclass
class MapBinStats:
    S_dir: torch.Tensor
    S_dir_scatter: torch.Tensor
    N_dir: torch.Tensor
    N_pos: torch.Tensor
    sum_p: torch.Tensor
    sum_ppT: torch.Tensor

    def _c(self):
        m = L.MapBinStats()
        m.S_dir, m.S_scatter, m.N_dir = L.ptr(self.S_dir), L.ptr(self.S_dir_scatter), L.ptr(self.N_dir)
        m.N_pos, m.sum_p, m.sum_ppT = L.ptr(self.N_pos), L.ptr(self.sum_p), L.ptr(self.sum_ppT)
        return m


@dataclass
class ScatterMetrics:
    eigenvalues: torch.Tensor
    eigenvectors: torch.Tensor
    linearity: float
    planarity: float
    sphericity: float
    anisotropy: float
    effective_rank: float


@dataclass
class MatrixFisherResult:
    R_mf: torch.Tensor
    L_rot: torch.Tensor
    h_rot: torch.Tensor
    delta_rot: torch.Tensor
    svd_singular_values: torch.Tensor
    map_scatter_metrics: ScatterMetrics
    scan_scatter_metrics: ScatterMetrics


@dataclass
class PlanarTranslationResult:
    t_wls: torch.Tensor
    L_trans: torch.Tensor
    h_trans: torch.Tensor
    delta_trans: torch.Tensor
    xy_info_scale: float
    z_info_scale: float


# --------------------------------------------------------------------------------------------------
# a1 PointBudgetResample
# --------------------------------------------------------------------------------------------------
def pc2_layout(fields, point_step: int) -> "L.Pc2Layout":
    """
    {name: (offset, PointField datatype)} -> gcs_pc2_layout.  Field selection as parse_pointcloud2_vlp16
    (backend_node.py:396-417): x, y, z, ring required (RuntimeError otherwise), per-point time from "t" else "time".
    """
    missing = [k for k in ("x", "y", "z", "ring") if k not in fields]
    if missing:
        raise RuntimeError(f"PointCloud2 (VLP-16 layout) missing required fields: {missing}. "
                           f"Present fields: {sorted(list(fields.keys()))}")
    lay = L.Pc2Layout()
    lay.point_step = int(point_step)
    for k in ("x", "y", "z", "ring"):
        setattr(lay, "off_" + k, int(fields[k][0]))
        setattr(lay, "type_" + k, int(fields[k][1]))
    tf = "t" if "t" in fields else ("time" if "time" in fields else None)
    lay.off_time, lay.type_time = (int(fields[tf][0]), int(fields[tf][1])) if tf else (-1, 0)
    return lay


def _pad16(data) -> torch.Tensor:
    """uint8 host payload -> uint8 host tensor whose length is a multiple of 16 (the device decoder reads 16-byte words)."""
    t = data if isinstance(data, torch.Tensor) else torch.from_numpy(np.frombuffer(data, dtype=np.uint8).copy() if isinstance(data, (bytes, bytearray, memoryview)) else np.ascontiguousarray(data, dtype=np.uint8))
    t = t.reshape(-1)
    pad = (-t.numel()) % 16
    return torch.cat([t, torch.zeros(pad, dtype=torch.uint8)]) if pad else t


def parse_pointcloud2_vlp16(data, n_points: int, point_step: int, fields, header_stamp_sec: float = 0.0,
                            R_base_lidar=None, t_base_lidar=None, device=None):
    """
    Device-side parse_pointcloud2_vlp16 (backend_node.py:377-468) fused with the LiDAR -> base transform of the scan
    callback (backend_node.py:1682-1684; skipped when R_base_lidar is None).  `data` is the message payload (bytes /
    uint8 array / uint8 tensor, host or device).  Returns (points (N,3), timestamps (N,), weights (N,), ring (N,) u8,
    tag (N,) u8) as CUDA tensors plus dict(n_nonfinite, time_rescaled).
    """
    io = _IO(device)
    lay = pc2_layout(fields, point_step)
    n = int(n_points)
    pts, t, w = io.empty(max(n, 0), 3), io.empty(max(n, 0)), io.empty(max(n, 0))
    ring, tag = io.empty(max(n, 0), dtype=torch.uint8), io.empty(max(n, 0), dtype=torch.uint8)
    if n <= 0:                                                                 # backend_node.py:386-394
        return pts, t, w, ring, tag, dict(n_nonfinite=0, time_rescaled=False)
    if isinstance(data, torch.Tensor) and data.is_cuda:
        d_dev = data.reshape(-1)
        if d_dev.data_ptr() % 16:          # a view with an odd storage offset: the decoder needs 16-byte alignment
            d_dev = d_dev.clone()
        if d_dev.numel() % 16:
            d_dev = torch.cat([d_dev, torch.zeros((-d_dev.numel()) % 16, dtype=torch.uint8, device=d_dev.device)])
    else:
        host = _pad16(data)
        d_dev = host.to(io.dev)
        io.h2d += host.numel()
    if d_dev.numel() < n * int(point_step):
        raise ValueError(f"payload has {d_dev.numel()} bytes, {n} points of {point_step} bytes need {n * int(point_step)}")
    stamp = io.dev_in(np.array([header_stamp_sec], np.float64))
    cert = io.zeros(1, L.PC_NCERT)
    R = None if R_base_lidar is None else (C.c_double * 9)(*np.asarray(R_base_lidar, np.float64).reshape(9))
    tb = None if t_base_lidar is None else (C.c_double * 3)(*np.asarray(t_base_lidar, np.float64).reshape(3))
    io.ctx.check(io.ctx.lib.gcs_parse_pointcloud2_vlp16(io.ctx.handle, io.stream(), L.ptr(d_dev), 1, n, C.byref(lay), L.ptr(stamp),
                                                        R, tb, L.ptr(pts), L.ptr(t), L.ptr(w), L.ptr(ring), L.ptr(tag), L.ptr(cert)))
    c = io.host(cert).reshape(-1)
    return pts, t, w, ring, tag, dict(n_nonfinite=int(c[L.PC_N_NONFINITE]), time_rescaled=bool(c[L.PC_TIME_RESCALED]))


def point_budget_resample(points, timestamps, weights, ring=None, tag=None,
                          n_points_cap: int = constants.GC_N_POINTS_CAP, chart_id: str = constants.GC_CHART_ID,
                          anchor_id: str = "initial") -> Tuple[PointBudgetResult, CertBundle, ExpectedEffect]:
    io = _IO()
    pts = io.dev_in(points, shape=(-1, 3))
    n_input = pts.shape[0]
    t = io.dev_in(timestamps, shape=(-1,))
    w = io.dev_in(weights, shape=(-1,))
    if t.shape[0] != n_input or w.shape[0] != n_input:
        raise ValueError(f"point_budget_resample: points {tuple(pts.shape)}, timestamps {tuple(t.shape)}, "
                         f"weights {tuple(w.shape)} disagree")
    rg = io.dev_in(ring, dtype=torch.uint8, shape=(-1,))
    tg = io.dev_in(tag, dtype=torch.uint8, shape=(-1,))
    cap = int(n_points_cap)
    o_pts, o_t, o_w = io.empty(cap, 3), io.empty(cap), io.empty(cap)
    o_rg, o_tg = io.empty(cap, dtype=torch.uint8), io.empty(cap, dtype=torch.uint8)
    cert_d = io.zeros(L.RS_NCERT)
    io.ctx.check(io.ctx.lib.gcs_point_budget_resample(
        io.ctx.handle, io.stream(), L.ptr(pts), L.ptr(t), L.ptr(w), L.ptr(rg), L.ptr(tg), n_input, cap,
        constants.GC_EPS_MASS, L.ptr(o_pts), L.ptr(o_t), L.ptr(o_w), L.ptr(o_rg), L.ptr(o_tg), L.ptr(cert_d)))
    c = io.host(cert_d)
    stride = max(1, int(math.ceil(n_input / cap)))
    n_sel = (n_input + stride - 1) // stride
    mass_in = float(c[L.RS_MASS_IN])
    ess = float(c[L.RS_ESS])
    result = PointBudgetResult(points=o_pts, timestamps=o_t, weights=o_w, ring=o_rg, tag=o_tg, n_input=int(n_input),
                               n_output=int(n_sel), total_mass_in=mass_in, total_mass_out=mass_in)
    cert = CertBundle.create_approx(
        chart_id=chart_id, anchor_id=anchor_id, triggers=["PointBudgetResample"],
        support=SupportCert(ess_total=ess, support_frac=float(min(1.0, cap / (n_input + constants.GC_EPS_MASS)))),
        influence=InfluenceCert.identity().with_overrides(
            mass_epsilon_ratio=constants.GC_EPS_MASS / (mass_in + constants.GC_EPS_MASS)),
        compute=io.compute())
    return result, cert, ExpectedEffect("predicted_ess", ess, None)


# --------------------------------------------------------------------------------------------------
# a2 DeskewConstantTwist
# --------------------------------------------------------------------------------------------------
def deskew_constant_twist(points, timestamps, weights, scan_start_time: float, scan_end_time: float, xi_body,
                          ess_imu: float, chart_id: str, anchor_id: str
                          ) -> Tuple[DeskewConstantTwistResult, CertBundle, ExpectedEffect]:
    return _drive(_deskew_constant_twist_gen(points, timestamps, weights, scan_start_time, scan_end_time, xi_body, ess_imu,
                                             chart_id, anchor_id))


def _deskew_constant_twist_gen(points, timestamps, weights, scan_start_time, scan_end_time, xi_body, ess_imu, chart_id,
                               anchor_id):
    io = _IO()
    pts = io.dev_in(points, shape=(-1, 3))
    t = io.dev_in(timestamps, shape=(-1,))
    w = io.dev_in(weights, shape=(-1,))
    n = pts.shape[0]
    if t.shape[0] != n or w.shape[0] != n:
        raise ValueError("deskew_constant_twist: points/timestamps/weights length mismatch")
    xi = _host_vec(xi_body, 6)
    o_pts, o_w = io.empty(n, 3), io.empty(n)
    cert_d = io.zeros(L.DK_NCERT)
    io.ctx.check(io.ctx.lib.gcs_deskew_constant_twist(
        io.ctx.handle, io.stream(), L.ptr(pts), L.ptr(t), L.ptr(w), n, _dptr(xi), float(scan_start_time),
        float(scan_end_time), L.ptr(o_pts), L.ptr(o_w), L.ptr(cert_d)))
    result = DeskewConstantTwistResult(points=o_pts, timestamps=t, weights=o_w, ess_imu=float(ess_imu))
    c = yield io, cert_d, result
    return (result,) + _deskew_cert(c, ess_imu, chart_id, anchor_id, io.compute())


def _deskew_cert(c, ess_imu, chart_id, anchor_id, compute):
    """Certificate + effect of deskew_constant_twist from its two mass sums (deskew_constant_twist.py:98-117)."""
    retained = float(c[L.DK_SUM_W_OUT] / (c[L.DK_SUM_W_IN] + constants.GC_EPS_MASS))
    cert = CertBundle.create_exact(chart_id=chart_id, anchor_id=anchor_id,
                                   support=SupportCert(ess_total=float(ess_imu), support_frac=retained),
                                   influence=InfluenceCert.identity(), compute=compute)
    return cert, ExpectedEffect("deskew_variance_reduction_proxy", 0.0, None)


def ray_directions(points, origin, eps: float = constants.GC_EPS_MASS) -> torch.Tensor:
    """Inline step of the pipeline (fl/backend/pipeline.py:589-593): d = (p - o) / (|p - o| + eps)."""
    io = _IO()
    pts = io.dev_in(points, shape=(-1, 3))
    out = io.empty(pts.shape[0], 3)
    io.ctx.check(io.ctx.lib.gcs_ray_directions(io.ctx.handle, io.stream(), L.ptr(pts), pts.shape[0],
                                               _dptr(_host_vec(origin, 3)), float(eps), L.ptr(out)))
    return out


# --------------------------------------------------------------------------------------------------
# a4 BinSoftAssign
# --------------------------------------------------------------------------------------------------
def create_fibonacci_atlas(n_bins: int = constants.GC_B_BINS) -> BinAtlas:
    """Fixed bin atlas (archive/bin_atlas.py:40-75): closed form, evaluated once on the host at init."""
    from .synth import fibonacci_atlas

    io = _IO()
    return BinAtlas(dirs=io.dev_in(fibonacci_atlas(n_bins)))


def bin_soft_assign(point_directions, bin_directions, tau: float = constants.GC_TAU_SOFT_ASSIGN,
                    chart_id: str = constants.GC_CHART_ID, anchor_id: str = "initial",
                    precision: int = L.PREC_F64) -> Tuple[BinSoftAssignResult, CertBundle, ExpectedEffect]:
    io = _IO()
    d = io.dev_in(point_directions, shape=(-1, 3))
    b = io.dev_in(bin_directions, shape=(-1, 3))
    n, nb = d.shape[0], b.shape[0]
    resp = io.empty(n, nb)
    cert_d = io.zeros(L.SA_NCERT)
    io.ctx.check(io.ctx.lib.gcs_bin_soft_assign(io.ctx.handle, io.stream(), L.ptr(d), n, L.ptr(b), nb, float(tau),
                                                constants.GC_EPS_MASS, int(precision), L.ptr(resp), L.ptr(cert_d)))
    c = io.host(cert_d)
    avg_entropy = float(c[L.SA_ENTROPY_SUM] / (float(n) + constants.GC_EPS_MASS))
    cert = CertBundle.create_exact(
        chart_id=chart_id, anchor_id=anchor_id,
        support=SupportCert(ess_total=float(math.exp(avg_entropy)), support_frac=float(c[L.SA_MAX_RESP])),
        compute=io.compute(largest_tensor_shape=(int(n), int(nb))))
    return BinSoftAssignResult(responsibilities=resp), cert, ExpectedEffect("predicted_assignment_entropy", avg_entropy, None)


# --------------------------------------------------------------------------------------------------
# a5 ScanBinMomentMatch
# --------------------------------------------------------------------------------------------------
def _alloc_stats(io, U, B):
    st = dict(N=io.empty(U, B), s_dir=io.empty(U, B, 3), S_scatter=io.empty(U, B, 3, 3), p_bar=io.empty(U, B, 3),
              Sigma_p=io.empty(U, B, 3, 3), kappa=io.empty(U, B), sum_p=io.empty(U, B, 3), sum_ppT=io.empty(U, B, 3, 3))
    cs = L.BinStats()
    for k, v in st.items():
        setattr(cs, k, L.ptr(v))
    return st, cs


def scan_bin_moment_match(points, point_covariances, weights, responsibilities, point_lambda=None,
                          direction_origin=None, eps_psd: float = constants.GC_EPS_PSD,
                          eps_mass: float = constants.GC_EPS_MASS, chart_id: str = constants.GC_CHART_ID,
                          anchor_id: str = "initial") -> Tuple[ScanBinStats, CertBundle, ExpectedEffect]:
    io = _IO()
    p = io.dev_in(points, shape=(-1, 3))
    n = p.shape[0]
    w = io.dev_in(weights, shape=(-1,))
    r = io.dev_in(responsibilities)
    if r.dim() != 2 or r.shape[0] != n:
        raise ValueError(f"responsibilities must be (N,B), got {tuple(r.shape)} for N={n}")
    nb = r.shape[1]
    cov = None
    if point_covariances is not None:
        cov = io.dev_in(point_covariances, shape=(-1, 3, 3))
        if cov.shape[0] != n:
            raise ValueError(f"point_covariances must be (N,3,3), got {tuple(cov.shape)}")
    if direction_origin is None:
        origin = np.zeros(3)
    else:
        origin = np.asarray(direction_origin.detach().cpu().numpy() if isinstance(direction_origin, torch.Tensor)
                            else direction_origin, dtype=np.float64).reshape(-1)
        if origin.shape[0] != 3:
            raise ValueError(f"direction_origin must be (3,), got {origin.shape}")
    lam = None
    if point_lambda is not None:
        lam = io.dev_in(point_lambda, shape=(-1,))
        if lam.shape[0] != n:
            raise ValueError(f"point_lambda must be (N,), got {tuple(lam.shape)} for N={n}")
    st, cs = _alloc_stats(io, 1, nb)
    cert_d = io.zeros(L.ST_NCERT)
    io.ctx.check(io.ctx.lib.gcs_scan_bin_moment_match(
        io.ctx.handle, io.stream(), L.ptr(p), L.ptr(cov), L.ptr(w), L.ptr(r), L.ptr(lam), _dptr(origin), n, nb,
        float(eps_psd), float(eps_mass), C.byref(cs), L.ptr(cert_d)))
    c = io.host(cert_d)
    result = ScanBinStats(N=st["N"][0], s_dir=st["s_dir"][0], S_dir_scatter=st["S_scatter"][0], p_bar=st["p_bar"][0],
                          Sigma_p=st["Sigma_p"][0], kappa_scan=st["kappa"][0], sum_p=st["sum_p"][0],
                          sum_ppT=st["sum_ppT"][0])
    ess = float(c[L.ST_ESS])
    cert = CertBundle.create_approx(
        chart_id=chart_id, anchor_id=anchor_id, triggers=["ScanBinMomentMatch"],
        support=SupportCert(ess_total=ess, support_frac=float(c[L.ST_SUPPORT_FRAC])),
        influence=InfluenceCert(lift_strength=0.0, psd_projection_delta=float(c[L.ST_PSD_DELTA]),
                                mass_epsilon_ratio=float(c[L.ST_MASS_EPS_RATIO]), anchor_drift_rho=0.0, dt_scale=1.0,
                                extrinsic_scale=1.0, trust_alpha=1.0),
        compute=io.compute(psd_projection_count=int(nb)))
    return result, cert, ExpectedEffect("predicted_ess", ess, None)


# --------------------------------------------------------------------------------------------------
# a6 KappaFromResultant
# --------------------------------------------------------------------------------------------------
def kappa_from_resultant_batch(R_bar, eps_r: float = constants.GC_EPS_R, d: int = 3,
                               r0: float = constants.GC_KAPPA_BLEND_R0, tau: float = constants.GC_KAPPA_BLEND_TAU):
    io = _IO()
    r = io.dev_in(R_bar)
    shape = r.shape
    r = r.reshape(-1).contiguous()
    out = io.empty(r.shape[0])
    io.ctx.check(io.ctx.lib.gcs_kappa_from_resultant_batch(io.ctx.handle, io.stream(), L.ptr(r), r.shape[0],
                                                           float(eps_r), float(d), float(r0), float(tau), L.ptr(out)))
    return out.reshape(shape)


def kappa_from_resultant_v2(R_bar: float, eps_r: float = constants.GC_EPS_R, eps_den: float = None,
                            chart_id: str = constants.GC_CHART_ID, anchor_id: str = "initial"
                            ) -> Tuple[KappaResult, CertBundle, ExpectedEffect]:
    """Scalar operator: same kernel as the batch variant (the reference pins batch == scalar at rtol 1e-10,
    test/test_audit_invariants.py:412-426)."""
    x = float(R_bar)
    Rc = min(max(x, 0.0), 1.0 - eps_r)
    k = float(kappa_from_resultant_batch(np.array([x]), eps_r=eps_r)[0].item())
    cert = CertBundle.create_approx(chart_id=chart_id, anchor_id=anchor_id, triggers=["KappaLowRApproximation"])
    return KappaResult(kappa=k, R_clamped=Rc, clamp_delta=abs(Rc - x)), cert, ExpectedEffect("kappa", k, None)


# --------------------------------------------------------------------------------------------------
# a9 map bin statistics
# --------------------------------------------------------------------------------------------------
def create_empty_map_stats(n_bins: int = constants.GC_B_BINS) -> MapBinStats:
    io = _IO()
    return MapBinStats(S_dir=io.zeros(n_bins, 3), S_dir_scatter=io.zeros(n_bins, 3, 3), N_dir=io.zeros(n_bins),
                       N_pos=io.zeros(n_bins), sum_p=io.zeros(n_bins, 3), sum_ppT=io.zeros(n_bins, 3, 3))


def map_stats_from_arrays(d) -> MapBinStats:
    io = _IO()
    return MapBinStats(S_dir=io.dev_in(d["S_dir"]), S_dir_scatter=io.dev_in(d["S_dir_scatter"]),
                       N_dir=io.dev_in(d["N_dir"]), N_pos=io.dev_in(d["N_pos"]), sum_p=io.dev_in(d["sum_p"]),
                       sum_ppT=io.dev_in(d["sum_ppT"]))


def update_map_stats(map_stats: MapBinStats, increments_S_dir, increments_S_dir_scatter, increments_N_dir,
                     increments_N_pos, increments_sum_p, increments_sum_ppT) -> MapBinStats:
    """Additive update (archive/bin_atlas.py:137-165).  Elementwise adds on six ~3 KB arrays: torch ops on device."""
    io = _IO()
    g = lambda x: io.dev_in(x)  # noqa: E731
    return MapBinStats(S_dir=map_stats.S_dir + g(increments_S_dir),
                       S_dir_scatter=map_stats.S_dir_scatter + g(increments_S_dir_scatter),
                       N_dir=map_stats.N_dir + g(increments_N_dir), N_pos=map_stats.N_pos + g(increments_N_pos),
                       sum_p=map_stats.sum_p + g(increments_sum_p), sum_ppT=map_stats.sum_ppT + g(increments_sum_ppT))


def apply_forgetting(map_stats: MapBinStats, forgetting_factor: float = 0.99) -> MapBinStats:
    g = float(forgetting_factor)
    return MapBinStats(S_dir=g * map_stats.S_dir, S_dir_scatter=g * map_stats.S_dir_scatter, N_dir=g * map_stats.N_dir,
                       N_pos=g * map_stats.N_pos, sum_p=g * map_stats.sum_p, sum_ppT=g * map_stats.sum_ppT)


def compute_map_derived_stats(map_stats: MapBinStats, eps_mass: float = constants.GC_EPS_MASS,
                              eps_psd: float = constants.GC_EPS_PSD):
    """-> (mu_dir, kappa, centroid, Sigma_c)   (archive/bin_atlas.py:200-221)."""
    io = _IO()
    B = map_stats.N_dir.shape[0]
    mu, kap, cen, Sc = io.empty(B, 3), io.empty(B), io.empty(B, 3), io.empty(B, 3, 3)
    m = map_stats._c()
    io.ctx.check(io.ctx.lib.gcs_map_bin_derived(io.ctx.handle, io.stream(), C.byref(m), B, float(eps_mass),
                                                float(eps_psd), L.ptr(mu), L.ptr(kap), L.ptr(cen), L.ptr(Sc)))
    return mu, kap, cen, Sc


def map_update_from_scan(map_stats: MapBinStats, scan: ScanBinStats, pose_start_of_scan, planar_z: bool = True,
                         forgetting_factor: float = 0.99) -> MapBinStats:
    """
    In place: map <- gamma * (map + pushforward(scan stats with the START-of-scan pose, t_z forced to 0)).
    Successor of the deleted PoseCovInflationPushforward (README.md:119; CHANGELOG.md:575-578,684-721): parity
    for this step is pinned only to the prose and to update_map_stats/apply_forgetting.
    """
    io = _IO()
    if scan.sum_p is None or scan.sum_ppT is None:
        raise ValueError("map_update_from_scan needs ScanBinStats.sum_p / sum_ppT (returned by this package)")
    B = map_stats.N_dir.shape[0]
    m = map_stats._c()
    pose = _host_vec(pose_start_of_scan, 6)
    # hold the (possibly re-laid-out) tensors until the call returns: ctypes only sees raw addresses
    sN, ssd, sS, ssp, sspp = (scan.N.contiguous(), scan.s_dir.contiguous(), scan.S_dir_scatter.contiguous(),
                              scan.sum_p.contiguous(), scan.sum_ppT.contiguous())
    io.ctx.check(io.ctx.lib.gcs_map_bin_update(
        io.ctx.handle, io.stream(), C.byref(m), L.ptr(sN), L.ptr(ssd), L.ptr(sS), L.ptr(ssp), L.ptr(sspp), B,
        _dptr(pose), 1 if planar_z else 0, float(forgetting_factor)))
    del sN, ssd, sS, ssp, sspp
    return map_stats


# --------------------------------------------------------------------------------------------------
# a7 / a8 evidence
# --------------------------------------------------------------------------------------------------
def _pose_of(belief_pred, eps_lift):
    if hasattr(belief_pred, "mean_world_pose"):
        pose = belief_pred.mean_world_pose(eps_lift=eps_lift)
        return _host_vec(pose, 6), belief_pred.chart_id, belief_pred.anchor_id
    return _host_vec(belief_pred, 6), constants.GC_CHART_ID, "initial"


def _metrics(rec, off):
    m = rec[off:off + 17]
    return ScatterMetrics(eigenvalues=torch.from_numpy(m[0:3].copy()), eigenvectors=torch.from_numpy(m[3:12].reshape(3, 3).copy()),
                          linearity=float(m[12]), planarity=float(m[13]), sphericity=float(m[14]),
                          anisotropy=float(m[15]), effective_rank=float(m[16]))


def _mf_from_record(rec_d, rec, io, chart_id, anchor_id):
    E = L.EV
    result = MatrixFisherResult(
        R_mf=rec_d[E["R_MF"]:E["R_MF"] + 9].reshape(3, 3), L_rot=rec_d[E["L_ROT"]:E["L_ROT"] + 9].reshape(3, 3),
        h_rot=rec_d[E["H_ROT"]:E["H_ROT"] + 3], delta_rot=rec_d[E["DELTA_ROT"]:E["DELTA_ROT"] + 3],
        svd_singular_values=rec_d[E["SVD_S"]:E["SVD_S"] + 3],
        map_scatter_metrics=_metrics(rec, E["MAP_METRICS"]), scan_scatter_metrics=_metrics(rec, E["SCAN_METRICS"]))
    cert = CertBundle.create_approx(
        chart_id=chart_id, anchor_id=anchor_id, triggers=["MatrixFisherRotationEvidence"],
        conditioning=ConditioningCert(eig_min=float(rec[E["MF_EIG_MIN"]]), eig_max=float(rec[E["MF_EIG_MAX"]]),
                                      cond=float(rec[E["MF_COND"]]), near_null_count=int(rec[E["MF_NEAR_NULL"]])),
        mismatch=MismatchCert(nll_per_ess=float(rec[E["MF_NLL_PER_ESS"]]), directional_score=float(rec[E["MF_DIR_SCORE"]])),
        influence=InfluenceCert(lift_strength=0.0, psd_projection_delta=float(rec[E["MF_PSD_DELTA"]]),
                                mass_epsilon_ratio=float(rec[E["MF_MASS_EPS"]]), anchor_drift_rho=0.0, dt_scale=1.0,
                                extrinsic_scale=1.0, trust_alpha=1.0),
        compute=io.compute(psd_projection_count=1))
    return result, cert, ExpectedEffect("predicted_rotation_nll", float(rec[E["MF_ROT_NLL"]]), None)


def _pt_from_record(rec_d, rec, io, chart_id, anchor_id):
    E = L.EV
    result = PlanarTranslationResult(
        t_wls=rec_d[E["T_WLS"]:E["T_WLS"] + 3], L_trans=rec_d[E["L_TRANS"]:E["L_TRANS"] + 9].reshape(3, 3),
        h_trans=rec_d[E["H_TRANS"]:E["H_TRANS"] + 3], delta_trans=rec_d[E["DELTA_TRANS"]:E["DELTA_TRANS"] + 3],
        xy_info_scale=float(rec[E["XY_INFO"]]), z_info_scale=float(rec[E["Z_INFO"]]))
    cert = CertBundle.create_approx(
        chart_id=chart_id, anchor_id=anchor_id, triggers=["PlanarTranslationEvidence"],
        conditioning=ConditioningCert(eig_min=float(rec[E["PT_EIG_MIN"]]), eig_max=float(rec[E["PT_EIG_MAX"]]),
                                      cond=float(rec[E["PT_COND"]]), near_null_count=int(rec[E["PT_NEAR_NULL"]])),
        mismatch=MismatchCert(nll_per_ess=float(rec[E["PT_NLL_PER_ESS"]]), directional_score=0.0),
        influence=InfluenceCert(lift_strength=0.0, psd_projection_delta=float(rec[E["PT_PSD_DELTA"]]),
                                mass_epsilon_ratio=float(rec[E["PT_MASS_EPS"]]), anchor_drift_rho=0.0, dt_scale=1.0,
                                extrinsic_scale=1.0, trust_alpha=1.0),
        compute=io.compute(psd_projection_count=1))
    return result, cert, ExpectedEffect("predicted_translation_nll", float(rec[E["PT_TRANS_NLL"]]), None)


def matrix_fisher_rotation_evidence(belief_pred, scan_s_dir, scan_S_dir_scatter, scan_N, map_S_dir, map_S_dir_scatter,
                                    map_N_dir, eps_psd: float = constants.GC_EPS_PSD,
                                    eps_lift: float = constants.GC_EPS_LIFT, eps_mass: float = constants.GC_EPS_MASS
                                    ) -> Tuple[MatrixFisherResult, CertBundle, ExpectedEffect]:
    """``belief_pred``: anything with mean_world_pose()/chart_id/anchor_id, or a 6-vector [t, rotvec]."""
    io = _IO()
    pose, chart_id, anchor_id = _pose_of(belief_pred, eps_lift)
    ssd, sS, sN = io.dev_in(scan_s_dir, shape=(-1, 3)), io.dev_in(scan_S_dir_scatter, shape=(-1, 3, 3)), io.dev_in(scan_N, shape=(-1,))
    msd, mS, mN = io.dev_in(map_S_dir, shape=(-1, 3)), io.dev_in(map_S_dir_scatter, shape=(-1, 3, 3)), io.dev_in(map_N_dir, shape=(-1,))
    B = sN.shape[0]
    if not (ssd.shape[0] == sS.shape[0] == msd.shape[0] == mS.shape[0] == mN.shape[0] == B):
        raise ValueError("matrix_fisher_rotation_evidence: per-bin arrays disagree in length")
    rec_d = io.zeros(L.EV_NREC)
    io.ctx.check(io.ctx.lib.gcs_matrix_fisher_rotation(
        io.ctx.handle, io.stream(), L.ptr(ssd), L.ptr(sS), L.ptr(sN), L.ptr(msd), L.ptr(mS), L.ptr(mN), B, _dptr(pose),
        float(eps_psd), float(eps_mass), L.ptr(rec_d)))
    return _mf_from_record(rec_d, io.host(rec_d), io, chart_id, anchor_id)


def planar_translation_evidence(belief_pred, scan_p_bar, scan_Sigma_p, scan_N, map_centroid, map_Sigma_c, map_N_pos,
                                map_S_dir_scatter, map_N_dir, R_hat, eps_psd: float = constants.GC_EPS_PSD,
                                eps_lift: float = constants.GC_EPS_LIFT, eps_mass: float = constants.GC_EPS_MASS
                                ) -> Tuple[PlanarTranslationResult, CertBundle, ExpectedEffect]:
    io = _IO()
    pose, chart_id, anchor_id = _pose_of(belief_pred, eps_lift)
    pb, Sp, sN = io.dev_in(scan_p_bar, shape=(-1, 3)), io.dev_in(scan_Sigma_p, shape=(-1, 3, 3)), io.dev_in(scan_N, shape=(-1,))
    mc, mSc, mNp = io.dev_in(map_centroid, shape=(-1, 3)), io.dev_in(map_Sigma_c, shape=(-1, 3, 3)), io.dev_in(map_N_pos, shape=(-1,))
    mS, mNd = io.dev_in(map_S_dir_scatter, shape=(-1, 3, 3)), io.dev_in(map_N_dir, shape=(-1,))
    B = sN.shape[0]
    Rh = _host_vec(R_hat, 9)
    if isinstance(R_hat, torch.Tensor) and R_hat.is_cuda:
        io.d2h += 72
        io.syncs += 1
    rec_d = io.zeros(L.EV_NREC)
    io.ctx.check(io.ctx.lib.gcs_planar_translation(
        io.ctx.handle, io.stream(), L.ptr(pb), L.ptr(Sp), L.ptr(sN), L.ptr(mc), L.ptr(mSc), L.ptr(mNp), L.ptr(mS),
        L.ptr(mNd), B, _dptr(Rh), _dptr(pose[:3]), float(eps_psd), float(eps_mass), L.ptr(rec_d)))
    return _pt_from_record(rec_d, io.host(rec_d), io, chart_id, anchor_id)


def build_combined_lidar_evidence_22d(mf_result: MatrixFisherResult, trans_result: PlanarTranslationResult):
    """22-D embedding (:729-756): translation block [0:3,0:3], rotation block [3:6,3:6]."""
    dev = mf_result.L_rot.device
    Lm = torch.zeros((constants.GC_D_Z, constants.GC_D_Z), dtype=F64, device=dev)
    h = torch.zeros(constants.GC_D_Z, dtype=F64, device=dev)
    Lm[0:3, 0:3] = trans_result.L_trans
    h[0:3] = trans_result.h_trans
    Lm[3:6, 3:6] = mf_result.L_rot
    h[3:6] = mf_result.h_rot
    return Lm, h


# --------------------------------------------------------------------------------------------------
# fused path
# --------------------------------------------------------------------------------------------------
@dataclass
class BinEvidenceBatch:
    """Outputs of the fused bin path for U = n_scans * n_hyp units (unit u = scan * n_hyp + hyp)."""
    n_scans: int
    n_hyp: int
    cap: int
    stats: dict               # N (U,B), s_dir (U,B,3), S_scatter, p_bar, Sigma_p, kappa, sum_p, sum_ppT
    evidence: Optional[torch.Tensor]   # (U, EV_NREC) packed MatrixFisher + planar-translation record
    L22: Optional[torch.Tensor]        # (U,22,22)
    h22: Optional[torch.Tensor]        # (U,22)
    cert: torch.Tensor                 # (U, BC_NCERT) packed certificate scalars (device)
    resampled: Optional[dict] = None   # points (S,cap,3), timestamps, weights, ring, tag
    deskewed: Optional[dict] = None    # points (U,cap,3), weights (U,cap)
    responsibilities: Optional[torch.Tensor] = None  # (U,cap,B)


class BinPathPlan:
    """
    Pre-marshalled launch of the fused bin path for fixed shapes: device buffers and the argument block are
    built once, ``run()`` only enqueues kernels (4 launches, no allocation, no host sync).
    """

    def __init__(self, n_scans, n_raw, cap, n_hyp=1, n_bins=constants.GC_B_BINS, tau=constants.GC_TAU_SOFT_ASSIGN,
                 origin=(0.0, 0.0, 0.0), precision=L.PREC_F64, want_evidence=True, materialize_resampled=False,
                 materialize_deskewed=True, materialize_responsibilities=False, device=None,
                 shard_row0=0, n_raw_total=0, cap_total=0, own_context=False):
        # own_context: a private gcs_ctx (workspace) so that this plan may run on its own stream concurrently with
        # other plans -- a gcs_ctx is not re-entrant (include/gcs_b200.h).
        if own_context:
            if device is None:
                device = torch.cuda.current_device()
            io = self.io = _IO(ctx=L.Context(device.index if isinstance(device, torch.device) else int(device)))
        else:
            io = self.io = _IO(device)
        self.S, self.H, self.U, self.B = int(n_scans), int(n_hyp), int(n_scans) * int(n_hyp), int(n_bins)
        self.n_raw, self.cap = int(n_raw), int(cap)
        S, U, B = self.S, self.U, self.B
        self.pts, self.t, self.w = io.empty(S, n_raw, 3), io.empty(S, n_raw), io.empty(S, n_raw)
        self.ring, self.tag = io.zeros(S, n_raw, dtype=torch.uint8), io.zeros(S, n_raw, dtype=torch.uint8)
        self.t0, self.t1 = io.zeros(S), io.zeros(S)
        self.xi, self.poses = io.zeros(U, 6), io.zeros(U, 6)
        self.bin_dirs = io.zeros(B, 3)
        self.bin_norm_max = 1.0
        self.map = create_empty_map_stats(B)
        self.stats, cstats = _alloc_stats(io, U, B)
        self.evidence = io.zeros(U, L.EV_NREC) if want_evidence else None
        self.L22 = io.zeros(U, 22, 22) if want_evidence else None
        self.h22 = io.zeros(U, 22) if want_evidence else None
        self.cert = io.zeros(U, L.BC_NCERT)
        self.rs = None
        if materialize_resampled:
            self.rs = dict(points=io.empty(S, cap, 3), timestamps=io.empty(S, cap), weights=io.empty(S, cap),
                           ring=io.empty(S, cap, dtype=torch.uint8), tag=io.empty(S, cap, dtype=torch.uint8))
        self.dk = dict(points=io.empty(U, cap, 3), weights=io.empty(U, cap)) if materialize_deskewed else None
        self.resp = io.empty(U, cap, B) if materialize_responsibilities else None
        self._map_c = self.map._c()
        a = self.args = L.BinsArgs()
        a.pts, a.t, a.w, a.ring, a.tag = L.ptr(self.pts), L.ptr(self.t), L.ptr(self.w), L.ptr(self.ring), L.ptr(self.tag)
        a.n_raw, a.cap, a.n_scans, a.n_hyp = self.n_raw, self.cap, S, self.H
        a.scan_t0, a.scan_t1, a.xi, a.poses = L.ptr(self.t0), L.ptr(self.t1), L.ptr(self.xi), L.ptr(self.poses)
        a.bin_dirs, a.n_bins, a.precision = L.ptr(self.bin_dirs), B, int(precision)
        a.origin = (C.c_double * 3)(*[float(v) for v in origin])
        a.tau, a.eps_mass, a.eps_psd = float(tau), constants.GC_EPS_MASS, constants.GC_EPS_PSD
        a.map = C.pointer(self._map_c) if want_evidence else None
        a.shard_row0, a.n_raw_total, a.cap_total, a.bin_norm_max = int(shard_row0), int(n_raw_total), int(cap_total), 1.0
        if self.rs:
            a.rs_pts, a.rs_t, a.rs_w = L.ptr(self.rs["points"]), L.ptr(self.rs["timestamps"]), L.ptr(self.rs["weights"])
            a.rs_ring, a.rs_tag = L.ptr(self.rs["ring"]), L.ptr(self.rs["tag"])
        if self.dk:
            a.dk_pts, a.dk_w = L.ptr(self.dk["points"]), L.ptr(self.dk["weights"])
        a.resp = L.ptr(self.resp)
        a.stats = cstats
        a.evidence, a.L22, a.h22, a.cert = L.ptr(self.evidence), L.ptr(self.L22), L.ptr(self.h22), L.ptr(self.cert)
        self.raw_len = int(io.ctx.lib.gcs_bins_raw_sums_len(B))

    # -- inputs ------------------------------------------------------------------------------------
    def set_bins(self, bin_dirs, tau=None):
        b = np.asarray(bin_dirs.detach().cpu().numpy() if isinstance(bin_dirs, torch.Tensor) else bin_dirs, np.float64)
        self.bin_dirs.copy_(torch.from_numpy(np.ascontiguousarray(b)))
        self.args.bin_norm_max = float(np.max(np.linalg.norm(b, axis=1)))
        if tau is not None:
            self.args.tau = float(tau)

    def set_map(self, map_stats):
        src = map_stats if isinstance(map_stats, dict) else map_stats.__dict__
        for k_dst, k_src in (("S_dir", "S_dir"), ("S_dir_scatter", "S_dir_scatter"), ("N_dir", "N_dir"),
                             ("N_pos", "N_pos"), ("sum_p", "sum_p"), ("sum_ppT", "sum_ppT")):
            v = src[k_src]
            v = v if isinstance(v, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(v, dtype=np.float64))
            getattr(self.map, k_dst).copy_(v)

    def upload(self, pts, t, w, ring=None, tag=None, t0=None, t1=None, xi=None, poses=None, non_blocking=True):
        """
        Host (ideally pinned) -> device copies of one batch; returns bytes copied.  The few-hundred-byte parameter copies
        go FIRST and the large arrays last: a small copy queued behind a large one on the same stream held back the
        kernels of the *other* stream of a double-buffered loop until the large copy had finished (measured: 4.06 ms per
        184.6 MB step with the small copies last, 3.42 ms with them first; tools/e2e_probe.py).
        """
        n = 0
        for dst, src in ((self.t0, t0), (self.t1, t1), (self.xi, xi), (self.poses, poses), (self.ring, ring), (self.tag, tag),
                         (self.t, t), (self.w, w), (self.pts, pts)):
            if src is None:
                continue
            src_t = src if isinstance(src, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(src))
            dst.copy_(src_t.reshape(dst.shape), non_blocking=non_blocking)
            n += dst.numel() * dst.element_size()
        return n

    def set_twist_from_imu(self, scan_index, imu_stamps, imu_gyro, imu_accel, scan_start_time, scan_end_time, sigma_warp,
                           rotvec_start_WB, gyro_bias, accel_bias, gravity_W=None, deskew_rotation_only=False):
        """
        SURVEY.md 8f-2: derive the n_hyp constant twists of scan `scan_index` from its IMU buffer on the device
        (window weights -> preintegration -> se3_log, fl/backend/pipeline.py:436-483) and write them straight into this
        plan's xi rows -- no host round trip between the IMU prologue and the deskew.  Returns the ImuTwistResult.
        """
        from . import imu
        s = int(scan_index)
        if not 0 <= s < self.S:
            raise ValueError(f"scan_index {s} outside [0, {self.S})")
        return imu.imu_scan_twist(imu_stamps, imu_gyro, imu_accel, scan_start_time, scan_end_time, sigma_warp,
                                  rotvec_start_WB, gyro_bias, accel_bias, gravity_W, deskew_rotation_only,
                                  xi_out=self.xi[s * self.H:(s + 1) * self.H])

    def enable_pointcloud2(self, fields, point_step: int, R_base_lidar=None, t_base_lidar=None):
        """
        Let the plan ingest PointCloud2 payloads directly: allocates the device staging buffer for
        n_scans x n_raw points of `point_step` bytes; upload_pointcloud2() then moves the wire bytes (22 B/point for the
        VLP-16 driver layout instead of 42 B/point of decoded arrays) and decodes them on the device.
        """
        io = self.io
        self._pc2_lay = pc2_layout(fields, point_step)
        nbytes = self.S * self.n_raw * int(point_step)
        self._pc2_dev = io.empty(nbytes + ((-nbytes) % 16), dtype=torch.uint8)
        self._pc2_stamp = io.zeros(self.S)
        self._pc2_cert = io.zeros(self.S, L.PC_NCERT)
        self._pc2_R = None if R_base_lidar is None else (C.c_double * 9)(*np.asarray(R_base_lidar, np.float64).reshape(9))
        self._pc2_t = None if t_base_lidar is None else (C.c_double * 3)(*np.asarray(t_base_lidar, np.float64).reshape(3))
        return nbytes

    def upload_pointcloud2(self, payload, header_stamps=None, t0=None, t1=None, xi=None, poses=None, non_blocking=True):
        """payload: uint8 host tensor (ideally pinned) holding the n_scans messages back to back.  Returns bytes copied."""
        io = self.io
        src = payload if isinstance(payload, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(payload, dtype=np.uint8))
        src = src.reshape(-1)
        n = self.S * self.n_raw * self._pc2_lay.point_step
        if src.numel() < n:
            raise ValueError(f"payload has {src.numel()} bytes, the plan needs {n}")
        moved = n
        # parameter copies first, payload last (see upload())
        for dst, s_ in ((self._pc2_stamp, header_stamps), (self.t0, t0), (self.t1, t1), (self.xi, xi), (self.poses, poses)):
            if s_ is None:
                continue
            st = s_ if isinstance(s_, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(s_, dtype=np.float64))
            dst.copy_(st.reshape(dst.shape), non_blocking=non_blocking)
            moved += dst.numel() * dst.element_size()
        self._pc2_dev[:n].copy_(src[:n], non_blocking=non_blocking)
        io.ctx.check(io.ctx.lib.gcs_parse_pointcloud2_vlp16(
            io.ctx.handle, io.stream(), L.ptr(self._pc2_dev), self.S, self.n_raw, C.byref(self._pc2_lay), L.ptr(self._pc2_stamp),
            self._pc2_R, self._pc2_t, L.ptr(self.pts), L.ptr(self.t), L.ptr(self.w), L.ptr(self.ring), L.ptr(self.tag),
            L.ptr(self._pc2_cert)))
        return moved

    # -- launches ----------------------------------------------------------------------------------
    def capture(self):
        """
        Capture run() (mass, memset, scan, reduce, finalize: seven small launches) into a CUDA graph; replay() then
        costs one graph launch.  The buffers of the plan are fixed, so upload() / upload_pointcloud2() between replays
        feed new scans.  Call after set_bins / set_map / the first upload.
        """
        self.run()                       # reserves the workspace and sets kernel attributes outside the capture
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            self.run()
        self._graph = g
        return self

    def replay(self):
        self._graph.replay()

    def run(self):
        io = self.io
        io.ctx.check(io.ctx.lib.gcs_lidar_evidence_bins(io.ctx.handle, io.stream(), C.byref(self.args)))

    def run_mass(self, mass):
        io = self.io
        io.ctx.check(io.ctx.lib.gcs_bins_mass(io.ctx.handle, io.stream(), C.byref(self.args), L.ptr(mass)))

    def run_accumulate(self, mass, raw_sums, raw_max):
        io = self.io
        io.ctx.check(io.ctx.lib.gcs_bins_accumulate(io.ctx.handle, io.stream(), C.byref(self.args), L.ptr(mass),
                                                    L.ptr(raw_sums), L.ptr(raw_max)))

    def run_finalize(self, mass, raw_sums, raw_max):
        io = self.io
        io.ctx.check(io.ctx.lib.gcs_bins_finalize(io.ctx.handle, io.stream(), C.byref(self.args), L.ptr(mass),
                                                  L.ptr(raw_sums), L.ptr(raw_max)))

    def outputs(self) -> BinEvidenceBatch:
        return BinEvidenceBatch(n_scans=self.S, n_hyp=self.H, cap=self.cap, stats=self.stats, evidence=self.evidence,
                                L22=self.L22, h22=self.h22, cert=self.cert, resampled=self.rs, deskewed=self.dk,
                                responsibilities=self.resp)


def lidar_evidence_bins(raw_points, raw_timestamps, raw_weights, raw_ring, raw_tag, n_points_cap, xi_body,
                        scan_start_time, scan_end_time, lidar_origin_base, bin_atlas, tau, map_stats, pose_pred,
                        precision: int = L.PREC_F64, materialize_responsibilities: bool = False,
                        chart_id: str = constants.GC_CHART_ID, anchor_id: str = "initial"):
    """
    Fused bin-family path for ONE scan and H >= 1 hypotheses: PointBudgetResample -> DeskewConstantTwist ->
    ray directions -> BinSoftAssign -> ScanBinMomentMatch (+kappa) -> MatrixFisherRotation ->
    PlanarTranslationEvidence -> 22-D LiDAR evidence.  Returns a dict with the same per-operator
    ``(result, CertBundle, ExpectedEffect)`` tuples the separate operators return (lists over hypotheses where the
    operator depends on the hypothesis), plus ``L`` (H,22,22) and ``h`` (H,22).
    """
    xi = np.asarray(xi_body.detach().cpu().numpy() if isinstance(xi_body, torch.Tensor) else xi_body, np.float64).reshape(-1, 6)
    poses = np.asarray(pose_pred.detach().cpu().numpy() if isinstance(pose_pred, torch.Tensor) else pose_pred,
                       np.float64).reshape(-1, 6)
    H = xi.shape[0]
    if poses.shape[0] != H:
        raise ValueError(f"xi_body has {H} hypotheses but pose_pred has {poses.shape[0]}")
    pts = np.asarray(raw_points, np.float64).reshape(-1, 3) if not isinstance(raw_points, torch.Tensor) else raw_points.reshape(-1, 3)
    n_raw = pts.shape[0]
    dirs = bin_atlas.dirs if isinstance(bin_atlas, BinAtlas) else bin_atlas
    B = dirs.shape[0]
    plan = BinPathPlan(1, n_raw, int(n_points_cap), n_hyp=H, n_bins=B, tau=tau, origin=_host_vec(lidar_origin_base, 3),
                       precision=precision, want_evidence=True, materialize_resampled=True, materialize_deskewed=True,
                       materialize_responsibilities=materialize_responsibilities)
    io = plan.io
    plan.set_bins(dirs)
    plan.set_map(map_stats)
    ring = np.zeros(n_raw, np.uint8) if raw_ring is None else raw_ring
    tag = np.zeros(n_raw, np.uint8) if raw_tag is None else raw_tag
    io.h2d += plan.upload(pts, raw_timestamps, raw_weights, ring, tag, np.array([scan_start_time], np.float64),
                          np.array([scan_end_time], np.float64), xi, poses, non_blocking=False)
    plan.run()
    out = plan.outputs()
    cert_h = io.host(torch.cat([out.cert.reshape(-1), out.evidence.reshape(-1)]))
    bc = cert_h[: H * L.BC_NCERT].reshape(H, L.BC_NCERT)
    ev = cert_h[H * L.BC_NCERT:].reshape(H, L.EV_NREC)
    cap = int(n_points_cap)
    stride = max(1, int(math.ceil(n_raw / cap)))
    n_sel = (n_raw + stride - 1) // stride
    eps = constants.GC_EPS_MASS

    mass_in = float(bc[0, L.BC_RS_MASS_IN])
    rs = PointBudgetResult(points=out.resampled["points"][0], timestamps=out.resampled["timestamps"][0],
                           weights=out.resampled["weights"][0], ring=out.resampled["ring"][0], tag=out.resampled["tag"][0],
                           n_input=n_raw, n_output=n_sel, total_mass_in=mass_in, total_mass_out=mass_in)
    rs_cert = CertBundle.create_approx(
        chart_id=chart_id, anchor_id=anchor_id, triggers=["PointBudgetResample"],
        support=SupportCert(ess_total=float(bc[0, L.BC_RS_ESS]), support_frac=float(min(1.0, cap / (n_raw + eps)))),
        influence=InfluenceCert.identity().with_overrides(mass_epsilon_ratio=eps / (mass_in + eps)), compute=io.compute())
    res = dict(resample=(rs, rs_cert, ExpectedEffect("predicted_ess", float(bc[0, L.BC_RS_ESS]), None)),
               deskew=[], soft_assign=[], stats=[], matrix_fisher=[], planar_translation=[], L=out.L22, h=out.h22)
    for h in range(H):
        c = bc[h]
        dk = DeskewConstantTwistResult(points=out.deskewed["points"][h], timestamps=out.resampled["timestamps"][0],
                                       weights=out.deskewed["weights"][h], ess_imu=1.0)
        dk_cert = CertBundle.create_exact(
            chart_id=chart_id, anchor_id=anchor_id,
            support=SupportCert(ess_total=1.0, support_frac=float(c[L.BC_DK_SUM_W_OUT] / (c[L.BC_DK_SUM_W_IN] + eps))),
            influence=InfluenceCert.identity(), compute=io.compute())
        res["deskew"].append((dk, dk_cert, ExpectedEffect("deskew_variance_reduction_proxy", 0.0, None)))
        avg_ent = float(c[L.BC_SA_ENTROPY_SUM] / (float(cap) + eps))
        sa_cert = CertBundle.create_exact(
            chart_id=chart_id, anchor_id=anchor_id,
            support=SupportCert(ess_total=float(math.exp(avg_ent)), support_frac=float(c[L.BC_SA_MAX_RESP])),
            compute=io.compute(largest_tensor_shape=(cap, B)))
        sa = BinSoftAssignResult(responsibilities=out.responsibilities[h] if out.responsibilities is not None else None)
        res["soft_assign"].append((sa, sa_cert, ExpectedEffect("predicted_assignment_entropy", avg_ent, None)))
        s = out.stats
        st = ScanBinStats(N=s["N"][h], s_dir=s["s_dir"][h], S_dir_scatter=s["S_scatter"][h], p_bar=s["p_bar"][h],
                          Sigma_p=s["Sigma_p"][h], kappa_scan=s["kappa"][h], sum_p=s["sum_p"][h], sum_ppT=s["sum_ppT"][h])
        st_cert = CertBundle.create_approx(
            chart_id=chart_id, anchor_id=anchor_id, triggers=["ScanBinMomentMatch"],
            support=SupportCert(ess_total=float(c[L.BC_ST_ESS]), support_frac=float(c[L.BC_ST_SUPPORT_FRAC])),
            influence=InfluenceCert(lift_strength=0.0, psd_projection_delta=float(c[L.BC_ST_PSD_DELTA]),
                                    mass_epsilon_ratio=float(c[L.BC_ST_MASS_EPS_RATIO]), anchor_drift_rho=0.0,
                                    dt_scale=1.0, extrinsic_scale=1.0, trust_alpha=1.0),
            compute=io.compute(psd_projection_count=B))
        res["stats"].append((st, st_cert, ExpectedEffect("predicted_ess", float(c[L.BC_ST_ESS]), None)))
        res["matrix_fisher"].append(_mf_from_record(out.evidence[h], ev[h], io, chart_id, anchor_id))
        res["planar_translation"].append(_pt_from_record(out.evidence[h], ev[h], io, chart_id, anchor_id))
    return res
