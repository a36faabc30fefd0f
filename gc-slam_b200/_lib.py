"""
ctypes binding of the C-ABI in include/gcs_b200.h.

No fallback: if libgcs_b200.so is missing or no B200 is visible, this module (or ``context()``) raises.
torch is used only for device memory (tensors' ``data_ptr()``) and stream handles.
"""

from __future__ import annotations

import ctypes as C
import os
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
# GCS_B200_LIB: a developer override (tools/build_variant.py builds experimental variants of the library next to the real one)
LIB_PATH = os.environ.get("GCS_B200_LIB") or os.path.join(_HERE, "lib", "libgcs_b200.so")

GCS_OK, GCS_EINVAL, GCS_ECUDA, GCS_ENOMEM, GCS_ECOMM = 0, -1, -2, -3, -4
PREC_F64, PREC_MIXED, PREC_TC = 0, 1, 2

# certificate / record layouts (enums of the header)
RS_MASS_IN, RS_MASS_SEL, RS_SUMSQ_SEL, RS_ESS, RS_MASS_SCALE, RS_NCERT = 0, 1, 2, 3, 4, 8
DK_SUM_W_OUT, DK_SUM_W_IN, DK_NCERT = 0, 1, 4
SA_ENTROPY_SUM, SA_MAX_RESP, SA_NCERT = 0, 1, 4
ST_ESS, ST_SUPPORT_FRAC, ST_PSD_DELTA, ST_MASS_EPS_RATIO, ST_NCERT = 0, 1, 2, 3, 8
(BC_RS_MASS_IN, BC_RS_ESS, BC_RS_MASS_SCALE, BC_DK_SUM_W_OUT, BC_DK_SUM_W_IN, BC_SA_ENTROPY_SUM, BC_SA_MAX_RESP,
 BC_ST_ESS, BC_ST_SUPPORT_FRAC, BC_ST_PSD_DELTA, BC_ST_MASS_EPS_RATIO) = range(11)
BC_NCERT = 16
EV = dict(R_MF=0, L_ROT=9, H_ROT=18, DELTA_ROT=21, SVD_S=24, SCAN_METRICS=27, MAP_METRICS=44, T_WLS=61, L_TRANS=64,
          H_TRANS=73, DELTA_TRANS=76, XY_INFO=79, Z_INFO=80, Z_SCALE=81,
          MF_EIG_MIN=82, MF_EIG_MAX=83, MF_COND=84, MF_NEAR_NULL=85, MF_NLL_PER_ESS=86, MF_DIR_SCORE=87,
          MF_PSD_DELTA=88, MF_MASS_EPS=89, MF_ROT_NLL=90, MF_N_EFF=91,
          PT_EIG_MIN=92, PT_EIG_MAX=93, PT_COND=94, PT_NEAR_NULL=95, PT_NLL_PER_ESS=96, PT_PSD_DELTA=97,
          PT_MASS_EPS=98, PT_TRANS_NLL=99, PT_N_EFF=100)
EV_NREC = 104

TIME_TAGS = dict(bin_scan=0, surfel_fit=1, map_view=2, topk=3, fuse=4, inflate=5, sinkhorn=6)

_vp = C.c_void_p
_i64 = C.c_int64
_dbl = C.c_double
_int = C.c_int


class BinStats(C.Structure):
    _fields_ = [(n, _vp) for n in ("N", "s_dir", "S_scatter", "p_bar", "Sigma_p", "kappa", "sum_p", "sum_ppT")]


class MapBinStats(C.Structure):
    _fields_ = [(n, _vp) for n in ("S_dir", "S_scatter", "N_dir", "N_pos", "sum_p", "sum_ppT")]


class BinsArgs(C.Structure):
    _fields_ = [
        ("pts", _vp), ("t", _vp), ("w", _vp), ("ring", _vp), ("tag", _vp),
        ("n_raw", _i64), ("cap", _i64), ("n_scans", C.c_int32), ("n_hyp", C.c_int32),
        ("scan_t0", _vp), ("scan_t1", _vp), ("xi", _vp), ("poses", _vp),
        ("bin_dirs", _vp), ("n_bins", C.c_int32), ("precision", C.c_int32),
        ("origin", _dbl * 3), ("tau", _dbl), ("eps_mass", _dbl), ("eps_psd", _dbl),
        ("map", C.POINTER(MapBinStats)),
        ("shard_row0", _i64), ("n_raw_total", _i64), ("cap_total", _i64), ("bin_norm_max", _dbl),
        ("rs_pts", _vp), ("rs_t", _vp), ("rs_w", _vp), ("rs_ring", _vp), ("rs_tag", _vp),
        ("dk_pts", _vp), ("dk_w", _vp), ("resp", _vp),
        ("stats", BinStats),
        ("evidence", _vp), ("L22", _vp), ("h22", _vp), ("cert", _vp),
    ]


class Pc2Layout(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("point_step", "off_x", "type_x", "off_y", "type_y", "off_z", "type_z", "off_ring",
                                         "type_ring", "off_time", "type_time")]


PC_N_NONFINITE, PC_TIME_RESCALED, PC_NCERT = 0, 1, 4
HB = dict(FLOOR_ADJUSTMENT=0, SPREAD_PROXY=1, PSD_PROJECTION_DELTA=2, PSD_SYM_DELTA=3, PSD_EIG_MIN=4, PSD_EIG_MAX=5,
          PSD_COND=6, PSD_NEAR_NULL=7, NCERT=8)
IMU_NPARAM, IMU_NOUT = 16, 40
IMU_OFF = dict(delta_pose=(0, 6), delta_R=(6, 15), delta_p=(15, 18), delta_v=(18, 21), ess=(21, 22), a_body_mean=(22, 25),
               a_world_nog_mean=(25, 28), a_world_mean=(28, 31), dt_eff_sum=(31, 32), xi_body=(32, 38))

# name -> (restype, argtypes); every symbol declared in include/gcs_b200.h appears here
PROTOTYPES = {
    "gcs_version": (_int, []),
    "gcs_version_string": (C.c_char_p, []),
    "gcs_create": (_int, [C.POINTER(_vp), _int]),
    "gcs_destroy": (_int, [_vp]),
    "gcs_last_error": (C.c_char_p, [_vp]),
    "gcs_reserve_workspace": (_int, [_vp, C.c_uint64]),
    "gcs_workspace_freeze": (_int, [_vp, _int]),
    "gcs_workspace_bytes": (C.c_uint64, [_vp]),
    "gcs_side_route": (_int, [_vp, _int]),
    "gcs_side_join": (_int, [_vp, _vp]),
    "gcs_device_sm_count": (_int, [_vp]),
    "gcs_kernel_launches": (C.c_uint64, [_vp]),
    "gcs_timing_enable": (_int, [_vp, _int]),
    "gcs_timing_collect": (_int, [_vp, C.POINTER(_dbl), C.POINTER(_int)]),
    "gcs_parse_pointcloud2_vlp16": (_int, [_vp, _vp, _vp, _int, _i64, C.POINTER(Pc2Layout), _vp, C.POINTER(_dbl), C.POINTER(_dbl),
                                           _vp, _vp, _vp, _vp, _vp, _vp]),
    "gcs_imu_scan_twist": (_int, [_vp, _vp, _vp, _vp, _vp, _i64, _vp, _vp, _int, _vp, _vp, _vp]),
    "gcs_hypothesis_barycenter": (_int, [_vp, _vp, _vp, _vp, _vp, _vp, _int, _int, _dbl, _dbl, _dbl, _vp, _vp, _vp, _vp, _vp, _vp]),
    "gcs_point_budget_resample": (_int, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _i64, _dbl, _vp, _vp, _vp, _vp, _vp, _vp]),
    "gcs_deskew_constant_twist": (_int, [_vp, _vp, _vp, _vp, _vp, _i64, C.POINTER(_dbl), _dbl, _dbl, _vp, _vp, _vp]),
    "gcs_ray_directions": (_int, [_vp, _vp, _vp, _i64, C.POINTER(_dbl), _dbl, _vp]),
    "gcs_bin_soft_assign": (_int, [_vp, _vp, _vp, _i64, _vp, _int, _dbl, _dbl, _int, _vp, _vp]),
    "gcs_scan_bin_moment_match": (_int, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, C.POINTER(_dbl), _i64, _int, _dbl, _dbl,
                                         C.POINTER(BinStats), _vp]),
    "gcs_kappa_from_resultant_batch": (_int, [_vp, _vp, _vp, _i64, _dbl, _dbl, _dbl, _dbl, _vp]),
    "gcs_map_bin_update": (_int, [_vp, _vp, C.POINTER(MapBinStats), _vp, _vp, _vp, _vp, _vp, _int, C.POINTER(_dbl), _int, _dbl]),
    "gcs_map_bin_derived": (_int, [_vp, _vp, C.POINTER(MapBinStats), _int, _dbl, _dbl, _vp, _vp, _vp, _vp]),
    "gcs_bin_evidence": (_int, [_vp, _vp, C.POINTER(BinStats), _int, _int, C.POINTER(MapBinStats), _vp, _dbl, _dbl, _vp, _vp, _vp]),
    "gcs_matrix_fisher_rotation": (_int, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _int, C.POINTER(_dbl), _dbl, _dbl, _vp]),
    "gcs_planar_translation": (_int, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _int, C.POINTER(_dbl),
                                      C.POINTER(_dbl), _dbl, _dbl, _vp]),
    "gcs_bins_raw_sums_len": (_int, [_int]),
    "gcs_lidar_evidence_bins": (_int, [_vp, _vp, C.POINTER(BinsArgs)]),
    "gcs_bins_mass": (_int, [_vp, _vp, C.POINTER(BinsArgs), _vp]),
    "gcs_bins_accumulate": (_int, [_vp, _vp, C.POINTER(BinsArgs), _vp, _vp, _vp]),
    "gcs_bins_finalize": (_int, [_vp, _vp, C.POINTER(BinsArgs), _vp, _vp, _vp]),
    "gcs_bins_reduce_gathered": (_int, [_vp, _vp, _vp, C.c_int32, _i64, _i64, _vp, _vp]),
    "gcs_peer_xchg_create": (_int, [_vp, C.c_int32, C.c_int32, C.c_uint64, C.POINTER(_vp), _vp]),
    "gcs_peer_xchg_connect": (_int, [_vp, _vp, _vp]),
    "gcs_peer_xchg_reduce": (_int, [_vp, _vp, _vp, _vp, _i64, _i64]),
    "gcs_peer_xchg_status": (_int, [_vp, _vp, C.POINTER(C.c_uint32)]),
    "gcs_peer_xchg_destroy": (_int, [_vp, _vp]),
}

_lib = None
_lib_lock = threading.Lock()


def register_prototypes(extra):
    """Other modules (primitive family) add their entry points here before the library is first loaded."""
    PROTOTYPES.update(extra)
    global _lib
    if _lib is not None:
        for name, (res, args) in extra.items():
            fn = getattr(_lib, name)
            fn.restype, fn.argtypes = res, args


def load():
    """dlopen the library (works without a GPU; nothing CUDA runs until gcs_create)."""
    global _lib
    with _lib_lock:
        if _lib is None:
            if not os.path.exists(LIB_PATH):
                raise RuntimeError(
                    f"{LIB_PATH} is missing: build it with `python -m gc_slam_b200.build` "
                    "(nvcc, sm_100a).  gc_slam_b200 has no CPU fallback.")
            lib = C.CDLL(LIB_PATH)
            for name, (res, args) in PROTOTYPES.items():
                fn = getattr(lib, name)  # AttributeError here = header/library mismatch
                fn.restype, fn.argtypes = res, args
            _lib = lib
    return _lib


class GcsError(RuntimeError):
    pass


class Context:
    """One gcs_ctx: bound to a device, not re-entrant (one per calling thread / stream)."""

    def __init__(self, device: int = 0):
        self.lib = load()
        h = _vp()
        rc = self.lib.gcs_create(C.byref(h), int(device))
        if rc != GCS_OK:
            msg = self.lib.gcs_last_error(None).decode()
            raise RuntimeError(f"gcs_create(device={device}) failed [{rc}]: {msg}")
        self.handle = h
        self.device = int(device)

    def check(self, rc: int):
        if rc == GCS_OK:
            return
        msg = self.lib.gcs_last_error(self.handle).decode()
        if rc == GCS_EINVAL:
            raise ValueError(msg)
        raise GcsError(f"[gcs {rc}] {msg}")

    @property
    def launches(self) -> int:
        return int(self.lib.gcs_kernel_launches(self.handle))

    @property
    def sm_count(self) -> int:
        return int(self.lib.gcs_device_sm_count(self.handle))

    def reserve_workspace(self, nbytes: int):
        self.check(self.lib.gcs_reserve_workspace(self.handle, int(nbytes)))

    def freeze_workspace(self, frozen: bool = True):
        """After this, a call that would have to grow the workspace raises instead (steady state never allocates)."""
        self.check(self.lib.gcs_workspace_freeze(self.handle, 1 if frozen else 0))

    def side_route(self, on: bool):
        """While on, gcs_map_recency_inflate / gcs_map_update run on the context's side stream (include/gcs_b200.h)."""
        self.check(self.lib.gcs_side_route(self.handle, 1 if on else 0))

    def side_join(self, stream):
        """`stream` waits for everything enqueued on the context's side stream."""
        self.check(self.lib.gcs_side_join(self.handle, stream))

    @property
    def workspace_bytes(self) -> int:
        return int(self.lib.gcs_workspace_bytes(self.handle))

    def timing_enable(self, on=True, only: str = None):
        """Bracket the dominant kernels with CUDA events; `only`: one of TIME_TAGS (e.g. "topk") to time just that kernel."""
        code = 0 if not on else (1 if only is None else 100 + TIME_TAGS[only])
        self.check(self.lib.gcs_timing_enable(self.handle, code))

    def timing_collect(self):
        """-> (summed device ms of the dominant kernel, number of launches) since the last collect."""
        ms, n = _dbl(0.0), _int(0)
        self.check(self.lib.gcs_timing_collect(self.handle, C.byref(ms), C.byref(n)))
        return float(ms.value), int(n.value)

    def close(self):
        if getattr(self, "handle", None):
            self.lib.gcs_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


_tls = threading.local()


def context(device: int | None = None) -> Context:
    """Per-thread, per-device context (the reference calls operators from a worker thread and a 2-thread pool)."""
    import torch

    if not torch.cuda.is_available():
        raise RuntimeError("gc_slam_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
    if isinstance(device, torch.device):
        device = device.index
    if device is None:
        device = torch.cuda.current_device()
    device = int(device)
    cache = getattr(_tls, "ctx", None)
    if cache is None:
        cache = _tls.ctx = {}
    if device not in cache:
        cache[device] = Context(device)
    return cache[device]


def stream_ptr(device=None):
    import torch

    return _vp(torch.cuda.current_stream(device).cuda_stream)


def ptr(t):
    """Device pointer of a contiguous CUDA tensor (None -> NULL)."""
    if t is None:
        return _vp(0)
    if not t.is_cuda:
        raise ValueError("expected a CUDA tensor")
    if not t.is_contiguous():
        raise ValueError("expected a contiguous tensor")
    return _vp(t.data_ptr())


def version_string() -> str:
    return load().gcs_version_string().decode()
