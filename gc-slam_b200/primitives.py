"""
Primitive-family LiDAR evidence operators -- host-side mirror of the reference's Python interface
(``fl/`` = fl_ws/src/fl_slam_poc/fl_slam_poc/ in whabacivch/GC-SLAM):

  MeasurementBatch, measurement_batch_from_camera_splats, create_empty_measurement_batch
                                     fl/backend/structures/measurement_batch.py:68-259
  SurfelExtractionConfig, extract_lidar_surfels
                                     fl/backend/operators/lidar_surfel_extraction.py:45-431
  AtlasMap (device tile pool), AtlasMapView, extract_atlas_map_view, primitive_map_recency_inflate,
  primitive_map_cull, primitive_map_forget
                                     fl/backend/structures/primitive_map.py:98-1484
  AssociationConfig, PrimitiveAssociationResult, associate_primitives_ot
                                     fl/backend/operators/primitive_association.py:71-553
  VisualPoseEvidenceResult, visual_pose_evidence
                                     fl/backend/operators/visual_pose_evidence.py:50-436
  map_update_step12b                 fl/backend/pipeline.py:1233-1447 (fuse x blocks x tiles, insert, cull, forget)
  primitive_map_fuse, primitive_map_insert_masked, primitive_map_cull, primitive_map_forget (one tile per call)
                                     fl/backend/structures/primitive_map.py:807-1384
  block_associations_for_fuse        fl/backend/operators/primitive_association.py:561-588
  compute_sparse_cost_matrix, sinkhorn_unbalanced_fixed_k (the association's inner functions, called by the reference's
  start-up warm-up)                  fl/backend/operators/primitive_association.py:105-197
  ma_hex_stencil_tile_ids, tile_ids_from_xyz_batch   fl/common/tiling.py:126-209 (host integer helpers)

Every operator returns ``(result, CertBundle, ExpectedEffect)``; arrays are torch CUDA tensors.  All arithmetic on
arrays happens in libgcs_b200.so; the host code here only marshals pointers and assembles certificates.
"""

from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from enum import Enum
from typing import Dict, List, Optional, Tuple

import numpy as np
import torch

from . import _lib as L
from . import constants
from .certs import (CertBundle, ComputeCert, ExpectedEffect, InfluenceCert, MapUpdateCert, OTCert, SupportCert)
from .operators import _IO, _Pending, _deskew_constant_twist_gen, _dptr, _drive, _host_vec, drive_group

F64 = torch.float64
_vp, _i32, _i64, _dbl = C.c_void_p, C.c_int32, C.c_int64, C.c_double


# --------------------------------------------------------------------------------------------------
# C structs
# --------------------------------------------------------------------------------------------------
class CMeasBatch(C.Structure):
    _fields_ = [(n, _vp) for n in ("Lambdas", "thetas", "etas", "weights", "sources", "source_indices", "valid",
                                   "timestamps", "colors")] + [("n_feat", _i32), ("n_surfel", _i32)]


class CSurfelCfg(C.Structure):
    _fields_ = [(n, _i32) for n in ("n_cells_1", "n_cells_2", "n_cells_z", "max_occupants", "min_points_per_voxel")] + \
               [(n, _dbl) for n in ("voxel_size_m", "sensor_noise_var_per_axis", "wishart_nu", "wishart_psi_scale",
                                    "kappa_main_scale", "kappa_min", "kappa_max", "eig_min", "eps_lift")]


_ATLAS_FIELDS = ("Lambdas", "thetas", "etas", "weights", "timestamps", "created_timestamps", "last_supported_scan_seq",
                 "last_update_scan_seq", "primitive_ids", "valid", "colors", "cam_mass", "lidar_mass", "rgb_cam_accum",
                 "rgb_cam_denom", "rgb")


class CAtlas(C.Structure):
    _fields_ = [(n, _vp) for n in _ATLAS_FIELDS] + [("m_tile", _i32), ("n_tiles_cap", _i32)]


_VIEW_FIELDS = ("candidate_tile_ids", "candidate_slots", "valid", "positions", "covariances", "directions", "kappas",
                "weights", "primitive_ids", "last_supported_scan_seq", "etas", "colors")


class CMapView(C.Structure):
    _fields_ = [(n, _vp) for n in _VIEW_FIELDS]


class CMapExport(C.Structure):
    _fields_ = [(n, _vp) for n in ("mu_world", "Sigma_world", "Lambda_world", "eta", "mass", "color", "primitive_ids",
                                   "last_supported_scan_seq", "cloud")]


class CAssocCfg(C.Structure):
    _fields_ = [(n, _i32) for n in ("k_assoc", "k_sinkhorn", "r_stencil_xy", "r_stencil_z", "a_policy", "reserved_")] + \
               [(n, _dbl) for n in ("beta", "epsilon", "tau_a", "tau_b", "eps_mass", "eps_lift", "h_tile",
                                    "recency_decay_lambda")] + [("scan_seq", _i64)]


class CAssocResult(C.Structure):
    _fields_ = [(n, _vp) for n in ("responsibilities", "candidate_pool_indices", "candidate_tile_ids", "candidate_slots",
                                   "row_masses", "cost_matrix")]


class CMapUpdateCfg(C.Structure):
    _fields_ = [(n, _i32) for n in ("k_insert_tile", "k_assoc", "assoc_block_size", "strict_tile_state")] + \
               [(n, _dbl) for n in ("recency_decay_lambda", "eps_lift", "eps_mass", "h_tile", "cull_weight_threshold",
                                    "forgetting_factor")] + [("scan_seq", _i64), ("next_global_id", _i64), ("timestamp", _dbl),
                                                           ("next_global_id_dev", _vp)]


OT = dict(MARGINAL_A=0, MARGINAL_B=1, MASS_TOTAL=2, SUM_A=3, SUM_M=4, SUM_NOVEL=5, P95_A=6, NONZERO_A=7, B_RECENCY_P95=8,
          ESS=9, TOTAL_COST=10, SUM_M2=11, NCERT=16)
VP = dict(L_TRANS=0, H_TRANS=9, L_ROT=12, H_ROT=21, TRANS_COST=24, ROT_COST=25, SUM_ROW_MASS=26, N_VALID_ROWS=27, SVD_S=28,
          DELTA_ROT=31, R_SCATTER=34, NREC=48)
MU = dict(FUSED_COUNT=0, FUSED_MASS=1, INSERT_COUNT=2, INSERT_MASS=3, INSERT_MASS_P95=4, EVICTED_COUNT=5, EVICTED_MASS=6,
          NEXT_GLOBAL_ID=7, TILE_COUNT0=8, NSTATS=24)

_int = C.c_int
L.register_prototypes({
    "gcs_batch_from_camera_splats": (_int, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i32, _dbl, C.POINTER(CMeasBatch)]),
    "gcs_extract_lidar_surfels": (_int, [_vp, _vp, _vp, _vp, _vp, _i64, C.POINTER(CSurfelCfg), C.POINTER(CMeasBatch), _vp, _vp, _vp]),
    "gcs_map_recency_inflate": (_int, [_vp, _vp, C.POINTER(CAtlas), C.POINTER(_i32), _i32, _i64, _dbl, _dbl, _vp]),
    "gcs_extract_atlas_map_view": (_int, [_vp, _vp, C.POINTER(CAtlas), C.POINTER(_i32), C.POINTER(_i64), _i32, _i32, _dbl, _dbl,
                                          C.POINTER(CMapView), _vp]),
    "gcs_associate_primitives_ot": (_int, [_vp, _vp, C.POINTER(CMeasBatch), C.POINTER(CMapView), C.POINTER(_i64), _i32, _i32,
                                           C.POINTER(CAssocCfg), C.POINTER(CAssocResult), _vp]),
    "gcs_visual_pose_evidence": (_int, [_vp, _vp, C.POINTER(CMeasBatch), C.POINTER(CMapView), C.POINTER(CAssocResult), _i32,
                                        C.POINTER(_dbl), _dbl, _dbl, _vp, _vp, _vp]),
    "gcs_map_update": (_int, [_vp, _vp, C.POINTER(CAtlas), C.POINTER(_i32), C.POINTER(_i64), _i32, C.POINTER(CMeasBatch),
                              C.POINTER(CAssocResult), C.POINTER(_dbl), C.POINTER(CMapUpdateCfg), _vp, _vp, _vp]),
    "gcs_map_merge_reduce": (_int, [_vp, _vp, C.POINTER(CAtlas), _i32, _dbl, _i32, _dbl, _dbl, _vp]),
    "gcs_export_map_points": (_int, [_vp, _vp, C.POINTER(CAtlas), C.POINTER(_i32), _i32, _dbl, C.POINTER(CMapExport), _i64, _vp]),
    "gcs_map_fuse": (_int, [_vp, _vp, C.POINTER(CAtlas), _i32, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i32, _dbl, _i64, _dbl, _vp]),
    "gcs_map_insert_masked": (_int, [_vp, _vp, C.POINTER(CAtlas), _i32, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i32, _dbl, _i64, _dbl,
                                     _i64, _vp, _vp, _vp]),
    "gcs_map_cull": (_int, [_vp, _vp, C.POINTER(CAtlas), _i32, _dbl, _i32, _vp]),
    "gcs_map_forget": (_int, [_vp, _vp, C.POINTER(CAtlas), _i32, _dbl]),
    "gcs_sparse_cost_matrix": (_int, [_vp, _vp, _vp, _vp, _vp, _i32, _vp, _vp, _vp, _i32, _vp, _i32, _dbl, _dbl, _vp]),
    "gcs_sinkhorn_unbalanced_fixed_k": (_int, [_vp, _vp, _vp, _vp, _vp, _i32, _i32, _dbl, _dbl, _dbl, _i32, _vp]),
})


# --------------------------------------------------------------------------------------------------
# tiling helpers (host integers; the reference also evaluates these on the host: fl/common/tiling.py)
# --------------------------------------------------------------------------------------------------
_BITS, _BIAS = 21, 1 << 20


def tile_id_from_cell_3d(c1: int, c2: int, cz: int) -> int:
    m = (1 << _BITS) - 1
    return int((((int(c1) + _BIAS) & m) << (2 * _BITS)) | (((int(c2) + _BIAS) & m) << _BITS) | ((int(cz) + _BIAS) & m))


_HEX_E1 = np.array([1.0, 0.0])
_HEX_E2 = np.array([0.5, 0.5 * np.sqrt(3.0)])


def ma_hex_cell_3d_from_xyz(xyz, h_tile: float):
    xyz = np.asarray(xyz, dtype=np.float64).ravel()
    if xyz.shape[0] < 3:
        raise ValueError(f"ma_hex_cell_3d_from_xyz: expected xyz (3,), got shape {xyz.shape}")
    h = max(float(h_tile), 1e-12)
    s1 = float(_HEX_E1 @ xyz[:2])
    s2 = float(_HEX_E2 @ xyz[:2])
    return int(np.floor(s1 / h)), int(np.floor(s2 / h)), int(np.floor(float(xyz[2]) / h))


def ma_hex_cells_3d_from_xyz_batch(xyz, h_tile: float):
    """Cells of many centres at once: vectorised, except for centres within 1e-9 cells of a boundary, which take the
    scalar function (so that the decision is the scalar one's bit for bit where rounding could matter)."""
    xyz = np.asarray(xyz, dtype=np.float64).reshape(-1, 3)
    h = max(float(h_tile), 1e-12)
    f = np.stack([xyz[:, 0], xyz[:, :2] @ _HEX_E2, xyz[:, 2]], axis=1) / h
    cells = np.floor(f).astype(np.int64)
    near = np.any(np.abs(f - np.rint(f)) < 1e-9, axis=1) | ~np.all(np.isfinite(f), axis=1)
    out = list(map(tuple, cells.tolist()))
    for k in np.nonzero(near)[0]:
        out[int(k)] = ma_hex_cell_3d_from_xyz(xyz[int(k)], h_tile)
    return out


def hex_disk_axial(radius: int):
    r = int(radius)
    out = [(q, rr) for q in range(-r, r + 1) for rr in range(max(-r, -q - r), min(r, -q + r) + 1)]
    out.sort()
    return out


def ma_hex_stencil_tile_ids(center_xyz, h_tile: float = constants.GC_H_TILE, radius_xy: int = constants.GC_R_STENCIL_TILES_XY,
                            radius_z: int = constants.GC_R_STENCIL_TILES_Z) -> List[int]:
    return list(stencil_of_cell(ma_hex_cell_3d_from_xyz(center_xyz, h_tile), radius_xy, radius_z))


_STENCIL_CACHE: dict = {}


def stencil_of_cell(cell, radius_xy: int = constants.GC_R_STENCIL_TILES_XY, radius_z: int = constants.GC_R_STENCIL_TILES_Z) -> tuple:
    """Stencil tile ids around a cell (c1, c2, cz), in the reference's order (tiling.py:171-186); memoised (a robot stays in
    a cell for many scans, and the hypotheses of a scan share one or two cells)."""
    key = (int(cell[0]), int(cell[1]), int(cell[2]), int(radius_xy), int(radius_z))
    out = _STENCIL_CACHE.get(key)
    if out is None:
        c1, c2, cz = key[:3]
        out = tuple(tile_id_from_cell_3d(c1 + dq, c2 + dr, cz + dz) for dz in range(-int(radius_z), int(radius_z) + 1)
                    for dq, dr in hex_disk_axial(radius_xy))
        if len(_STENCIL_CACHE) > 4096:
            _STENCIL_CACHE.clear()
        _STENCIL_CACHE[key] = out
    return out


# --------------------------------------------------------------------------------------------------
# MeasurementBatch
# --------------------------------------------------------------------------------------------------
@dataclass
class MeasurementBatch:
    Lambdas: torch.Tensor
    thetas: torch.Tensor
    etas: torch.Tensor
    weights: torch.Tensor
    sources: torch.Tensor
    source_indices: torch.Tensor
    valid_mask: torch.Tensor  # uint8 on device (0/1)
    timestamps: torch.Tensor
    colors: torch.Tensor
    n_feat: int
    n_surfel: int
    n_camera_valid: int
    n_lidar_valid: int

    @property
    def n_total(self) -> int:
        return self.n_feat + self.n_surfel

    @property
    def n_valid(self) -> int:
        return self.n_camera_valid + self.n_lidar_valid

    @property
    def camera_slice(self) -> slice:
        return slice(0, self.n_feat)

    @property
    def lidar_slice(self) -> slice:
        return slice(self.n_feat, self.n_total)

    def _c(self) -> CMeasBatch:
        b = CMeasBatch()
        b.Lambdas, b.thetas, b.etas, b.weights = L.ptr(self.Lambdas), L.ptr(self.thetas), L.ptr(self.etas), L.ptr(self.weights)
        b.sources, b.source_indices, b.valid = L.ptr(self.sources), L.ptr(self.source_indices), L.ptr(self.valid_mask)
        b.timestamps, b.colors, b.n_feat, b.n_surfel = L.ptr(self.timestamps), L.ptr(self.colors), self.n_feat, self.n_surfel
        return b

    def clone(self) -> "MeasurementBatch":
        return MeasurementBatch(**{k: (v.clone() if isinstance(v, torch.Tensor) else v) for k, v in self.__dict__.items()})


def create_empty_measurement_batch(n_feat: int = constants.GC_N_FEAT, n_surfel: int = constants.GC_N_SURFEL, device=None
                                   ) -> MeasurementBatch:
    io = _IO(device)
    n = n_feat + n_surfel
    return MeasurementBatch(Lambdas=io.zeros(n, 3, 3), thetas=io.zeros(n, 3), etas=io.zeros(n, constants.GC_VMF_N_LOBES, 3),
                            weights=io.zeros(n), sources=io.zeros(n, dtype=torch.int32),
                            source_indices=io.zeros(n, dtype=torch.int32), valid_mask=io.zeros(n, dtype=torch.uint8),
                            timestamps=io.zeros(n), colors=io.zeros(n, 3), n_feat=int(n_feat), n_surfel=int(n_surfel),
                            n_camera_valid=0, n_lidar_valid=0)


def measurement_batch_from_camera_splats(positions, covariances, directions, kappas, weights, timestamps, colors=None,
                                         n_feat: int = constants.GC_N_FEAT, n_surfel: int = constants.GC_N_SURFEL,
                                         eps_lift: float = constants.GC_EPS_LIFT) -> MeasurementBatch:
    io = _IO()
    b = create_empty_measurement_batch(n_feat, n_surfel)
    pos = io.dev_in(positions, shape=(-1, 3))
    n = pos.shape[0]
    nv = min(n, n_feat)
    if nv > 0:
        cb = b._c()
        # keep every staged tensor referenced until the call returns (ctypes only sees raw addresses)
        cov, dirs = io.dev_in(covariances, shape=(-1, 3, 3)), io.dev_in(directions, shape=(-1, 3))
        kap, wts, ts = io.dev_in(kappas, shape=(-1,)), io.dev_in(weights, shape=(-1,)), io.dev_in(timestamps, shape=(-1,))
        col = io.dev_in(colors, shape=(-1, 3)) if colors is not None else None
        io.ctx.check(io.ctx.lib.gcs_batch_from_camera_splats(
            io.ctx.handle, io.stream(), L.ptr(pos), L.ptr(cov), L.ptr(dirs), L.ptr(kap), L.ptr(wts), L.ptr(ts), L.ptr(col), n,
            float(eps_lift), C.byref(cb)))
        del cov, dirs, kap, wts, ts, col
    b.n_camera_valid = nv
    return b


# --------------------------------------------------------------------------------------------------
# a10 surfel extraction
# --------------------------------------------------------------------------------------------------
@dataclass
class SurfelExtractionConfig:
    n_surfel: int = constants.GC_N_SURFEL
    n_feat: int = constants.GC_N_FEAT
    voxel_size_m: float = 0.1
    hex3d_num_cells_1: int = 32
    hex3d_num_cells_2: int = 32
    hex3d_num_cells_z: int = 8
    hex3d_max_occupants: int = 32
    min_points_per_voxel: int = 3
    sensor_noise_var_per_axis: float = 1e-6
    wishart_nu: float = 5.0
    wishart_psi_scale: float = 0.1
    kappa_main_scale: float = 10.0
    kappa_min: float = 0.1
    kappa_max: float = 100.0
    eig_min: float = 1e-12
    eps_lift: float = constants.GC_EPS_LIFT

    def _c(self) -> CSurfelCfg:
        return CSurfelCfg(self.hex3d_num_cells_1, self.hex3d_num_cells_2, self.hex3d_num_cells_z, self.hex3d_max_occupants,
                          self.min_points_per_voxel, self.voxel_size_m, self.sensor_noise_var_per_axis, self.wishart_nu,
                          self.wishart_psi_scale, self.kappa_main_scale, self.kappa_min, self.kappa_max, self.eig_min,
                          self.eps_lift)


def extract_lidar_surfels(points, timestamps, weights, config: Optional[SurfelExtractionConfig] = None,
                          base_batch: Optional[MeasurementBatch] = None, chart_id: str = constants.GC_CHART_ID,
                          anchor_id: str = "surfel_extraction", return_bucket: bool = False
                          ) -> Tuple[MeasurementBatch, CertBundle, ExpectedEffect]:
    return _drive(_extract_lidar_surfels_gen(points, timestamps, weights, config, base_batch, chart_id, anchor_id, return_bucket))


def _extract_lidar_surfels_gen(points, timestamps, weights, config, base_batch, chart_id, anchor_id, return_bucket=False):
    if config is None:
        config = SurfelExtractionConfig()
    io = _IO()
    pts = io.dev_in(points, shape=(-1, 3))
    n = pts.shape[0]
    t = io.dev_in(timestamps, shape=(-1,))
    w = io.dev_in(weights, shape=(-1,))
    if t.shape[0] != n or w.shape[0] != n:
        raise ValueError("extract_lidar_surfels: points/timestamps/weights length mismatch")
    batch = base_batch.clone() if base_batch is not None else create_empty_measurement_batch(config.n_feat, config.n_surfel)
    if batch.n_surfel != config.n_surfel or batch.n_feat != config.n_feat:
        raise ValueError("extract_lidar_surfels: base_batch budget differs from config")
    n_cells = config.hex3d_num_cells_1 * config.hex3d_num_cells_2 * config.hex3d_num_cells_z
    nv_d = io.zeros(1, dtype=torch.int32)
    bucket = io.empty(n_cells, config.hex3d_max_occupants, dtype=torch.int32) if return_bucket else None
    count = io.empty(n_cells, dtype=torch.int32) if return_bucket else None
    cb, cc = batch._c(), config._c()
    io.ctx.check(io.ctx.lib.gcs_extract_lidar_surfels(io.ctx.handle, io.stream(), L.ptr(pts), L.ptr(t), L.ptr(w), n, C.byref(cc),
                                                      C.byref(cb), L.ptr(nv_d), L.ptr(bucket), L.ptr(count)))
    batch.n_lidar_valid = -1                      # not known on the host yet
    n_use = int((yield io, nv_d, batch)[0])
    batch.n_lidar_valid = n_use
    cert, effect = _surfel_cert(n_use, config, chart_id, anchor_id, io.compute())
    if return_bucket:
        batch._bucket, batch._bucket_count = bucket, count
    return batch, cert, effect


def _surfel_cert(n_use, config, chart_id, anchor_id, compute):
    support_frac = float(n_use) / float(max(config.n_surfel, 1))
    cert = CertBundle.create_approx(chart_id=chart_id, anchor_id=anchor_id,
                                    triggers=["ma_hex3d_binning", "plane_fit_batched", "wishart_regularization"],
                                    support=SupportCert(ess_total=float(n_use), support_frac=support_frac),
                                    influence=InfluenceCert.identity(), compute=compute)
    return cert, ExpectedEffect("surfel_extraction", float(n_use), float(n_use))


# --------------------------------------------------------------------------------------------------
# a11 atlas: device tile pool
# --------------------------------------------------------------------------------------------------
_ATLAS_SHAPES = dict(Lambdas=(3, 3), thetas=(3,), etas=(constants.GC_VMF_N_LOBES, 3), weights=(), timestamps=(),
                     created_timestamps=(), last_supported_scan_seq=(), last_update_scan_seq=(), primitive_ids=(), valid=(),
                     colors=(3,), cam_mass=(), lidar_mass=(), rgb_cam_accum=(3,), rgb_cam_denom=(), rgb=(3,))
_ATLAS_DTYPES = dict(last_supported_scan_seq=torch.int64, last_update_scan_seq=torch.int64, primitive_ids=torch.int64,
                     valid=torch.uint8)
_NP_NAME = dict(valid="valid_mask")


class AtlasMap:
    """
    Atlas of primitive map tiles, resident in HBM: one (T_cap, M_TILE, ...) array per PrimitiveMapTile field
    (15.6 MB per 50,000-slot tile) and a host dict tile_id -> pool row.  Mirrors AtlasMap / PrimitiveMapTile
    (primitive_map.py:98-211); tiles are created on demand as the reference's fuse / insert do.
    """

    def __init__(self, m_tile: int = constants.GC_PRIMITIVE_MAP_MAX_SIZE, n_tiles_cap: int = 32, device=None):
        io = _IO(device)
        self.device = io.dev
        self.m_tile, self.n_tiles_cap = int(m_tile), int(n_tiles_cap)
        self.fields: Dict[str, torch.Tensor] = {}
        for name in _ATLAS_FIELDS:
            self.fields[name] = torch.zeros((self.n_tiles_cap, self.m_tile) + _ATLAS_SHAPES[name],
                                            dtype=_ATLAS_DTYPES.get(name, F64), device=self.device)
        self.fields["rgb"].fill_(0.5)
        self.tiles: Dict[int, int] = {}       # tile_id -> pool row
        self.next_global_id = 0
        self.total_count = 0

    @property
    def n_tiles(self) -> int:
        return len(self.tiles)

    @property
    def tile_ids(self) -> List[int]:
        return list(self.tiles.keys())

    def _c(self) -> CAtlas:
        a = CAtlas()
        for name in _ATLAS_FIELDS:
            setattr(a, name, L.ptr(self.fields[name]))
        a.m_tile, a.n_tiles_cap = self.m_tile, self.n_tiles_cap
        return a

    def ensure_tile(self, tile_id: int) -> int:
        tile_id = int(tile_id)
        if tile_id not in self.tiles:
            if len(self.tiles) >= self.n_tiles_cap:
                raise RuntimeError(f"AtlasMap pool is full ({self.n_tiles_cap} tiles); create it with a larger n_tiles_cap")
            self.tiles[tile_id] = len(self.tiles)  # pool rows are zero-initialised = create_empty_tile
        return self.tiles[tile_id]

    def index_list(self, tile_ids, create: bool = False):
        return [(self.ensure_tile(t) if create else self.tiles.get(int(t), -1)) for t in tile_ids]

    def tile_count(self, tile_id: int) -> int:
        row = self.tiles.get(int(tile_id))
        return 0 if row is None else int(self.fields["valid"][row].sum().item())

    def upload_tile(self, tile_id: int, tile: dict):
        """Host NumPy tile (field names of PrimitiveMapTile) -> pool row."""
        row = self.ensure_tile(tile_id)
        for name in _ATLAS_FIELDS:
            src = np.asarray(tile[_NP_NAME.get(name, name)])
            t = torch.from_numpy(np.ascontiguousarray(src.astype(np.uint8) if name == "valid" else src))
            self.fields[name][row].copy_(t.to(self.fields[name].dtype))

    def download_tile(self, tile_id: int) -> dict:
        row = self.tiles[int(tile_id)]
        out = {"tile_id": int(tile_id)}
        for name in _ATLAS_FIELDS:
            a = self.fields[name][row].cpu().numpy()
            out[_NP_NAME.get(name, name)] = a.astype(bool) if name == "valid" else a
        out["count"] = int(out["valid_mask"].sum())
        return out

    @classmethod
    def from_numpy(cls, atlas: dict, n_tiles_cap: Optional[int] = None, device=None) -> "AtlasMap":
        cap = n_tiles_cap or max(8, len(atlas["tiles"]) + 8)
        a = cls(m_tile=atlas["m_tile"], n_tiles_cap=cap, device=device)
        for tid, t in atlas["tiles"].items():
            a.upload_tile(tid, t)
        a.next_global_id, a.total_count = int(atlas["next_global_id"]), int(atlas["total_count"])
        return a


def create_empty_atlas_map(m_tile: int = constants.GC_PRIMITIVE_MAP_MAX_SIZE, n_tiles_cap: int = 32) -> AtlasMap:
    return AtlasMap(m_tile=m_tile, n_tiles_cap=n_tiles_cap)


@dataclass
class PrimitiveMapRecencyInflateStats:
    staleness_inflation_strength: float
    staleness_cov_inflation_trace: float
    stale_precision_downscale_total: float


def _i32arr(v):
    return (_i32 * len(v))(*[int(x) for x in v])


def _i64arr(v):
    return (_i64 * len(v))(*[int(x) for x in v])


def primitive_map_recency_inflate(atlas_map: AtlasMap, tile_ids: List[int], scan_seq: int,
                                  recency_decay_lambda: float = constants.GC_RECENCY_DECAY_LAMBDA,
                                  min_scale: float = constants.GC_RECENCY_MIN_SCALE, chart_id: str = constants.GC_CHART_ID,
                                  anchor_id: str = "primitive_map_recency_inflate"):
    """In place on the device pool (the reference rebuilds the tiles functionally); returns (atlas, cert, effect, stats)."""
    return _drive(_recency_inflate_gen(atlas_map, tile_ids, scan_seq, recency_decay_lambda, min_scale, chart_id, anchor_id))


def _recency_inflate_gen(atlas_map, tile_ids, scan_seq, recency_decay_lambda=constants.GC_RECENCY_DECAY_LAMBDA,
                         min_scale=constants.GC_RECENCY_MIN_SCALE, chart_id=constants.GC_CHART_ID,
                         anchor_id="primitive_map_recency_inflate"):
    io = _IO(atlas_map.device)
    idx = atlas_map.index_list(tile_ids, create=False)
    stats_d = io.zeros(4)
    ca = atlas_map._c()
    io.ctx.check(io.ctx.lib.gcs_map_recency_inflate(io.ctx.handle, io.stream(), C.byref(ca), _i32arr(idx), len(idx),
                                                    int(scan_seq), float(recency_decay_lambda), float(min_scale), L.ptr(stats_d)))
    s = yield io, stats_d, atlas_map
    return (atlas_map,) + _inflate_finish(s, chart_id, anchor_id, io.compute())


def _inflate_finish(s, chart_id, anchor_id, compute):
    stats = PrimitiveMapRecencyInflateStats(staleness_inflation_strength=float(s[0] / max(s[2], 1.0)),
                                            staleness_cov_inflation_trace=float(s[1]),
                                            stale_precision_downscale_total=float(s[0]))
    cert = CertBundle.create_exact(chart_id=chart_id, anchor_id=anchor_id, compute=compute)
    return cert, ExpectedEffect("primitive_map_recency_inflate", float(s[2]), float(s[2])), stats


@dataclass
class AtlasMapView:
    candidate_tile_ids: torch.Tensor
    candidate_slots: torch.Tensor
    valid_mask: torch.Tensor
    tile_ids: List[int]
    m_tile_view: int
    positions: torch.Tensor
    covariances: torch.Tensor
    directions: torch.Tensor
    kappas: torch.Tensor
    weights: torch.Tensor
    primitive_ids: torch.Tensor
    last_supported_scan_seq: torch.Tensor
    etas: torch.Tensor
    colors: torch.Tensor
    n_valid: int = 0

    @property
    def count(self) -> int:
        return int(self.positions.shape[0])

    def _c(self) -> CMapView:
        v = CMapView()
        v.candidate_tile_ids, v.candidate_slots, v.valid = L.ptr(self.candidate_tile_ids), L.ptr(self.candidate_slots), L.ptr(self.valid_mask)
        v.positions, v.covariances, v.directions, v.kappas = L.ptr(self.positions), L.ptr(self.covariances), L.ptr(self.directions), L.ptr(self.kappas)
        v.weights, v.primitive_ids, v.last_supported_scan_seq = L.ptr(self.weights), L.ptr(self.primitive_ids), L.ptr(self.last_supported_scan_seq)
        v.etas, v.colors = L.ptr(self.etas), L.ptr(self.colors)
        return v


def _empty_view(io, tile_ids, m_tile_view) -> AtlasMapView:
    P = len(tile_ids) * int(m_tile_view)
    return AtlasMapView(candidate_tile_ids=io.empty(P, dtype=torch.int64), candidate_slots=io.empty(P, dtype=torch.int32),
                        valid_mask=io.empty(P, dtype=torch.uint8), tile_ids=[int(t) for t in tile_ids], m_tile_view=int(m_tile_view),
                        positions=io.empty(P, 3), covariances=io.empty(P, 3, 3), directions=io.empty(P, 3), kappas=io.empty(P),
                        weights=io.empty(P), primitive_ids=io.empty(P, dtype=torch.int64),
                        last_supported_scan_seq=io.empty(P, dtype=torch.int64), etas=io.empty(P, constants.GC_VMF_N_LOBES, 3),
                        colors=io.empty(P, 3))


def extract_atlas_map_view(atlas_map: AtlasMap, tile_ids: List[int], m_tile_view: int, eps_lift: float = constants.GC_EPS_LIFT,
                           eps_mass: float = constants.GC_EPS_MASS) -> AtlasMapView:
    return _drive(_extract_atlas_map_view_gen(atlas_map, tile_ids, m_tile_view, eps_lift, eps_mass))


def _extract_atlas_map_view_gen(atlas_map, tile_ids, m_tile_view, eps_lift=constants.GC_EPS_LIFT, eps_mass=constants.GC_EPS_MASS):
    if m_tile_view <= 0:
        raise ValueError(f"extract_atlas_map_view: m_tile_view must be > 0, got {m_tile_view}")
    io = _IO(atlas_map.device)
    view = _empty_view(io, tile_ids, m_tile_view)
    nv = io.zeros(1, dtype=torch.int32)
    ca, cv = atlas_map._c(), view._c()
    idx = atlas_map.index_list(tile_ids, create=False)
    io.ctx.check(io.ctx.lib.gcs_extract_atlas_map_view(io.ctx.handle, io.stream(), C.byref(ca), _i32arr(idx), _i64arr(tile_ids),
                                                       len(tile_ids), int(m_tile_view), float(eps_lift), float(eps_mass),
                                                       C.byref(cv), L.ptr(nv)))
    view.n_valid = -1                             # not known on the host yet
    view.n_valid = int((yield io, nv, view)[0])
    return view


# --------------------------------------------------------------------------------------------------
# a12 association
# --------------------------------------------------------------------------------------------------
class MeasurementMassPolicy(Enum):
    UNIFORM = "uniform"
    WEIGHT_PROPORTIONAL = "weight_proportional"
    FEATURE_CONFIDENCE = "feature_confidence"


class MapMassPolicy(Enum):
    UNIFORM = "uniform"
    PRIMITIVE_MASS = "primitive_mass"
    MASS_TEMPERED = "mass_tempered"


@dataclass
class AssociationConfig:
    k_assoc: int = constants.GC_K_ASSOC
    k_sinkhorn: int = constants.GC_K_SINKHORN
    beta: float = 0.5
    epsilon: float = 0.1
    tau_a: float = 0.5
    tau_b: float = 0.5
    cost_subtract_row_min: bool = True
    cost_scale_by_median: bool = False
    a_policy: MeasurementMassPolicy = MeasurementMassPolicy.UNIFORM
    b_policy: MapMassPolicy = MapMassPolicy.UNIFORM
    eps_mass: float = constants.GC_EPS_MASS
    h_tile: float = constants.GC_H_TILE
    r_stencil_tiles_xy: int = constants.GC_R_STENCIL_TILES_XY
    r_stencil_tiles_z: int = constants.GC_R_STENCIL_TILES_Z
    scan_seq: int = 0
    recency_decay_lambda: float = constants.GC_RECENCY_DECAY_LAMBDA


@dataclass
class PrimitiveAssociationResult:
    responsibilities: torch.Tensor
    candidate_pool_indices: torch.Tensor
    candidate_tile_ids: torch.Tensor
    candidate_slots: torch.Tensor
    row_masses: torch.Tensor
    cost_matrix: torch.Tensor

    def _c(self) -> CAssocResult:
        r = CAssocResult()
        r.responsibilities, r.candidate_pool_indices = L.ptr(self.responsibilities), L.ptr(self.candidate_pool_indices)
        r.candidate_tile_ids, r.candidate_slots = L.ptr(self.candidate_tile_ids), L.ptr(self.candidate_slots)
        r.row_masses, r.cost_matrix = L.ptr(self.row_masses), L.ptr(self.cost_matrix)
        return r


def _empty_assoc(io, N, K, zero=True):
    """Result arrays; zero=False: uninitialised, for the kernels that write every entry (six fill launches per scan less)."""
    z = io.zeros if zero else io.empty
    return PrimitiveAssociationResult(responsibilities=z(N, K), candidate_pool_indices=z(N, K, dtype=torch.int32),
                                      candidate_tile_ids=z(N, K, dtype=torch.int64),
                                      candidate_slots=z(N, K, dtype=torch.int64), row_masses=z(N),
                                      cost_matrix=z(N, K))


def associate_primitives_ot(measurement_batch: MeasurementBatch, map_view: AtlasMapView, config: AssociationConfig = None,
                            eps_lift: float = constants.GC_EPS_LIFT, eps_mass: float = constants.GC_EPS_MASS,
                            chart_id: str = constants.GC_CHART_ID, anchor_id: str = "primitive_ot"
                            ) -> Tuple[PrimitiveAssociationResult, CertBundle, ExpectedEffect]:
    return _drive(_associate_primitives_ot_gen(measurement_batch, map_view, config, eps_lift, eps_mass, chart_id, anchor_id))


def _associate_primitives_ot_gen(measurement_batch, map_view, config=None, eps_lift=constants.GC_EPS_LIFT,
                                 eps_mass=constants.GC_EPS_MASS, chart_id=constants.GC_CHART_ID, anchor_id="primitive_ot"):
    if config is None:
        config = AssociationConfig()
    _check_assoc_config(config)
    io = _IO(measurement_batch.Lambdas.device)
    N, K = measurement_batch.n_total, int(config.k_assoc)
    if measurement_batch.n_valid == 0 or map_view.n_valid == 0:
        cert = CertBundle.create_exact(chart_id=chart_id, anchor_id=anchor_id)
        return _empty_assoc(io, N, K), cert, ExpectedEffect("primitive_association_ot", 0.0, 0.0)
    res = _empty_assoc(io, N, K, zero=False)
    cfg = _c_assoc_cfg(config, eps_lift)
    cert_d = io.empty(OT["NCERT"])     # the kernel writes all NCERT entries
    cb, cv, cr = measurement_batch._c(), map_view._c(), res._c()
    io.ctx.check(io.ctx.lib.gcs_associate_primitives_ot(io.ctx.handle, io.stream(), C.byref(cb), C.byref(cv),
                                                        _i64arr(map_view.tile_ids), len(map_view.tile_ids),
                                                        int(map_view.m_tile_view), C.byref(cfg), C.byref(cr), L.ptr(cert_d)))
    c = yield io, cert_d, res
    return (res,) + _assoc_cert(c, N, K, config, chart_id, anchor_id, io)


def _check_assoc_config(config):
    if config.a_policy not in (MeasurementMassPolicy.UNIFORM, MeasurementMassPolicy.WEIGHT_PROPORTIONAL):
        raise ValueError(f"Unsupported measurement mass policy: {config.a_policy}. Only UNIFORM and WEIGHT_PROPORTIONAL are implemented.")
    if config.b_policy != MapMassPolicy.UNIFORM:
        raise ValueError(f"Unsupported map mass policy: {config.b_policy}. Only UNIFORM is implemented.")
    if not config.cost_subtract_row_min or config.cost_scale_by_median:
        raise ValueError("associate_primitives_ot: only cost_subtract_row_min=True, cost_scale_by_median=False is built")


def _c_assoc_cfg(config, eps_lift) -> CAssocCfg:
    return CAssocCfg(int(config.k_assoc), int(config.k_sinkhorn), int(config.r_stencil_tiles_xy), int(config.r_stencil_tiles_z),
                     1 if config.a_policy == MeasurementMassPolicy.WEIGHT_PROPORTIONAL else 0, 0, float(config.beta), float(config.epsilon), float(config.tau_a), float(config.tau_b), float(config.eps_mass),
                     float(eps_lift), float(config.h_tile), float(config.recency_decay_lambda), int(config.scan_seq))


def _assoc_cert(c, N, K, config, chart_id, anchor_id, io):
    """OTCert bundle + effect from the association kernel's certificate sums (primitive_association.py:480-553)."""
    tm = float(c[OT["MASS_TOTAL"]])
    n_nonzero_a = int(c[OT["NONZERO_A"]])
    compute = io.compute(alloc_bytes_est=int(N * K * 8 * 4), largest_tensor_shape=(int(N), int(K)), segment_sum_k=int(K),
                         psd_projection_count=0, chol_solve_count=0)
    cert = CertBundle.create_approx(
        chart_id=chart_id, anchor_id=anchor_id, triggers=["sinkhorn_fixed_iter", "sinkhorn_unbalanced_kl_relax"],
        frobenius_applied=False,
        support=SupportCert(ess_total=float(c[OT["ESS"]]), support_frac=float(n_nonzero_a) / float(max(N, 1))),
        influence=InfluenceCert.identity().with_overrides(mass_epsilon_ratio=float(config.eps_mass) / (tm + config.eps_mass)),
        compute=compute)
    b_val = 1.0 / float(K)
    cert.ot = OTCert(marginal_defect_a=float(c[OT["MARGINAL_A"]]), marginal_defect_b=float(c[OT["MARGINAL_B"]]),
                     transport_mass_total=tm, dual_gap_proxy=0.0, sum_a=float(c[OT["SUM_A"]]), sum_b=float(b_val * K),
                     sum_m=float(c[OT["SUM_M"]]), sum_novel=float(c[OT["SUM_NOVEL"]]), p95_a=float(c[OT["P95_A"]]), p95_b=b_val,
                     nonzero_a=n_nonzero_a, nonzero_b=int(K if b_val > config.eps_mass else 0), epsilon=float(config.epsilon),
                     tau_a=float(config.tau_a), tau_b=float(config.tau_b), n_iters=int(config.k_sinkhorn),
                     b_policy=str(config.b_policy.value), b_recency_decay_lambda=float(config.recency_decay_lambda),
                     b_recency_p95=float(c[OT["B_RECENCY_P95"]]))
    total_cost = float(c[OT["TOTAL_COST"]])
    return cert, ExpectedEffect("primitive_association_ot", total_cost, total_cost)


# --------------------------------------------------------------------------------------------------
# a13 pose evidence
# --------------------------------------------------------------------------------------------------
@dataclass
class VisualPoseEvidenceResult:
    L_pose: torch.Tensor
    h_pose: torch.Tensor
    L_trans: torch.Tensor
    h_trans: torch.Tensor
    L_rot: torch.Tensor
    h_rot: torch.Tensor
    total_weighted_cost: float
    n_associations: int
    mean_transported_mass: float


def visual_pose_evidence(association_result: PrimitiveAssociationResult, measurement_batch: MeasurementBatch,
                         map_view: AtlasMapView, belief_pred, eps_lift: float = constants.GC_EPS_LIFT,
                         eps_mass: float = constants.GC_EPS_MASS, chart_id: str = constants.GC_CHART_ID,
                         anchor_id: str = "visual_pose_evidence", z_lin_pose=None
                         ) -> Tuple[VisualPoseEvidenceResult, CertBundle, ExpectedEffect]:
    return _drive(_visual_pose_evidence_gen(association_result, measurement_batch, map_view, belief_pred, eps_lift, eps_mass,
                                            chart_id, anchor_id, z_lin_pose))


def _visual_pose_evidence_gen(association_result, measurement_batch, map_view, belief_pred, eps_lift=constants.GC_EPS_LIFT,
                              eps_mass=constants.GC_EPS_MASS, chart_id=constants.GC_CHART_ID, anchor_id="visual_pose_evidence",
                              z_lin_pose=None):
    io = _IO(measurement_batch.Lambdas.device)
    N_meas = measurement_batch.n_valid
    N_assoc, K = association_result.responsibilities.shape
    if N_meas == 0 or N_assoc == 0 or map_view.n_valid == 0:
        return _empty_pose_evidence(io, eps_lift, chart_id, anchor_id)
    if z_lin_pose is not None:
        pose = _host_vec(z_lin_pose.detach().cpu().numpy().ravel()[:6] if isinstance(z_lin_pose, torch.Tensor)
                         else np.asarray(z_lin_pose, np.float64).ravel()[:6], 6)
    elif hasattr(belief_pred, "mean_world_pose"):
        pose = _host_vec(belief_pred.mean_world_pose(eps_lift=eps_lift), 6)
    else:
        pose = _host_vec(belief_pred, 6)
    L22, h22, rec_d = io.empty(22, 22), io.empty(22), io.empty(VP["NREC"])   # the kernel writes all NREC entries
    cb, cv, cr = measurement_batch._c(), map_view._c(), association_result._c()
    io.ctx.check(io.ctx.lib.gcs_visual_pose_evidence(io.ctx.handle, io.stream(), C.byref(cb), C.byref(cv), C.byref(cr), int(K),
                                                     _dptr(pose), float(eps_lift), float(eps_mass), L.ptr(L22), L.ptr(h22), L.ptr(rec_d)))
    r = yield io, rec_d, None
    return _pose_evidence_finish(r, L22, h22, rec_d, N_meas, K, eps_lift, chart_id, anchor_id, io.compute())


def _empty_pose_evidence(io, eps_lift, chart_id, anchor_id):
    res = VisualPoseEvidenceResult(L_pose=eps_lift * torch.eye(22, dtype=F64, device=io.dev), h_pose=io.zeros(22),
                                   L_trans=io.zeros(3, 3), h_trans=io.zeros(3), L_rot=io.zeros(3, 3), h_rot=io.zeros(3),
                                   total_weighted_cost=0.0, n_associations=0, mean_transported_mass=0.0)
    return res, CertBundle.create_exact(chart_id=chart_id, anchor_id=anchor_id), ExpectedEffect("visual_pose_evidence", 0.0, 0.0)


def _pose_evidence_finish(r, L22, h22, rec_d, N_meas, K, eps_lift, chart_id, anchor_id, compute):
    n_rows = int(r[VP["N_VALID_ROWS"]])
    total_cost = float(r[VP["TRANS_COST"]] + r[VP["ROT_COST"]])
    res = VisualPoseEvidenceResult(L_pose=L22, h_pose=h22, L_trans=rec_d[0:9].reshape(3, 3), h_trans=rec_d[9:12],
                                   L_rot=rec_d[12:21].reshape(3, 3), h_rot=rec_d[21:24], total_weighted_cost=total_cost,
                                   n_associations=int(n_rows * K),
                                   mean_transported_mass=float(r[VP["SUM_ROW_MASS"]] / max(n_rows, 1)))
    cert = CertBundle.create_approx(
        chart_id=chart_id, anchor_id=anchor_id, triggers=["linearization", "ot_soft_correspondence"], frobenius_applied=True,
        support=SupportCert(ess_total=float(r[VP["SUM_ROW_MASS"]]), support_frac=float(n_rows) / float(max(N_meas, 1))),
        influence=InfluenceCert.identity().with_overrides(lift_strength=eps_lift), compute=compute)
    return res, cert, ExpectedEffect("visual_pose_evidence", total_cost, total_cost)


# --------------------------------------------------------------------------------------------------
# a14 map update (pipeline step 12b)
# --------------------------------------------------------------------------------------------------
@dataclass
class MapUpdateResult:
    atlas_map: AtlasMap
    n_fused: int
    n_inserted: int
    n_culled: int
    new_ids: torch.Tensor       # (n_tiles, k_insert) int64, -1 where nothing was inserted
    insert_slots: torch.Tensor  # (n_tiles, k_insert) int32
    tile_counts: List[int]


def map_update_step12b(atlas_map: AtlasMap, measurement_batch: MeasurementBatch, association_result: PrimitiveAssociationResult,
                       active_tile_ids: List[int], z_t, scan_seq: int, timestamp: float,
                       k_insert_tile: int = constants.GC_K_INSERT_TILE,
                       recency_decay_lambda: float = constants.GC_RECENCY_DECAY_LAMBDA, eps_lift: float = constants.GC_EPS_LIFT,
                       eps_mass: float = constants.GC_EPS_MASS, h_tile: float = constants.GC_H_TILE,
                       cull_weight_threshold: float = constants.GC_PRIMITIVE_CULL_WEIGHT_THRESHOLD,
                       forgetting_factor: float = constants.GC_PRIMITIVE_FORGETTING_FACTOR,
                       assoc_block_size: int = constants.GC_ASSOC_BLOCK_SIZE, strict_tile_state: bool = True,
                       inflate_stats: Optional[PrimitiveMapRecencyInflateStats] = None,
                       chart_id: str = constants.GC_CHART_ID) -> Tuple[MapUpdateResult, CertBundle, ExpectedEffect]:
    return _drive(_map_update_step12b_gen(atlas_map, measurement_batch, association_result, active_tile_ids, z_t, scan_seq,
                                          timestamp, k_insert_tile, recency_decay_lambda, eps_lift, eps_mass, h_tile,
                                          cull_weight_threshold, forgetting_factor, assoc_block_size, strict_tile_state,
                                          inflate_stats, chart_id))


def _map_update_step12b_gen(atlas_map, measurement_batch, association_result, active_tile_ids, z_t, scan_seq, timestamp,
                            k_insert_tile=constants.GC_K_INSERT_TILE, recency_decay_lambda=constants.GC_RECENCY_DECAY_LAMBDA,
                            eps_lift=constants.GC_EPS_LIFT, eps_mass=constants.GC_EPS_MASS, h_tile=constants.GC_H_TILE,
                            cull_weight_threshold=constants.GC_PRIMITIVE_CULL_WEIGHT_THRESHOLD,
                            forgetting_factor=constants.GC_PRIMITIVE_FORGETTING_FACTOR,
                            assoc_block_size=constants.GC_ASSOC_BLOCK_SIZE, strict_tile_state=True, inflate_stats=None,
                            chart_id=constants.GC_CHART_ID, next_id_dev=None):
    """
    Whole primitive-map update of one scan in one call (in place on the device pool): rigid pushforward of the
    measurement batch with z_t, PoE fuse into the associated slots, novelty-driven insertion into the lowest-retention
    slots, cull, forget.  Replaces the Python loops of pipeline.py:1258-1447 (6 blocks x 7 tiles of primitive_map_fuse,
    7 x insert_masked, 7 x cull / forget).
    """
    io = _IO(atlas_map.device)
    n_exist_before = len([t for t in active_tile_ids if int(t) in atlas_map.tiles])
    idx = atlas_map.index_list(active_tile_ids, create=True)
    nt = len(active_tile_ids)
    K = association_result.responsibilities.shape[1]
    cfg = CMapUpdateCfg(int(k_insert_tile), int(K), int(assoc_block_size), 1 if strict_tile_state else 0,
                        float(recency_decay_lambda), float(eps_lift), float(eps_mass), float(h_tile),
                        float(cull_weight_threshold), float(forgetting_factor), int(scan_seq), int(atlas_map.next_global_id),
                        float(timestamp), L.ptr(next_id_dev) if next_id_dev is not None else None)
    # (next_id_dev: a device int64[1] counter the kernels read and advance -- the caller may then enqueue the next scan's
    # update before this one's statistics have been read; the host attribute catches up when they are)
    active_set = set(int(x) for x in active_tile_ids)
    inactive = [int(t) for t in atlas_map.tile_ids if int(t) not in active_set]   # as of this scan, not of the read-back
    new_ids = io.empty(nt, int(k_insert_tile), dtype=torch.int64)
    slots = io.empty(nt, int(k_insert_tile), dtype=torch.int32)
    stats_d = io.zeros(MU["NSTATS"])
    ca, cb, cr = atlas_map._c(), measurement_batch._c(), association_result._c()
    pose = _host_vec(z_t, 6)
    io.ctx.check(io.ctx.lib.gcs_map_update(io.ctx.handle, io.stream(), C.byref(ca), _i32arr(idx), _i64arr(active_tile_ids), nt,
                                           C.byref(cb), C.byref(cr), _dptr(pose), C.byref(cfg), L.ptr(new_ids), L.ptr(slots),
                                           L.ptr(stats_d)))
    s = yield io, stats_d, atlas_map
    if callable(inflate_stats):          # resolved late: the caller learns the inflation statistics in the same read-back
        inflate_stats = inflate_stats()
    n_ins, n_cull = int(s[MU["INSERT_COUNT"]]), int(s[MU["EVICTED_COUNT"]])
    # (with the device-resident counter several updates may be read back late: ids only grow)
    atlas_map.next_global_id = max(int(atlas_map.next_global_id), int(s[MU["NEXT_GLOBAL_ID"]])) if next_id_dev is not None \
        else int(s[MU["NEXT_GLOBAL_ID"]])
    atlas_map.total_count = atlas_map.total_count + n_ins - n_cull
    result = MapUpdateResult(atlas_map=atlas_map, n_fused=int(s[MU["FUSED_COUNT"]]), n_inserted=n_ins, n_culled=n_cull,
                             new_ids=new_ids, insert_slots=slots,
                             tile_counts=[int(s[MU["TILE_COUNT0"] + a]) for a in range(min(nt, 16))])
    mu = MapUpdateCert(
        n_active_tiles=nt, tile_ids_active=[int(t) for t in active_tile_ids], n_inactive_tiles=len(inactive),
        tile_ids_inactive=inactive, tile_cache_hits=nt, tile_cache_misses=0,
        insert_count_total=n_ins, insert_mass_total=float(s[MU["INSERT_MASS"]]), insert_mass_p95=float(s[MU["INSERT_MASS_P95"]]),
        evicted_count=n_cull, evicted_mass_total=float(s[MU["EVICTED_MASS"]]), fused_count=int(s[MU["FUSED_COUNT"]]),
        fused_mass_total=float(s[MU["FUSED_MASS"]]), merged_count=0,
        staleness_inflation_strength=float(inflate_stats.staleness_inflation_strength) if inflate_stats else 0.0,
        staleness_cov_inflation_trace=float(inflate_stats.staleness_cov_inflation_trace) if inflate_stats else 0.0,
        stale_precision_downscale_total=float(inflate_stats.stale_precision_downscale_total) if inflate_stats else 0.0)
    _ = n_exist_before
    cert = CertBundle.create_exact(chart_id=chart_id, anchor_id="map_update", map_update=mu, compute=io.compute())
    return result, cert, ExpectedEffect("map_update", float(n_ins), float(n_ins))


# --------------------------------------------------------------------------------------------------
# fused fast entry: the primitive-family LiDAR evidence path of one scan with TWO host synchronisations
# --------------------------------------------------------------------------------------------------
def lidar_evidence_primitives(points, timestamps, weights, scan_start_time: float, scan_end_time: float, xi_body,
                              atlas_map: AtlasMap, active_tile_ids: List[int], pose_pred, scan_seq: int,
                              base_batch: Optional[MeasurementBatch] = None, z_t=None,
                              surfel_config: Optional[SurfelExtractionConfig] = None,
                              association_config: Optional[AssociationConfig] = None,
                              m_tile_view: int = constants.GC_M_TILE_VIEW, ess_imu: float = 1.0, update_map: bool = True,
                              map_update_kwargs: Optional[dict] = None, chart_id: str = constants.GC_CHART_ID,
                              anchor_id: str = "lidar_evidence_primitives") -> dict:
    """
    The call sequence of process_scan_single_hypothesis for the primitive family (fl/backend/pipeline.py:569-587 deskew,
    :780-800 surfels, :835-853 recency inflate + map view, :855-877 association, :998-1010 pose evidence, :1233-1447 map
    update) in one call.  The same C entry points run in the same order on the current stream and every stage returns
    exactly the tuple its stand-alone operator returns -- results are bit-identical to calling the operators one by
    one -- but the seven per-operator certificate read-backs collapse into two: one after the map view (the only point
    where the host must decide something: empty measurement batch / empty view take the reference's early exits) and
    one at the end.  Keys: deskew, surfels, recency_inflate, map_view, association, pose_evidence, map_update.
    """
    if z_t is None:
        z_t = pose_pred
    if association_config is None:
        association_config = AssociationConfig(scan_seq=int(scan_seq))
    g_dk = _Pending(_deskew_constant_twist_gen(points, timestamps, weights, scan_start_time, scan_end_time, xi_body, ess_imu,
                                               chart_id, anchor_id))
    dk = g_dk.provisional
    g_sf = _Pending(_extract_lidar_surfels_gen(dk.points, dk.timestamps, dk.weights, surfel_config, base_batch, chart_id,
                                               "surfel_extraction"))
    g_ri = _Pending(_recency_inflate_gen(atlas_map, active_tile_ids, scan_seq, association_config.recency_decay_lambda))
    g_mv = _Pending(_extract_atlas_map_view_gen(atlas_map, active_tile_ids, m_tile_view))
    dk_out, sf_out, ri_out, view = drive_group([g_dk, g_sf, g_ri, g_mv])
    batch = sf_out[0]
    inflate_stats = ri_out[3]
    g_as = _Pending(_associate_primitives_ot_gen(batch, view, association_config))
    assoc = g_as.provisional
    g_pe = _Pending(_visual_pose_evidence_gen(assoc, batch, view, pose_pred, z_lin_pose=pose_pred))
    group = [g_as, g_pe]
    if update_map:
        group.append(_Pending(_map_update_step12b_gen(atlas_map, batch, assoc, active_tile_ids, z_t, scan_seq, scan_end_time,
                                                      inflate_stats=inflate_stats, **(map_update_kwargs or {}))))
    outs = drive_group(group)
    return dict(deskew=dk_out, surfels=sf_out, recency_inflate=ri_out, map_view=view, association=outs[0],
                pose_evidence=outs[1], map_update=outs[2] if update_map else None)


# --------------------------------------------------------------------------------------------------
# map export (export half): valid primitives of the selected tiles -> renderable batch + /gc/map/points
# --------------------------------------------------------------------------------------------------
@dataclass
class RenderablePrimitiveBatch:
    """Canonical renderable batch (world frame), field names of primitive_map.py:452-470; device tensors, newest first."""
    mu_world: torch.Tensor
    Sigma_world: torch.Tensor
    Lambda_world: torch.Tensor
    eta: torch.Tensor
    mass: torch.Tensor
    color: torch.Tensor
    primitive_ids: torch.Tensor
    last_supported_scan_seq: torch.Tensor
    cloud: torch.Tensor          # uint8 (N * 16): x, y, z, intensity float32 LE records of /gc/map/points
    point_step: int = 16

    @property
    def count(self) -> int:
        return int(self.mass.shape[0])


def export_map_points(atlas_map: AtlasMap, tile_ids: Optional[List[int]] = None, max_primitives: Optional[int] = None,
                      eps_lift: float = constants.GC_EPS_LIFT) -> RenderablePrimitiveBatch:
    """
    PrimitiveMapPublisher.publish without the ROS objects (fl/backend/map_publisher.py:131-258): every valid primitive of
    the selected tiles (default: all, sorted tile ids) -> mu, Sigma, Lambda_world, eta, mass, colour in newest-first
    order (ties by primitive id) and the PointCloud2 payload of /gc/map/points, all on the device; one host read (the
    count).  `max_primitives` (per-tile down-selection) is only built for the publisher's default, None.
    """
    if max_primitives is not None:
        raise ValueError("export_map_points: max_primitives is not built (the reference's publisher default is None)")
    io = _IO(atlas_map.device)
    if tile_ids is None:
        tile_ids = sorted(int(t) for t in atlas_map.tile_ids)
    tile_ids = [int(t) for t in tile_ids if int(t) in atlas_map.tiles]
    if not tile_ids:
        z = lambda *s, dt=F64: io.zeros(*s, dtype=dt)
        return RenderablePrimitiveBatch(z(0, 3), z(0, 3, 3), z(0, 3, 3), z(0, constants.GC_VMF_N_LOBES, 3), z(0), z(0, 3),
                                        z(0, dt=torch.int64), z(0, dt=torch.int64), z(0, dt=torch.uint8))
    if len(tile_ids) > 256:
        raise ValueError(f"export_map_points: {len(tile_ids)} tiles in one call (limit 256): export in groups")
    idx = atlas_map.index_list(tile_ids, create=False)
    cap = len(tile_ids) * atlas_map.m_tile
    mu, Sig, Lam, eta = io.empty(cap, 3), io.empty(cap, 3, 3), io.empty(cap, 3, 3), io.empty(cap, constants.GC_VMF_N_LOBES, 3)
    mass, col = io.empty(cap), io.empty(cap, 3)
    pid, rec = io.empty(cap, dtype=torch.int64), io.empty(cap, dtype=torch.int64)
    cloud = io.empty(cap * 16, dtype=torch.uint8)
    n_d = io.zeros(1, dtype=torch.int32)
    ex = CMapExport(L.ptr(mu), L.ptr(Sig), L.ptr(Lam), L.ptr(eta), L.ptr(mass), L.ptr(col), L.ptr(pid), L.ptr(rec), L.ptr(cloud))
    ca = atlas_map._c()
    io.ctx.check(io.ctx.lib.gcs_export_map_points(io.ctx.handle, io.stream(), C.byref(ca), _i32arr(idx), len(idx), float(eps_lift),
                                                  C.byref(ex), int(cap), L.ptr(n_d)))
    n = int(io.host(n_d)[0])
    return RenderablePrimitiveBatch(mu[:n], Sig[:n], Lam[:n], eta[:n], mass[:n], col[:n], pid[:n], rec[:n], cloud[:n * 16])


# --------------------------------------------------------------------------------------------------
# merge-reduce (merge half)
# --------------------------------------------------------------------------------------------------
@dataclass
class PrimitiveMapMergeReduceResult:
    atlas_map: AtlasMap
    tile_id: int
    n_merged: int
    frobenius_correction: float


def primitive_map_merge_reduce(atlas_map: AtlasMap, tile_id: int,
                               merge_threshold: float = constants.GC_PRIMITIVE_MERGE_THRESHOLD,
                               max_pairs: int = constants.GC_K_MERGE_PAIRS_PER_TILE,
                               max_tile_size: int = constants.GC_PRIMITIVE_MERGE_MAX_TILE_SIZE,
                               eps_psd: float = constants.GC_EPS_PSD, eps_lift: float = constants.GC_EPS_LIFT,
                               chart_id: str = constants.GC_CHART_ID, anchor_id: str = "primitive_map"
                               ) -> Tuple[PrimitiveMapMergeReduceResult, CertBundle, ExpectedEffect]:
    """
    primitive_map_merge_reduce (fl/backend/structures/primitive_map.py:1809-2031), in place on the device pool: same
    no-op ladder (missing tile / fewer than two primitives / no pair budget -> exact no-op; tile larger than
    max_tile_size -> approximate no-op with the budget-cap trigger), same triggers and certificate fields otherwise.
    """
    def no_op(predicted, triggers=None, influence=None):
        res = PrimitiveMapMergeReduceResult(atlas_map=atlas_map, tile_id=int(tile_id), n_merged=0, frobenius_correction=0.0)
        if triggers:
            cert = CertBundle.create_approx(chart_id=chart_id, anchor_id=anchor_id, triggers=triggers, frobenius_applied=True,
                                            influence=influence or InfluenceCert.identity())
        else:
            cert = CertBundle.create_exact(chart_id=chart_id, anchor_id=anchor_id)
        return res, cert, ExpectedEffect("primitive_map_merge_reduce", float(predicted), 0.0)

    tile_id = int(tile_id)
    if tile_id not in atlas_map.tiles:
        return no_op(0.0)
    M = int(atlas_map.m_tile)
    if M < 2 or int(max_pairs) <= 0:
        return no_op(float(max_pairs))
    if int(max_tile_size) > 0 and M > int(max_tile_size):
        over = float(M - int(max_tile_size)) / float(max(M, 1))
        return no_op(float(max_pairs), ["merge_reduce_budget_cap"], InfluenceCert.identity().with_overrides(mass_epsilon_ratio=over))
    if M > 2048 or int(max_pairs) > 64:
        raise ValueError(f"primitive_map_merge_reduce: m_tile={M}, max_pairs={max_pairs} exceed the built budgets (2048 slots, 64 pairs)")
    io = _IO(atlas_map.device)
    stats_d = io.zeros(4)
    ca = atlas_map._c()
    row = atlas_map.index_list([tile_id], create=False)[0]
    io.ctx.check(io.ctx.lib.gcs_map_merge_reduce(io.ctx.handle, io.stream(), C.byref(ca), int(row), float(merge_threshold),
                                                 int(max_pairs), float(eps_psd), float(eps_lift), L.ptr(stats_d)))
    n_merged = int(io.host(stats_d)[0])
    if n_merged <= 0:
        return no_op(float(max_pairs))
    atlas_map.total_count = atlas_map.total_count - n_merged
    cert = CertBundle.create_approx(chart_id=chart_id, anchor_id=anchor_id, triggers=["primitive_map_merge_reduce"],
                                    frobenius_applied=True,
                                    influence=InfluenceCert.identity().with_overrides(mass_epsilon_ratio=float(n_merged) / float(max(M, 1))),
                                    compute=io.compute())
    res = PrimitiveMapMergeReduceResult(atlas_map=atlas_map, tile_id=tile_id, n_merged=n_merged, frobenius_correction=float(n_merged))
    return res, cert, ExpectedEffect("primitive_map_merge_reduce", float(max_pairs), float(n_merged))


# --------------------------------------------------------------------------------------------------
# the per-tile map operators, one call = one tile.  map_update_step12b runs the same arithmetic for all
# active tiles of a scan in one pass; these keep the reference's own granularity and signatures.
# --------------------------------------------------------------------------------------------------
def _optr(t):
    return L.ptr(t) if t is not None else None


def block_associations_for_fuse(result: PrimitiveAssociationResult, valid_mask, block_size: int = constants.GC_ASSOC_BLOCK_SIZE):
    """
    block_associations_for_fuse (fl/backend/operators/primitive_association.py:561-588): the (N, K) association cut into
    blocks of block_size rows (row indices clipped to N-1, rows past N or with valid_mask clear get responsibility 0).
    Pure re-indexing of device arrays.  Returns (meas_idx, candidate_tile_ids, candidate_slots, responsibilities,
    valid_rows) with leading shape (n_blocks, block_size).
    """
    resp = result.responsibilities
    n_total, _k = resp.shape
    block = int(max(1, block_size))
    n_blocks = (n_total + block - 1) // block
    meas_idx = torch.arange(n_blocks * block, dtype=torch.int32, device=resp.device).reshape(n_blocks, block)
    clipped = torch.clamp(meas_idx, max=n_total - 1)
    vm = _IO(resp.device).dev_in(valid_mask, torch.uint8).reshape(-1).to(torch.bool)
    li = clipped.long()
    valid_rows = (meas_idx < n_total) & vm[li]
    return (clipped, result.candidate_tile_ids[li], result.candidate_slots[li],
            resp[li] * valid_rows[:, :, None].to(resp.dtype), valid_rows)


@dataclass
class PrimitiveMapFuseResult:
    atlas_map: AtlasMap
    tile_id: int
    n_fused: int


def primitive_map_fuse(atlas_map: AtlasMap, tile_id: int, target_slots, Lambdas_meas, thetas_meas, etas_meas, weights_meas,
                       responsibilities, timestamp: float, scan_seq: int = 0, valid_mask=None, colors_meas=None,
                       sources_meas=None, eps_psd: float = constants.GC_EPS_PSD, eps_mass: float = constants.GC_EPS_MASS,
                       fuse_chunk_size: int = constants.GC_FUSE_CHUNK_SIZE, chart_id: str = constants.GC_CHART_ID,
                       anchor_id: str = "primitive_map") -> Tuple[PrimitiveMapFuseResult, CertBundle, ExpectedEffect]:
    """
    primitive_map_fuse (fl/backend/structures/primitive_map.py:992-1163), in place on the device pool.  The tile is
    created if missing; an empty proposal list is the reference's exact no-op (the new tile is then not kept, :1031-1041).
    fuse_chunk_size only sets the reference's chunking of one scatter-add and does not change its result: ignored.
    """
    del eps_psd, fuse_chunk_size
    tile_id = int(tile_id)
    io = _IO(atlas_map.device)
    slots = io.dev_in(target_slots, torch.int32).reshape(-1)
    n = int(slots.shape[0])
    if n == 0:
        return (PrimitiveMapFuseResult(atlas_map=atlas_map, tile_id=tile_id, n_fused=0),
                CertBundle.create_exact(chart_id=chart_id, anchor_id=anchor_id), ExpectedEffect("primitive_map_fuse", 0.0, 0.0))
    if n > 16384:      # checked before the tile is created: a rejected call leaves the map as it was
        raise ValueError(f"primitive_map_fuse: {n} proposals in one call exceed the built budget of 16384 (the pipeline "
                         f"fuses assoc_block_size x K_ASSOC = 2048 per call)")
    Lm = io.dev_in(Lambdas_meas, F64, (n, 3, 3))
    th = io.dev_in(thetas_meas, F64, (n, 3))
    et = io.dev_in(etas_meas, F64, (n, constants.GC_VMF_N_LOBES, 3))
    wm = io.dev_in(weights_meas, F64, (n,))
    rs = io.dev_in(responsibilities, F64, (n,))
    vm = io.dev_in(valid_mask, torch.uint8, (n,)) if valid_mask is not None else None
    # the reference ignores colour / source arrays shorter than the proposal list (:1081-1095)
    cm = io.dev_in(colors_meas, F64) if colors_meas is not None and len(colors_meas) >= n else None
    sm = io.dev_in(sources_meas, torch.int32).reshape(-1) if sources_meas is not None and len(sources_meas) >= n else None
    row = atlas_map.ensure_tile(tile_id)
    stats_d = io.zeros(4)
    ca = atlas_map._c()
    io.ctx.check(io.ctx.lib.gcs_map_fuse(io.ctx.handle, io.stream(), C.byref(ca), int(row), L.ptr(slots), L.ptr(Lm), L.ptr(th),
                                         L.ptr(et), L.ptr(wm), L.ptr(rs), _optr(vm), _optr(cm), _optr(sm), n, float(timestamp),
                                         int(scan_seq), float(eps_mass), L.ptr(stats_d)))
    n_unique = int(io.host(stats_d)[0])
    res = PrimitiveMapFuseResult(atlas_map=atlas_map, tile_id=tile_id, n_fused=n_unique)
    cert = CertBundle.create_exact(chart_id=chart_id, anchor_id=anchor_id, compute=io.compute())
    return res, cert, ExpectedEffect("primitive_map_fuse", float(n), float(n_unique))


@dataclass
class PrimitiveMapInsertResult:
    atlas_map: AtlasMap
    tile_id: int
    n_inserted: int
    new_ids: torch.Tensor
    target_slots: Optional[torch.Tensor] = None   # not in the reference's result: the evicted / filled slots, (K,) int32


def primitive_map_insert_masked(atlas_map: AtlasMap, tile_id: int, Lambdas_new, thetas_new, etas_new, weights_new,
                                timestamp: float, valid_new_mask, scan_seq: int = 0,
                                recency_decay_lambda: float = constants.GC_RECENCY_DECAY_LAMBDA, colors_new=None,
                                sources_new=None, chart_id: str = constants.GC_CHART_ID,
                                anchor_id: str = "primitive_map_insert_masked"
                                ) -> Tuple[PrimitiveMapInsertResult, CertBundle, ExpectedEffect]:
    """primitive_map_insert_masked (fl/backend/structures/primitive_map.py:807-981), in place on the device pool."""
    tile_id = int(tile_id)
    io = _IO(atlas_map.device)
    Ln = io.dev_in(Lambdas_new, F64)
    k = int(Ln.shape[0])
    if k < 1 or k > min(1024, atlas_map.m_tile):
        raise ValueError(f"primitive_map_insert_masked: K={k} proposals outside [1, min(1024, m_tile={atlas_map.m_tile})]")
    Ln = Ln.reshape(k, 3, 3)
    th = io.dev_in(thetas_new, F64, (k, 3))
    et = io.dev_in(etas_new, F64, (k, constants.GC_VMF_N_LOBES, 3))
    wn = io.dev_in(weights_new, F64, (k,))
    vm = io.dev_in(valid_new_mask, torch.uint8, (k,))
    cn = io.dev_in(colors_new, F64, (k, 3)) if colors_new is not None else None
    sn = io.dev_in(sources_new, torch.int32, (k,)) if sources_new is not None else None
    row = atlas_map.ensure_tile(tile_id)
    ids_d = io.empty(k, dtype=torch.int64)
    slots_d = io.empty(k, dtype=torch.int32)
    stats_d = io.zeros(4)
    ca = atlas_map._c()
    io.ctx.check(io.ctx.lib.gcs_map_insert_masked(io.ctx.handle, io.stream(), C.byref(ca), int(row), L.ptr(Ln), L.ptr(th), L.ptr(et),
                                                  L.ptr(wn), L.ptr(vm), _optr(cn), _optr(sn), k, float(timestamp), int(scan_seq),
                                                  float(recency_decay_lambda), int(atlas_map.next_global_id), L.ptr(ids_d),
                                                  L.ptr(slots_d), L.ptr(stats_d)))
    st = io.host(stats_d)
    n_inserted, dropped = int(st[0]), int(st[1])
    atlas_map.next_global_id = int(atlas_map.next_global_id + n_inserted)
    atlas_map.total_count = int(atlas_map.total_count + n_inserted)
    cert = (CertBundle.create_approx(chart_id=chart_id, anchor_id=anchor_id, triggers=["insert_unfilled_budget"],
                                     frobenius_applied=False, compute=io.compute())
            if dropped > 0 else CertBundle.create_exact(chart_id=chart_id, anchor_id=anchor_id, compute=io.compute()))
    res = PrimitiveMapInsertResult(atlas_map=atlas_map, tile_id=tile_id, n_inserted=n_inserted, new_ids=ids_d, target_slots=slots_d)
    return res, cert, ExpectedEffect("primitive_map_insert_masked", float(n_inserted), float(n_inserted))


@dataclass
class PrimitiveMapCullResult:
    atlas_map: AtlasMap
    tile_id: int
    n_culled: int
    mass_dropped: float


def primitive_map_cull(atlas_map: AtlasMap, tile_id: int,
                       weight_threshold: float = constants.GC_PRIMITIVE_CULL_WEIGHT_THRESHOLD,
                       max_primitives: Optional[int] = None, chart_id: str = constants.GC_CHART_ID,
                       anchor_id: str = "primitive_map") -> Tuple[PrimitiveMapCullResult, CertBundle, ExpectedEffect]:
    """
    primitive_map_cull (fl/backend/structures/primitive_map.py:1175-1304), in place on the device pool: an absent or
    empty tile and "nothing below the threshold" are exact no-ops; otherwise the budgeting certificate with
    mass_epsilon_ratio = mass_dropped / (sum of the tile's weights + eps_mass).
    """
    tile_id = int(tile_id)

    def no_op():
        return (PrimitiveMapCullResult(atlas_map=atlas_map, tile_id=tile_id, n_culled=0, mass_dropped=0.0),
                CertBundle.create_exact(chart_id=chart_id, anchor_id=anchor_id), ExpectedEffect("primitive_map_cull", 0.0, 0.0))

    row = atlas_map.tiles.get(tile_id)
    if row is None:
        return no_op()
    io = _IO(atlas_map.device)
    stats_d = io.zeros(4)
    ca = atlas_map._c()
    io.ctx.check(io.ctx.lib.gcs_map_cull(io.ctx.handle, io.stream(), C.byref(ca), int(row), float(weight_threshold),
                                         -1 if max_primitives is None else int(max_primitives), L.ptr(stats_d)))
    st = io.host(stats_d)
    n_culled, mass_dropped = int(st[0]), float(st[1])
    if n_culled == 0:
        return no_op()
    atlas_map.total_count = atlas_map.total_count - n_culled
    cert = CertBundle.create_approx(chart_id=chart_id, anchor_id=anchor_id, triggers=["budgeting", "mass_drop"],
                                    influence=InfluenceCert.identity().with_overrides(
                                        mass_epsilon_ratio=mass_dropped / (float(st[2]) + constants.GC_EPS_MASS)),
                                    compute=io.compute())
    res = PrimitiveMapCullResult(atlas_map=atlas_map, tile_id=tile_id, n_culled=n_culled, mass_dropped=mass_dropped)
    return res, cert, ExpectedEffect("primitive_map_cull", float(n_culled), float(n_culled))


@dataclass
class PrimitiveMapForgetResult:
    atlas_map: AtlasMap
    tile_id: int


def primitive_map_forget(atlas_map: AtlasMap, tile_id: int,
                         forgetting_factor: float = constants.GC_PRIMITIVE_FORGETTING_FACTOR,
                         chart_id: str = constants.GC_CHART_ID, anchor_id: str = "primitive_map"
                         ) -> Tuple[PrimitiveMapForgetResult, CertBundle, ExpectedEffect]:
    """primitive_map_forget (fl/backend/structures/primitive_map.py:1314-1384): weights *= gamma over the tile; exact."""
    tile_id = int(tile_id)
    res = PrimitiveMapForgetResult(atlas_map=atlas_map, tile_id=tile_id)
    row = atlas_map.tiles.get(tile_id)
    if row is None:
        return res, CertBundle.create_exact(chart_id=chart_id, anchor_id=anchor_id), ExpectedEffect("primitive_map_forget", 0.0, 0.0)
    gamma = float(forgetting_factor)
    io = _IO(atlas_map.device)
    ca = atlas_map._c()
    io.ctx.check(io.ctx.lib.gcs_map_forget(io.ctx.handle, io.stream(), C.byref(ca), int(row), gamma))
    return (res, CertBundle.create_exact(chart_id=chart_id, anchor_id=anchor_id, compute=io.compute()),
            ExpectedEffect("primitive_map_forget", 1.0 - gamma, 1.0 - gamma))


# --------------------------------------------------------------------------------------------------
# the association's inner functions on their own (fl/backend/backend_node.py:884-905 calls them directly)
# --------------------------------------------------------------------------------------------------
def compute_sparse_cost_matrix(meas_positions, meas_directions, meas_kappas, map_positions, map_directions, map_kappas,
                               candidate_indices, beta: float = 0.5, eig_min: float = 1e-12) -> torch.Tensor:
    """_compute_sparse_cost_matrix_jax (primitive_association.py:152-197): (N, K) costs of the candidate pairs, on the device."""
    io = _IO()
    mp = io.dev_in(meas_positions)
    n = int(mp.shape[0])
    vp = io.dev_in(map_positions)
    m = int(vp.shape[0])
    cand = io.dev_in(candidate_indices, torch.int32)
    if cand.dim() != 2 or int(cand.shape[0]) != n:
        raise ValueError(f"candidate_indices must be ({n}, K), got {tuple(cand.shape)}")
    k = int(cand.shape[1])
    md, mk = io.dev_in(meas_directions, F64, (n, 3)), io.dev_in(meas_kappas, F64, (n,))
    vd, vk = io.dev_in(map_directions, F64, (m, 3)), io.dev_in(map_kappas, F64, (m,))
    out = io.empty(n, k)
    io.ctx.check(io.ctx.lib.gcs_sparse_cost_matrix(io.ctx.handle, io.stream(), L.ptr(mp.reshape(n, 3)), L.ptr(md), L.ptr(mk), n,
                                                   L.ptr(vp.reshape(m, 3)), L.ptr(vd), L.ptr(vk), m, L.ptr(cand), k, float(beta),
                                                   float(eig_min), L.ptr(out)))
    return out


def sinkhorn_unbalanced_fixed_k(C, a, b, epsilon: float, tau_a: float, tau_b: float, K: int) -> torch.Tensor:
    """_sinkhorn_unbalanced_fixed_k_jax (primitive_association.py:105-138): transport plan after K fixed iterations."""
    io = _IO()
    Cm = io.dev_in(C)
    if Cm.dim() != 2:
        raise ValueError(f"C must be (N, M), got {tuple(Cm.shape)}")
    n, m = int(Cm.shape[0]), int(Cm.shape[1])
    if m > 32:
        raise ValueError(f"sinkhorn_unbalanced_fixed_k: {m} columns exceed the built budget of 32 (K_ASSOC is 8)")
    av, bv = io.dev_in(a, F64, (n,)), io.dev_in(b, F64, (m,))
    pi = io.empty(n, m)
    io.ctx.check(io.ctx.lib.gcs_sinkhorn_unbalanced_fixed_k(io.ctx.handle, io.stream(), L.ptr(Cm), L.ptr(av), L.ptr(bv), n, m,
                                                            float(epsilon), float(tau_a), float(tau_b), int(K), L.ptr(pi)))
    return pi


# the reference's private names, so that its warm-up block runs unchanged after the import swap
_compute_sparse_cost_matrix_jax = compute_sparse_cost_matrix
_sinkhorn_unbalanced_fixed_k_jax = sinkhorn_unbalanced_fixed_k
