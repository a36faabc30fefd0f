"""
The per-scan hypothesis loop of the reference, batched on the device
(``fl/`` = fl_ws/src/fl_slam_poc/fl_slam_poc/ in whabacivch/GC-SLAM).

  fl/backend/backend_node.py:2036-2083   for i, belief in enumerate(self.hypotheses):
                                             result = process_scan_single_hypothesis(... primitive_map=self.primitive_map ...)
                                             if i == 0: self.primitive_map = result.primitive_map_updated
  fl/backend/pipeline.py:569-587, 780-877, 998-1010   deskew -> surfels -> recency inflate -> map view -> association ->
                                             pose evidence of one hypothesis

Every hypothesis deskews the SAME raw scan with its own twist, extracts its own surfels, and associates them with a view
of the map around its own predicted pose.  Only hypothesis 0's map update is kept (:2079-2083) -- and because it is
stored inside the loop, hypotheses 1.. see the map AFTER that update; their own recency inflation acts on a copy that is
thrown away.  ``lidar_evidence_primitives_batched`` reproduces exactly that order:

  1. hypothesis 0 runs the single-hypothesis path (``primitives.lidar_evidence_primitives``: in-place inflation + map
     update), if ``update_map``;
  2. the remaining hypotheses run as ONE batch per distinct stencil (hypotheses whose predicted positions fall into the
     same map tile share one read-only view; a hypothesis batch normally has one stencil): unit axis in the grid of every
     kernel (gcs_*_batched), the view gathered once with the inflation applied functionally
     (gcs_extract_atlas_map_view_inflated), one host synchronisation for all certificates of all hypotheses.

Results are bit-identical to calling the operators hypothesis by hypothesis (tests/test_gpu_hypothesis_batch.py).
"""

from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Dict, List, Optional

import numpy as np
import torch

from . import _lib as L
from . import constants
from .certs import CertBundle, ExpectedEffect
from .operators import _IO, DeskewConstantTwistResult, _deskew_cert, _pinned_like
from . import primitives as PR
from .primitives import (CAssocResult, CMeasBatch, CMapView, CSurfelCfg, CAssocCfg, CAtlas, OT, VP, AssociationConfig,
                         AtlasMap, AtlasMapView, MeasurementBatch, PrimitiveAssociationResult, SurfelExtractionConfig,
                         _i32arr, _i64arr)

F64 = torch.float64
_vp, _i32, _i64, _dbl, _int = C.c_void_p, C.c_int32, C.c_int64, C.c_double, C.c_int

L.register_prototypes({
    "gcs_deskew_constant_twist_batched": (_int, [_vp, _vp, _vp, _vp, _vp, _i64, _vp, _i32, _dbl, _dbl, _vp, _vp, _vp]),
    "gcs_extract_lidar_surfels_batched": (_int, [_vp, _vp, _vp, _vp, _vp, _i64, _i32, _i32, C.POINTER(CSurfelCfg),
                                                 C.POINTER(CMeasBatch), _vp]),
    "gcs_extract_atlas_map_view_inflated": (_int, [_vp, _vp, C.POINTER(CAtlas), C.POINTER(_i32), C.POINTER(_i64), _i32, _i32, _dbl,
                                                   _dbl, _i64, _dbl, _dbl, C.POINTER(CMapView), _vp, _vp]),
    "gcs_associate_primitives_ot_batched": (_int, [_vp, _vp, C.POINTER(CMeasBatch), _i32, C.POINTER(CMapView), C.POINTER(_i64), _i32,
                                                   _i32, C.POINTER(CAssocCfg), C.POINTER(CAssocResult), _vp]),
    "gcs_visual_pose_evidence_batched": (_int, [_vp, _vp, C.POINTER(CMeasBatch), _i32, C.POINTER(CMapView), C.POINTER(CAssocResult),
                                                _i32, _vp, _dbl, _dbl, _vp, _vp, _vp]),
})

# layout of one unit's row in the packed certificate buffer (float64 words)
_ROW = L.DK_NCERT + 1 + OT["NCERT"] + VP["NREC"]
_O_DK, _O_NV, _O_OT, _O_VP = 0, L.DK_NCERT, L.DK_NCERT + 1, L.DK_NCERT + 1 + OT["NCERT"]


@dataclass
class HypothesisGroup:
    """Hypotheses that share one stencil (one read-only view); stacked device results, unit axis first."""
    units: List[int]                       # hypothesis indices, in unit order
    tile_ids: List[int]
    deskewed_points: torch.Tensor          # (U, n, 3)
    deskewed_weights: torch.Tensor         # (U, n)
    batch: MeasurementBatch                # stacked: every tensor (U, N_total, ...)
    view: AtlasMapView
    association: PrimitiveAssociationResult  # stacked (U, N_total, K)
    L_pose: torch.Tensor                   # (U, 22, 22)
    h_pose: torch.Tensor                   # (U, 22)
    rec: torch.Tensor                      # (U, GCS_VP_NREC)
    scalars: Optional[np.ndarray] = None   # host copy of the packed certificate rows (U, _ROW) after the synchronisation
    view_scalars: Optional[np.ndarray] = None  # [n_valid of the view, 4 inflation statistics]


class BatchedPrimitiveEvidence:
    """
    Per-hypothesis results of ``lidar_evidence_primitives_batched``.  ``L_pose`` (H, 22, 22) / ``h_pose`` (H, 22) are
    the stacked evidence in hypothesis order (the input of fusion.evidence_fusion_batched); ``unit(h)`` assembles the
    same dict of operator tuples ``primitives.lidar_evidence_primitives`` returns for one hypothesis (certificate
    objects are only built on access).
    """

    def __init__(self, n_hyp, groups, first, cfg):
        self.n_hyp, self.groups, self.first, self._cfg = n_hyp, groups, first, cfg
        self._where: Dict[int, tuple] = {}
        for g in groups:
            for k, h in enumerate(g.units):
                self._where[h] = (g, k)
        dev = cfg["dev"]
        self.L_pose = torch.empty(n_hyp, 22, 22, dtype=F64, device=dev)
        self.h_pose = torch.empty(n_hyp, 22, dtype=F64, device=dev)
        if first is not None:
            self.L_pose[0].copy_(first["pose_evidence"][0].L_pose)
            self.h_pose[0].copy_(first["pose_evidence"][0].h_pose)
        for g in groups:
            if g.units == list(range(g.units[0], g.units[0] + len(g.units))):
                self.L_pose[g.units[0]:g.units[0] + len(g.units)].copy_(g.L_pose)
                self.h_pose[g.units[0]:g.units[0] + len(g.units)].copy_(g.h_pose)
            else:
                ix = torch.tensor(g.units, device=dev)
                self.L_pose.index_copy_(0, ix, g.L_pose)
                self.h_pose.index_copy_(0, ix, g.h_pose)

    @property
    def map_update(self):
        return None if self.first is None else self.first["map_update"]

    def unit(self, h: int) -> dict:
        if self.first is not None and h == 0:
            return self.first
        g, k = self._where[int(h)]
        c, cfg = g.scalars[k], self._cfg
        io = _IO(cfg["dev"])
        chart = cfg["chart_id"]
        N, K = g.batch.n_total, int(cfg["assoc"].k_assoc)
        dk = DeskewConstantTwistResult(points=g.deskewed_points[k], timestamps=cfg["timestamps"], weights=g.deskewed_weights[k],
                                       ess_imu=float(cfg["ess_imu"]))
        dk_cert, dk_eff = _deskew_cert(c[_O_DK:_O_DK + L.DK_NCERT], cfg["ess_imu"], chart, cfg["anchor_id"], io.compute())
        n_use = int(round(c[_O_NV]))
        b = MeasurementBatch(**{f: (getattr(g.batch, f)[k] if isinstance(getattr(g.batch, f), torch.Tensor) else getattr(g.batch, f))
                                for f in g.batch.__dataclass_fields__})
        b.n_lidar_valid = n_use
        sf_cert, sf_eff = PR._surfel_cert(n_use, cfg["surfel"], chart, "surfel_extraction", io.compute())
        view = g.view
        ri_cert, ri_eff, ri_stats = PR._inflate_finish(g.view_scalars[1:5], chart, "primitive_map_recency_inflate", io.compute())
        assoc = PrimitiveAssociationResult(**{f: getattr(g.association, f)[k] for f in g.association.__dataclass_fields__})
        if b.n_valid == 0 or view.n_valid == 0:      # the reference's early exits (primitive_association.py:275-290)
            assoc = PR._empty_assoc(io, N, K)
            as_out = (assoc, CertBundle.create_exact(chart_id=chart, anchor_id="primitive_ot"),
                      ExpectedEffect("primitive_association_ot", 0.0, 0.0))
            pe_out = PR._empty_pose_evidence(io, cfg["eps_lift"], chart, "visual_pose_evidence")
        else:
            as_out = (assoc,) + PR._assoc_cert(c[_O_OT:_O_OT + OT["NCERT"]], N, K, cfg["assoc"], chart, "primitive_ot", io)
            pe_out = PR._pose_evidence_finish(c[_O_VP:_O_VP + VP["NREC"]], g.L_pose[k], g.h_pose[k], g.rec[k], b.n_valid, K,
                                              cfg["eps_lift"], chart, "visual_pose_evidence", io.compute())
        return dict(deskew=(dk, dk_cert, dk_eff), surfels=(b, sf_cert, sf_eff),
                    recency_inflate=(cfg["atlas"], ri_cert, ri_eff, ri_stats), map_view=view, association=as_out,
                    pose_evidence=pe_out, map_update=None)


def _stack_batch(io, base: Optional[MeasurementBatch], U: int, cfg: SurfelExtractionConfig) -> MeasurementBatch:
    """U copies of the camera slice (base_batch) stacked along a new unit axis; LiDAR rows zero."""
    if base is None:
        base = PR.create_empty_measurement_batch(cfg.n_feat, cfg.n_surfel, io.dev)
    if base.n_surfel != cfg.n_surfel or base.n_feat != cfg.n_feat:
        raise ValueError("lidar_evidence_primitives_batched: base_batch budget differs from the surfel config")
    kw = {}
    for f in base.__dataclass_fields__:
        v = getattr(base, f)
        kw[f] = v.unsqueeze(0).expand((U,) + tuple(v.shape)).contiguous() if isinstance(v, torch.Tensor) else v
    return MeasurementBatch(**kw)


def _run_group(io, units, tile_ids, pts, t, w, n, xi_d, poses_d, t0, t1, atlas_map, scan_seq, base_batch, scfg, acfg, m_tile_view,
               eps_lift, eps_mass, min_scale) -> HypothesisGroup:
    U = len(units)
    lib, h, st = io.ctx.lib, io.ctx.handle, io.stream()
    N, K = scfg.n_feat + scfg.n_surfel, int(acfg.k_assoc)
    sel = None if units == list(range(units[0], units[0] + U)) else torch.tensor(units, device=io.dev)
    xi_g = xi_d[units[0]:units[0] + U] if sel is None else xi_d.index_select(0, sel)
    po_g = poses_d[units[0]:units[0] + U] if sel is None else poses_d.index_select(0, sel)
    scal = io.empty(U, _ROW)
    vscal = io.zeros(8)
    dk_cert = io.empty(U, L.DK_NCERT)
    dk_p, dk_w = io.empty(U, n, 3), io.empty(U, n)
    io.ctx.check(lib.gcs_deskew_constant_twist_batched(h, st, L.ptr(pts), L.ptr(t), L.ptr(w), n, L.ptr(xi_g.contiguous()), U,
                                                       float(t0), float(t1), L.ptr(dk_p), L.ptr(dk_w), L.ptr(dk_cert)))
    batch = _stack_batch(io, base_batch, U, scfg)
    nv_d = io.zeros(U, dtype=torch.int32)
    cb, cc = batch._c(), scfg._c()
    io.ctx.check(lib.gcs_extract_lidar_surfels_batched(h, st, L.ptr(dk_p), L.ptr(t), L.ptr(dk_w), n, U, 1, C.byref(cc), C.byref(cb),
                                                       L.ptr(nv_d)))
    view = PR._empty_view(io, tile_ids, m_tile_view)
    nvv = vscal[5:6].view(torch.int32)        # 2 int32 words inside the packed buffer; [0] = n_valid of the view
    ca, cv = atlas_map._c(), view._c()
    idx = atlas_map.index_list(tile_ids, create=False)
    io.ctx.check(lib.gcs_extract_atlas_map_view_inflated(h, st, C.byref(ca), _i32arr(idx), _i64arr(tile_ids), len(tile_ids),
                                                         int(m_tile_view), float(eps_lift), float(eps_mass), int(scan_seq),
                                                         float(acfg.recency_decay_lambda), float(min_scale), C.byref(cv),
                                                         L.ptr(nvv), L.ptr(vscal[0:4])))
    z = io.empty
    assoc = PrimitiveAssociationResult(responsibilities=z(U, N, K), candidate_pool_indices=z(U, N, K, dtype=torch.int32),
                                       candidate_tile_ids=z(U, N, K, dtype=torch.int64), candidate_slots=z(U, N, K, dtype=torch.int64),
                                       row_masses=z(U, N), cost_matrix=z(U, N, K))
    ot_cert = io.empty(U, OT["NCERT"])
    cfg_c = PR._c_assoc_cfg(acfg, eps_lift)
    cr = assoc._c()
    io.ctx.check(lib.gcs_associate_primitives_ot_batched(h, st, C.byref(cb), U, C.byref(cv), _i64arr(tile_ids), len(tile_ids),
                                                         int(m_tile_view), C.byref(cfg_c), C.byref(cr), L.ptr(ot_cert)))
    L22, h22, rec = io.empty(U, 22, 22), io.empty(U, 22), io.empty(U, VP["NREC"])
    io.ctx.check(lib.gcs_visual_pose_evidence_batched(h, st, C.byref(cb), U, C.byref(cv), C.byref(cr), K, L.ptr(po_g.contiguous()),
                                                      float(eps_lift), float(eps_mass), L.ptr(L22), L.ptr(h22), L.ptr(rec)))
    # one packed row of certificate scalars per unit
    scal[:, _O_DK:_O_DK + L.DK_NCERT] = dk_cert
    scal[:, _O_NV] = nv_d.to(F64)
    scal[:, _O_OT:_O_OT + OT["NCERT"]] = ot_cert
    scal[:, _O_VP:_O_VP + VP["NREC"]] = rec
    g = HypothesisGroup(units=list(units), tile_ids=[int(x) for x in tile_ids], deskewed_points=dk_p, deskewed_weights=dk_w,
                        batch=batch, view=view, association=assoc, L_pose=L22, h_pose=h22, rec=rec)
    g._scal_d, g._vscal_d = scal, vscal
    return g


def lidar_evidence_primitives_batched(points, timestamps, weights, scan_start_time: float, scan_end_time: float, xi_bodies,
                                      atlas_map: AtlasMap, poses_pred, scan_seq: int,
                                      base_batch: Optional[MeasurementBatch] = None,
                                      surfel_config: Optional[SurfelExtractionConfig] = None,
                                      association_config: Optional[AssociationConfig] = None,
                                      m_tile_view: int = constants.GC_M_TILE_VIEW, ess_imu: float = 1.0, update_map: bool = True,
                                      map_update_kwargs: Optional[dict] = None, eps_lift: float = constants.GC_EPS_LIFT,
                                      eps_mass: float = constants.GC_EPS_MASS,
                                      recency_min_scale: float = constants.GC_RECENCY_MIN_SCALE,
                                      chart_id: str = constants.GC_CHART_ID,
                                      anchor_id: str = "lidar_evidence_primitives") -> BatchedPrimitiveEvidence:
    """
    Primitive-family LiDAR evidence of H pose hypotheses of one scan (module docstring).  ``xi_bodies`` (H, 6): twist of
    every hypothesis (host array or device tensor, e.g. the rows gcs_imu_scan_twist wrote); ``poses_pred`` (H, 6)
    [t, rotvec] predicted world poses (map stencil centre + linearisation point).  ``update_map``: hypothesis 0 updates
    the map first (the reference's order); False: every hypothesis sees the current map read-only.
    """
    io = _IO(atlas_map.device)
    pts = io.dev_in(points, shape=(-1, 3))
    n = int(pts.shape[0])
    t = io.dev_in(timestamps, shape=(-1,))
    w = io.dev_in(weights, shape=(-1,))
    if t.shape[0] != n or w.shape[0] != n:
        raise ValueError("lidar_evidence_primitives_batched: points/timestamps/weights length mismatch")
    xi_d = io.dev_in(xi_bodies, shape=(-1, 6))
    H = int(xi_d.shape[0])
    poses_h = (poses_pred.detach().cpu().numpy() if isinstance(poses_pred, torch.Tensor) else np.asarray(poses_pred, np.float64)).reshape(-1, 6)
    if poses_h.shape[0] != H or H < 1:
        raise ValueError(f"lidar_evidence_primitives_batched: {H} twists but {poses_h.shape[0]} poses")
    poses_d = io.dev_in(poses_pred, shape=(-1, 6))
    scfg = surfel_config if surfel_config is not None else SurfelExtractionConfig()
    acfg = association_config if association_config is not None else AssociationConfig(scan_seq=int(scan_seq))
    PR._check_assoc_config(acfg)
    first = None
    rest = list(range(H))
    if update_map:
        active0 = PR.ma_hex_stencil_tile_ids(poses_h[0, :3], acfg.h_tile, acfg.r_stencil_tiles_xy, acfg.r_stencil_tiles_z)
        xi0 = xi_d[0].detach().cpu().numpy()
        first = PR.lidar_evidence_primitives(pts, t, w, scan_start_time, scan_end_time, xi0, atlas_map, active0, poses_h[0],
                                             scan_seq, base_batch=base_batch, surfel_config=scfg, association_config=acfg,
                                             m_tile_view=m_tile_view, ess_imu=ess_imu, update_map=True,
                                             map_update_kwargs=map_update_kwargs, chart_id=chart_id, anchor_id=anchor_id)
        rest = rest[1:]
    by_stencil: Dict[tuple, List[int]] = {}
    for hh in rest:
        tl = tuple(PR.ma_hex_stencil_tile_ids(poses_h[hh, :3], acfg.h_tile, acfg.r_stencil_tiles_xy, acfg.r_stencil_tiles_z))
        by_stencil.setdefault(tl, []).append(hh)
    groups = [_run_group(io, units, list(tl), pts, t, w, n, xi_d, poses_d, scan_start_time, scan_end_time, atlas_map, scan_seq,
                         base_batch, scfg, acfg, m_tile_view, eps_lift, eps_mass, recency_min_scale)
              for tl, units in by_stencil.items()]
    cfg = dict(dev=io.dev, chart_id=chart_id, anchor_id=anchor_id, ess_imu=ess_imu, surfel=scfg, assoc=acfg, eps_lift=eps_lift,
               atlas=atlas_map, timestamps=t)
    out = BatchedPrimitiveEvidence(H, groups, first, cfg)
    # ONE synchronisation for the certificates of every hypothesis
    bufs = []
    for k, g in enumerate(groups):
        b1, b2 = _pinned_like(g._scal_d, ("hb", k, 0)), _pinned_like(g._vscal_d, ("hb", k, 1))
        b1.copy_(g._scal_d.reshape(-1), non_blocking=True)
        b2.copy_(g._vscal_d, non_blocking=True)
        bufs.append((b1, b2))
    if groups:
        torch.cuda.current_stream(io.dev).synchronize()
    for g, (b1, b2) in zip(groups, bufs):
        g.scalars = b1.numpy().reshape(len(g.units), _ROW).copy()
        vs = b2.numpy().copy()
        g.view_scalars = np.concatenate([[float(vs[5:6].view(np.int32)[0])], vs[0:4]])
        g.view.n_valid = int(g.view_scalars[0])
        g.batch.n_lidar_valid = -1             # per unit: see unit(h)
        # the reference's early exits (empty measurement batch / empty view): evidence = eps_lift * I, h = 0
        n_cam = g.batch.n_camera_valid
        for k, hh in enumerate(g.units):
            if g.view.n_valid == 0 or n_cam + int(round(g.scalars[k, _O_NV])) == 0:
                out.L_pose[hh] = eps_lift * torch.eye(22, dtype=F64, device=io.dev)
                out.h_pose[hh] = 0.0
    return out
