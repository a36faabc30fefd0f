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
import threading
from typing import Dict, List, Optional

import numpy as np
import torch

from . import _lib as L
from . import constants
from .certs import CertBundle, ExpectedEffect
from .operators import _IO, DeskewConstantTwistResult, _deskew_cert
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
                                                _i32, _vp, _dbl, _dbl, _vp, _vp, _vp, _vp, _i32, _vp]),
})

# layout of one unit's row in the packed certificate buffer (float64 words)
_ROW = L.DK_NCERT + 1 + OT["NCERT"] + VP["NREC"]
_O_DK, _O_NV, _O_OT, _O_VP = 0, L.DK_NCERT, L.DK_NCERT + 1, L.DK_NCERT + 1 + OT["NCERT"]


class HypothesisGroup:
    """
    Hypotheses that share one stencil (one read-only view).  The stacked device results (unit axis first) live in one
    arena allocation; the tensor views below are made on first access, not per call.
      deskewed_points (U, n, 3), deskewed_weights (U, n), batch: MeasurementBatch with every tensor (U, N_total, ...),
      view: AtlasMapView, association: PrimitiveAssociationResult (U, N_total, K), L_pose (U, 22, 22), h_pose (U, 22),
      rec (U, GCS_VP_NREC)
    """

    def __init__(self, units, tile_ids, arena, n_scalar_bytes, scfg, n_camera_valid, m_tile_view):
        self.units, self.tile_ids, self._arena = list(units), [int(x) for x in tile_ids], arena
        self._scal_d = arena.buf[:n_scalar_bytes]
        self._scfg, self._n_cam, self._m_view = scfg, int(n_camera_valid), int(m_tile_view)
        self.scalars: Optional[np.ndarray] = None       # host copy of the packed certificate rows (U, _ROW) after wait()
        self.view_scalars: Optional[np.ndarray] = None  # [n_valid of the view, 4 inflation statistics]
        self._view_n_valid = -1
        self._cache = {}

    def _lazy(self, key, make):
        if key not in self._cache:
            self._cache[key] = make()
        return self._cache[key]

    deskewed_points = property(lambda self: self._lazy("dk_pts", lambda: self._arena.view("dk_pts")))
    deskewed_weights = property(lambda self: self._lazy("dk_w", lambda: self._arena.view("dk_w")))
    L_pose = property(lambda self: self._lazy("L22", lambda: self._arena.view("L22")))
    h_pose = property(lambda self: self._lazy("h22", lambda: self._arena.view("h22")))
    rec = property(lambda self: self._lazy("rec", lambda: self._arena.view("rec")))

    @property
    def batch(self) -> MeasurementBatch:
        return self._lazy("batch", lambda: MeasurementBatch(
            **{f: self._arena.view("b_" + f) for f, _, _ in _BATCH_FIELDS}, n_feat=self._scfg.n_feat, n_surfel=self._scfg.n_surfel,
            n_camera_valid=self._n_cam, n_lidar_valid=-1))

    @property
    def view(self) -> AtlasMapView:
        v = self._lazy("view", lambda: AtlasMapView(**{f: self._arena.view("v_" + f) for f, _, _ in _VIEW_FIELDS},
                                                    tile_ids=list(self.tile_ids), m_tile_view=self._m_view))
        v.n_valid = self._view_n_valid
        return v

    @property
    def association(self) -> PrimitiveAssociationResult:
        return self._lazy("assoc", lambda: PrimitiveAssociationResult(**{f: self._arena.view("a_" + f) for f, _, _ in _ASSOC_FIELDS}))


class BatchedPrimitiveEvidence:
    """
    Per-hypothesis results of ``lidar_evidence_primitives_batched``.  ``L_pose`` (H, 22, 22) / ``h_pose`` (H, 22) are
    the stacked evidence in hypothesis order (the input of fusion.evidence_fusion_batched); ``unit(h)`` assembles the
    same dict of operator tuples ``primitives.lidar_evidence_primitives`` returns for one hypothesis (certificate
    objects are only built on access); ``map_update`` is hypothesis 0's (result, cert, effect) when the map was updated.
    """

    def __init__(self, n_hyp, groups, cfg):
        self.n_hyp, self.groups, self._cfg = n_hyp, groups, cfg
        self._pending = None
        self._owned = []
        self._gens = {}          # "inflate" / "update": (generator at its yield, pinned host buffer, shape)
        self._first_extra = None
        self._where: Dict[int, tuple] = {}
        for g in groups:
            for k, h in enumerate(g.units):
                self._where[h] = (g, k)
        dev = cfg["dev"]
        if len(groups) == 1 and groups[0].units == list(range(n_hyp)):
            self.L_pose, self.h_pose = groups[0].L_pose, groups[0].h_pose    # one stencil: the group's arrays are the stack
            return
        self.L_pose = torch.empty(n_hyp, 22, 22, dtype=F64, device=dev)
        self.h_pose = torch.empty(n_hyp, 22, dtype=F64, device=dev)
        for g in groups:
            if g.units == list(range(g.units[0], g.units[0] + len(g.units))):
                self.L_pose[g.units[0]:g.units[0] + len(g.units)].copy_(g.L_pose)
                self.h_pose[g.units[0]:g.units[0] + len(g.units)].copy_(g.h_pose)
            else:
                ix = torch.tensor(g.units, device=dev)
                self.L_pose.index_copy_(0, ix, g.L_pose)
                self.h_pose.index_copy_(0, ix, g.h_pose)

    def wait(self):
        """Block until the certificates of every hypothesis (and the map update's) are on the host (idempotent)."""
        if self._pending is None:
            return self
        self._event.synchronize()
        for g, b1 in zip(self.groups, self._pending):
            _unpack_scalars(g, b1.numpy().copy())
            g._view_n_valid = int(g.view_scalars[0])
        self._pending = None
        if self._gens:
            fin = {}
            for name in ("inflate", "update"):      # the update's certificate reads the inflation statistics
                gen, buf, shape = self._gens[name]
                try:
                    gen.send(buf.view(F64).numpy().reshape(shape).copy())
                    raise RuntimeError("operator body yielded twice")
                except StopIteration as e:
                    fin[name] = e.value
                if name == "inflate":
                    self._inflate_stats = fin[name][3]
            self._first_extra = fin
            self._gens = {}
            if self._cfg["atlas"]._pending_update is self:
                self._cfg["atlas"]._pending_update = None
        for buf in getattr(self, "_owned", []):
            _PinnedRing.release(buf)
        self._owned = []
        return self

    def __del__(self):      # a deferred result dropped without wait(): its pinned buffers go back once the copies are done
        try:
            if self._owned:
                self._event.synchronize()
                for buf in self._owned:
                    _PinnedRing.release(buf)
        except Exception:
            pass

    @property
    def map_update(self):
        self.wait()
        return None if self._first_extra is None else self._first_extra["update"]

    def unit(self, h: int) -> dict:
        self.wait()
        g, k = self._where[int(h)]
        c, cfg = g.scalars[k], self._cfg
        io = _IO(cfg["dev"])
        chart = cfg["chart_id"]
        N, K = g.batch.n_total, int(cfg["assoc"].k_assoc)
        dk = DeskewConstantTwistResult(points=g.deskewed_points[k], timestamps=cfg["timestamps"], weights=g.deskewed_weights[k],
                                       ess_imu=float(cfg["ess_imu"]))
        dk_cert, dk_eff = _deskew_cert(c[_O_DK:_O_DK + L.DK_NCERT], cfg["ess_imu"], chart, cfg["anchor_id"], io.compute())
        n_use = int(round(c[_O_NV]))
        b = _batch_unit(g.batch, k)
        b.n_lidar_valid = n_use
        sf_cert, sf_eff = PR._surfel_cert(n_use, cfg["surfel"], chart, "surfel_extraction", io.compute())
        view = g.view
        updating = self._first_extra is not None and int(h) == 0
        if updating:         # hypothesis 0 inflated the map in place: that operator's own tuple
            ri = self._first_extra["inflate"]
        else:
            ri = (cfg["atlas"],) + PR._inflate_finish(g.view_scalars[1:5], chart, "primitive_map_recency_inflate", io.compute())
        assoc = PrimitiveAssociationResult(**{f: getattr(g.association, f)[k] for f in g.association.__dataclass_fields__})
        if b.n_valid == 0 or view.n_valid == 0:      # the reference's early exits (primitive_association.py:275-290)
            as_out = (assoc, CertBundle.create_exact(chart_id=chart, anchor_id="primitive_ot"),
                      ExpectedEffect("primitive_association_ot", 0.0, 0.0))
            pe_out = PR._empty_pose_evidence(io, cfg["eps_lift"], chart, "visual_pose_evidence")
        else:
            as_out = (assoc,) + PR._assoc_cert(c[_O_OT:_O_OT + OT["NCERT"]], N, K, cfg["assoc"], chart, "primitive_ot", io)
            pe_out = PR._pose_evidence_finish(c[_O_VP:_O_VP + VP["NREC"]], g.L_pose[k], g.h_pose[k], g.rec[k], b.n_valid, K,
                                              cfg["eps_lift"], chart, "visual_pose_evidence", io.compute())
        return dict(deskew=(dk, dk_cert, dk_eff), surfels=(b, sf_cert, sf_eff), recency_inflate=ri, map_view=view,
                    association=as_out, pose_evidence=pe_out, map_update=self._first_extra["update"] if updating else None)


def _batch_unit(stacked: MeasurementBatch, k: int) -> MeasurementBatch:
    return MeasurementBatch(**{f: (getattr(stacked, f)[k] if isinstance(getattr(stacked, f), torch.Tensor) else getattr(stacked, f))
                               for f in stacked.__dataclass_fields__})


class CPrimBatchArgs(C.Structure):
    """gcs_prim_batch_args (include/gcs_b200.h)."""
    _fields_ = [("pts", _vp), ("t", _vp), ("w", _vp), ("n", _i64), ("n_units", _i32), ("inflate", _i32), ("xi", _vp), ("poses", _vp),
                ("scan_start_time", _dbl), ("scan_end_time", _dbl), ("dk_pts", _vp), ("dk_w", _vp), ("dk_cert", _vp),
                ("surfel_cfg", CSurfelCfg), ("base", CMeasBatch), ("batch", CMeasBatch), ("n_lidar_valid", _vp), ("n_camera_valid", _i32), ("reserved_", _i32),
                ("atlas", C.POINTER(CAtlas)), ("tile_index", _i32 * 16), ("tile_ids", _i64 * 16), ("n_tiles", _i32),
                ("m_tile_view", _i32), ("eps_lift", _dbl), ("eps_mass", _dbl), ("recency_min_scale", _dbl), ("view", CMapView),
                ("view_n_valid", _vp), ("inflate_stats", _vp), ("assoc_cfg", CAssocCfg), ("assoc", CAssocResult), ("ot_cert", _vp),
                ("L22", _vp), ("h22", _vp), ("rec", _vp)]


L.register_prototypes({"gcs_lidar_evidence_primitives_batched": (_int, [_vp, _vp, C.POINTER(CPrimBatchArgs)])})

_DT = {F64: 8, torch.int32: 4, torch.int64: 8, torch.uint8: 1}
import os as _os
_NO_SIDE_ROUTE = bool(_os.environ.get("GCS_NO_SIDE_ROUTE"))   # developer switch: keep the map update on the caller's stream


class _Arena:
    """One device allocation per call, carved into the stacked result arrays; tensor views are made on access only.
    The layout (name -> offset, shape, dtype) depends on the shapes alone and is computed once per shape key."""
    _layouts: Dict[tuple, tuple] = {}

    def __init__(self, dev, spec=None, size=0):
        self.dev, self.spec, self.size, self.buf = dev, (spec if spec is not None else {}), size, None

    def add(self, name, shape, dtype=F64):
        n = 1
        for x in shape:
            n *= int(x)
        nbytes = n * _DT[dtype]
        self.spec[name] = (self.size, tuple(int(x) for x in shape), dtype, nbytes)
        self.size += (nbytes + 255) & ~255

    def alloc(self):
        self.buf = torch.empty(self.size, dtype=torch.uint8, device=self.dev)
        self.base = self.buf.data_ptr()

    def ptr(self, name):
        return _vp(self.base + self.spec[name][0])

    def view(self, name):
        off, shape, dtype, nbytes = self.spec[name]
        return self.buf[off:off + nbytes].view(dtype).reshape(shape)


_BATCH_FIELDS = (("Lambdas", (3, 3), F64), ("thetas", (3,), F64), ("etas", (constants.GC_VMF_N_LOBES, 3), F64), ("weights", (), F64),
                 ("sources", (), torch.int32), ("source_indices", (), torch.int32), ("valid_mask", (), torch.uint8),
                 ("timestamps", (), F64), ("colors", (3,), F64))
_VIEW_FIELDS = (("candidate_tile_ids", (), torch.int64), ("candidate_slots", (), torch.int32), ("valid_mask", (), torch.uint8),
                ("positions", (3,), F64), ("covariances", (3, 3), F64), ("directions", (3,), F64), ("kappas", (), F64),
                ("weights", (), F64), ("primitive_ids", (), torch.int64), ("last_supported_scan_seq", (), torch.int64),
                ("etas", (constants.GC_VMF_N_LOBES, 3), F64), ("colors", (3,), F64))
_ASSOC_FIELDS = (("responsibilities", True, F64), ("candidate_pool_indices", True, torch.int32), ("candidate_tile_ids", True, torch.int64),
                 ("candidate_slots", True, torch.int64), ("row_masses", False, F64), ("cost_matrix", True, F64))
_C_BATCH_NAME = dict(valid_mask="valid")
_C_VIEW_NAME = dict(valid_mask="valid")


_PTR_SLOTS: Dict[tuple, tuple] = {}


def _arena_pointer_slots(key, spec):
    """(word indices into CPrimBatchArgs, arena offsets) of every struct field that points into the arena; per layout."""
    hit = _PTR_SLOTS.get(key)
    if hit is not None:
        return hit
    T = CPrimBatchArgs
    pairs = [(T.dk_pts.offset, "dk_pts"), (T.dk_w.offset, "dk_w"), (T.dk_cert.offset, "dk_cert"), (T.n_lidar_valid.offset, "n_valid"),
             (T.view_n_valid.offset, "view_n_valid"), (T.inflate_stats.offset, "inflate_stats"), (T.ot_cert.offset, "ot_cert"),
             (T.L22.offset, "L22"), (T.h22.offset, "h22"), (T.rec.offset, "rec")]
    for f, _, _ in _BATCH_FIELDS:
        pairs.append((T.batch.offset + getattr(CMeasBatch, _C_BATCH_NAME.get(f, f)).offset, "b_" + f))
    for f, _, _ in _VIEW_FIELDS:
        pairs.append((T.view.offset + getattr(CMapView, _C_VIEW_NAME.get(f, f)).offset, "v_" + f))
    for f, _, _ in _ASSOC_FIELDS:
        pairs.append((T.assoc.offset + getattr(CAssocResult, f).offset, "a_" + f))
    assert all(o % 8 == 0 for o, _ in pairs)
    words = np.array([o // 8 for o, _ in pairs], dtype=np.int64)
    offs = np.array([spec[name][0] for _, name in pairs], dtype=np.uint64)
    _PTR_SLOTS[key] = (words, offs)
    return words, offs


def _run_group(io, units, tile_ids, pts, t, w, n, xi_d, poses_d, t0, t1, atlas_map, scan_seq, base_batch, scfg, acfg, m_tile_view,
               eps_lift, eps_mass, min_scale, inflate=True) -> HypothesisGroup:
    U = len(units)
    N, K = scfg.n_feat + scfg.n_surfel, int(acfg.k_assoc)
    P = len(tile_ids) * int(m_tile_view)
    if base_batch is None:
        base_batch = PR.create_empty_measurement_batch(scfg.n_feat, scfg.n_surfel, io.dev)
    if base_batch.n_surfel != scfg.n_surfel or base_batch.n_feat != scfg.n_feat:
        raise ValueError("lidar_evidence_primitives_batched: base_batch budget differs from the surfel config")
    if units == list(range(units[0], units[0] + U)):
        xi_g, po_g = xi_d[units[0]:units[0] + U], poses_d[units[0]:units[0] + U]
    else:
        sel = torch.tensor(units, device=io.dev)
        xi_g, po_g = xi_d.index_select(0, sel), poses_d.index_select(0, sel)
    key = (U, n, N, P, K)
    lay = _Arena._layouts.get(key)
    if lay is None:
        A = _Arena(io.dev)
        # packed certificate scalars first (one contiguous device -> host copy): _ROW float64 words per unit, then the view's
        A.add("dk_cert", (U, L.DK_NCERT)); A.add("n_valid", (U,), torch.int32); A.add("ot_cert", (U, OT["NCERT"])); A.add("rec", (U, VP["NREC"]))
        A.add("view_n_valid", (2,), torch.int32); A.add("inflate_stats", (4,))
        n_scalar_bytes = A.size
        A.add("dk_pts", (U, n, 3)); A.add("dk_w", (U, n))
        for f, shp, dt in _BATCH_FIELDS:
            A.add("b_" + f, (U, N) + shp, dt)
        for f, shp, dt in _VIEW_FIELDS:
            A.add("v_" + f, (P,) + shp, dt)
        for f, has_k, dt in _ASSOC_FIELDS:
            A.add("a_" + f, (U, N, K) if has_k else (U, N), dt)
        A.add("L22", (U, 22, 22)); A.add("h22", (U, 22))
        lay = _Arena._layouts[key] = (A.spec, A.size, n_scalar_bytes)
    spec, size, n_scalar_bytes = lay
    A = _Arena(io.dev, spec, size)
    A.alloc()
    A.buf[:n_scalar_bytes].zero_()
    a = CPrimBatchArgs()
    # every pointer into the arena in one vectorised store: (8-byte word of the struct) <- arena base + offset
    words, offs = _arena_pointer_slots(key, spec)
    np.frombuffer(a, dtype=np.uint64)[words] = offs + np.uint64(A.base)
    a.pts, a.t, a.w, a.n, a.n_units, a.inflate = L.ptr(pts), L.ptr(t), L.ptr(w), n, U, 1 if inflate else 0
    a.xi, a.poses = L.ptr(xi_g.contiguous()), L.ptr(po_g.contiguous())
    a.scan_start_time, a.scan_end_time = float(t0), float(t1)
    a.surfel_cfg = scfg._c()
    a.base = base_batch._c()
    a.batch.n_feat, a.batch.n_surfel = scfg.n_feat, scfg.n_surfel
    a.n_camera_valid = int(base_batch.n_camera_valid)
    ca = atlas_map._c()
    a.atlas = C.pointer(ca)
    idx = atlas_map.index_list(tile_ids, create=False)
    for k in range(len(tile_ids)):
        a.tile_index[k], a.tile_ids[k] = int(idx[k]), int(tile_ids[k])
    a.n_tiles, a.m_tile_view = len(tile_ids), int(m_tile_view)
    a.eps_lift, a.eps_mass, a.recency_min_scale = float(eps_lift), float(eps_mass), float(min_scale)
    a.assoc_cfg = PR._c_assoc_cfg(acfg, eps_lift)
    io.ctx.check(io.ctx.lib.gcs_lidar_evidence_primitives_batched(io.ctx.handle, io.stream(), C.byref(a)))
    del xi_g, po_g
    return HypothesisGroup(units, tile_ids, A, n_scalar_bytes, scfg, base_batch.n_camera_valid, m_tile_view)


def _unpack_scalars(g: HypothesisGroup, host_bytes: np.ndarray):
    """Host copy of the arena's scalar region -> per-unit rows in the _ROW layout + the view's scalars."""
    A, U = g._arena, len(g.units)

    def part(name, dtype):
        off, shape, _, nbytes = A.spec[name]
        return host_bytes[off:off + nbytes].view(dtype).reshape(shape)
    rows = np.empty((U, _ROW))
    rows[:, _O_DK:_O_DK + L.DK_NCERT] = part("dk_cert", np.float64)
    rows[:, _O_NV] = part("n_valid", np.int32)
    rows[:, _O_OT:_O_OT + OT["NCERT"]] = part("ot_cert", np.float64)
    rows[:, _O_VP:_O_VP + VP["NREC"]] = part("rec", np.float64)
    g.scalars = rows
    g.view_scalars = np.concatenate([[float(part("view_n_valid", np.int32)[0])], part("inflate_stats", np.float64)])


class _PinnedRing:
    """Pinned host buffers per (thread, size): staging for asynchronous copies in both directions (a pageable cudaMemcpy
    waits for everything enqueued before it, which would serialise back-to-back scans).  A buffer handed out is busy until
    its owner releases it (a deferred result does so in wait()); a new one is only allocated when all are busy, so the
    steady state of a double-buffered loop touches two or three buffers and never allocates."""
    _tls = threading.local()

    @classmethod
    def get(cls, nbytes: int) -> torch.Tensor:
        pool = getattr(cls._tls, "pool", None)
        if pool is None:
            pool = cls._tls.pool = {}
        slots = pool.setdefault(int(nbytes), [])
        for slot in slots:
            if not slot[1]:
                slot[1] = True
                return slot[0]
        slots.append([torch.empty(int(nbytes), dtype=torch.uint8).pin_memory(), True])
        return slots[-1][0]

    @classmethod
    def release(cls, buf: torch.Tensor):
        for slot in getattr(cls._tls, "pool", {}).get(int(buf.numel()), []):
            if slot[0] is buf:
                slot[1] = False
                return


def _h2d_async(io, a: np.ndarray, shape, owned: list) -> torch.Tensor:
    """Host array -> device through a pinned staging buffer; the buffer joins `owned` and is released with the result."""
    a = np.ascontiguousarray(a, dtype=np.float64)
    stage = _PinnedRing.get(a.nbytes)
    owned.append(stage)
    stage.view(F64).copy_(torch.from_numpy(a).reshape(-1))
    io.h2d += a.nbytes
    return stage.view(F64).to(io.dev, non_blocking=True).reshape(shape)


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
                                      anchor_id: str = "lidar_evidence_primitives", defer: bool = False,
                                      z_lin_poses=None, z_t=None) -> BatchedPrimitiveEvidence:
    """
    Primitive-family LiDAR evidence of H pose hypotheses of one scan (module docstring).  ``xi_bodies`` (H, 6): twist of
    every hypothesis (host array or device tensor, e.g. the rows gcs_imu_scan_twist wrote); ``poses_pred`` (H, 6)
    [t, rotvec] predicted world poses: the centre of the map stencil (pipeline.py:802-829) and, unless ``z_lin_poses``
    (H, 6) gives the IMU + odometry informed points of pipeline.py:998-1008, the linearisation points of the pose
    evidence.  ``update_map``: hypothesis 0 updates the map first (the reference's order) with the pose ``z_t`` (6,)
    (pipeline.py:1244: the post-recompose pose; default: hypothesis 0's linearisation pose -- a caller that needs the
    fused pose there runs hypothesis 0 with ``update_map=False`` and ``primitives.map_update_step12b`` itself); False:
    every hypothesis sees the current map read-only.  ``defer``: return as soon as everything is enqueued; ``out.wait()``
    (or the first ``out.unit(h)``) blocks for the certificates.
    """
    io = _IO(atlas_map.device)
    pts = io.dev_in(points, shape=(-1, 3))
    n = int(pts.shape[0])
    t = io.dev_in(timestamps, shape=(-1,))
    w = io.dev_in(weights, shape=(-1,))
    if t.shape[0] != n or w.shape[0] != n:
        raise ValueError("lidar_evidence_primitives_batched: points/timestamps/weights length mismatch")
    owned: list = []      # pinned buffers of this call, released in wait()
    xi_d = (io.dev_in(xi_bodies, shape=(-1, 6)) if isinstance(xi_bodies, torch.Tensor)
            else _h2d_async(io, np.asarray(xi_bodies, np.float64).reshape(-1, 6), (-1, 6), owned))
    H = int(xi_d.shape[0])
    # the stencil of every hypothesis is decided on the host (as the reference does, pipeline.py:808-829): poses given as a
    # device tensor cost one blocking read here -- pass host poses to keep back-to-back scans asynchronous
    poses_h = (poses_pred.detach().cpu().numpy() if isinstance(poses_pred, torch.Tensor) else np.asarray(poses_pred, np.float64)).reshape(-1, 6)
    if poses_h.shape[0] != H or H < 1:
        raise ValueError(f"lidar_evidence_primitives_batched: {H} twists but {poses_h.shape[0]} poses")
    lin = poses_pred if z_lin_poses is None else z_lin_poses
    lin_h = None
    if isinstance(lin, torch.Tensor) and lin.is_cuda:
        poses_d = io.dev_in(lin, shape=(-1, 6))
    else:
        lin_h = poses_h if z_lin_poses is None else (lin.detach().cpu().numpy() if isinstance(lin, torch.Tensor)
                                                     else np.asarray(lin, np.float64)).reshape(-1, 6)
        poses_d = _h2d_async(io, lin_h, (-1, 6), owned)
    if int(poses_d.shape[0]) != H:
        raise ValueError(f"lidar_evidence_primitives_batched: {H} twists but {int(poses_d.shape[0])} linearisation poses")
    scfg = surfel_config if surfel_config is not None else SurfelExtractionConfig()
    acfg = association_config if association_config is not None else AssociationConfig(scan_seq=int(scan_seq))
    PR._check_assoc_config(acfg)
    # A deferred earlier scan still owes the map's host state (id counter, counts).  Its update keeps the id counter on the
    # device, so ANOTHER deferred scan with an update can be enqueued behind it without that state (results are then
    # waited for in order); anything else waits first.
    pend = getattr(atlas_map, "_pending_update", None)
    if pend is not None and not (update_map and defer and pend._pending is not None and getattr(pend, "_ids_on_device", False)):
        pend.wait()
        pend = None
    elif pend is not None and pend._pending is None:
        pend = None
    run = lambda units, tl: _run_group(io, units, list(tl), pts, t, w, n, xi_d, poses_d, scan_start_time, scan_end_time, atlas_map,
                                       scan_seq, base_batch, scfg, acfg, m_tile_view, eps_lift, eps_mass, recency_min_scale)
    groups, gens = [], {}
    rest = list(range(H))
    if update_map:
        # hypothesis 0 (pipeline.py:835-877, 998-1010, 1233-1447): evidence against the view of the inflated map -- gathered
        # functionally, bit-identical to inflating first -- then the inflation in place and the map update, all enqueued
        # behind each other without a host synchronisation; the other hypotheses then see the updated map
        active0 = PR.ma_hex_stencil_tile_ids(poses_h[0, :3], acfg.h_tile, acfg.r_stencil_tiles_xy, acfg.r_stencil_tiles_z)
        g0 = run([0], active0)
        groups.append(g0)
        # the in-place inflation and the map update go to the context's side stream (behind hypothesis 0's evidence): they
        # overlap the deskew + surfel extraction of the remaining hypotheses, whose view is prepared on the same side stream
        # behind the update
        route = H > 1 and not _NO_SIDE_ROUTE
        gens["inflate"] = PR._recency_inflate_gen(atlas_map, active0, scan_seq, acfg.recency_decay_lambda, recency_min_scale, chart_id)
        io.ctx.side_route(route)
        try:
            io_i, stats_i, _ = next(gens["inflate"])
        finally:
            io.ctx.side_route(False)
        out_holder = []
        b0 = _batch_unit(g0.batch, 0)
        a0 = PrimitiveAssociationResult(**{f: getattr(g0.association, f)[0] for f in g0.association.__dataclass_fields__})
        if z_t is None:
            z_t = poses_h[0] if z_lin_poses is None else (lin_h[0] if lin_h is not None else poses_d[0])
        ctr = getattr(atlas_map, "_next_id_dev", None)
        if ctr is None:
            ctr = atlas_map._next_id_dev = torch.zeros(1, dtype=torch.int64, device=io.dev)
        if pend is None:          # the host attribute is current: put it on the device (stream-ordered, from pinned memory)
            stage = _PinnedRing.get(8)
            owned.append(stage)
            stage.view(torch.int64)[0] = int(atlas_map.next_global_id)
            ctr.copy_(stage.view(torch.int64), non_blocking=True)
        gens["update"] = PR._map_update_step12b_gen(atlas_map, b0, a0, active0, z_t, scan_seq, scan_end_time,
                                                    inflate_stats=lambda: out_holder[0]._inflate_stats, next_id_dev=ctr,
                                                    **(map_update_kwargs or {}))
        io.ctx.side_route(route)
        try:
            io_u, stats_u, _ = next(gens["update"])
        finally:
            io.ctx.side_route(False)
        rest = rest[1:]
    by_stencil: Dict[tuple, List[int]] = {}
    if rest:
        cells = PR.ma_hex_cells_3d_from_xyz_batch(poses_h[rest, :3], acfg.h_tile)
        for hh, cell in zip(rest, cells):
            by_stencil.setdefault(PR.stencil_of_cell(cell, acfg.r_stencil_tiles_xy, acfg.r_stencil_tiles_z), []).append(hh)
    groups += [run(units, tl) for tl, units in by_stencil.items()]
    cfg = dict(dev=io.dev, chart_id=chart_id, anchor_id=anchor_id, ess_imu=ess_imu, surfel=scfg, assoc=acfg, eps_lift=eps_lift,
               atlas=atlas_map, timestamps=t)
    out = BatchedPrimitiveEvidence(H, groups, cfg)
    # ONE synchronisation for the certificates of every hypothesis: the packed scalars go to pinned memory behind the
    # kernels; `defer` leaves the wait to the caller (out.wait()), so that the next scan can be enqueued meanwhile
    out._pending = []
    out._owned = owned
    if update_map:
        io.ctx.side_join(io.stream())      # the update's outputs (and the map) are read from the caller's stream from here on
    for g in groups:
        b1 = _PinnedRing.get(g._scal_d.numel())
        owned.append(b1)
        b1.copy_(g._scal_d, non_blocking=True)
        out._pending.append(b1)
    if update_map:
        out_holder.append(out)
        for name, st_d in (("inflate", stats_i), ("update", stats_u)):
            buf = _PinnedRing.get(st_d.numel() * 8)
            owned.append(buf)
            buf.view(F64).copy_(st_d.reshape(-1), non_blocking=True)
            out._gens[name] = (gens[name], buf, tuple(st_d.shape))
        atlas_map._pending_update = out
        out._ids_on_device = True
    out._event = torch.cuda.Event()
    out._event.record(torch.cuda.current_stream(io.dev))
    if not defer:
        out.wait()
    return out
