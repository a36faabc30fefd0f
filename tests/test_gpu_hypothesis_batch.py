"""
GPU parity of the hypothesis-batched primitive path (pytest -m gpu): H hypotheses of one scan through
lidar_evidence_primitives_batched (unit axis in every kernel's grid, one shared read-only view per stencil, functional
recency inflation) against

  * the single-hypothesis operators called hypothesis by hypothesis in the reference's order
    (fl/backend/backend_node.py:2036-2083: hypothesis 0 updates the map, the others see the updated map and inflate a
    copy of it) -- bit for bit, and
  * the NumPy oracle of the same loop (oracle/prim_path.py), which is pinned to the reference-generated goldens.
"""
import numpy as np
import pytest

from conftest import rel_err

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")


def _np(t):
    return t.detach().cpu().numpy()


@pytest.fixture(scope="module")
def mods():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from gc_slam_b200 import hypothesis_batch, operators, primitives
    return primitives, operators, hypothesis_batch


def _inputs(n, H, seed, spread=False):
    from gc_slam_b200 import synth
    pts, t, w, _, _ = synth.vlp16_scan(n, seed, t0=synth.EPOCH_T0)
    xis = np.stack([synth.scan_twist(seed + 7 * h) for h in range(H)])
    poses = synth.hypothesis_poses(H, seed + 1)
    poses[:, :3] += np.array([0.1, -0.2, 0.0])
    if spread:                      # two of the hypotheses predict a position in the neighbouring map tile
        poses[2, 0] += 2.1
        poses[H - 1, 0] += 2.1
    cam = synth.camera_splats(150, seed + 2)
    return pts, t, w, xis, poses, cam


def _base(P, cam):
    return P.measurement_batch_from_camera_splats(cam["positions"], cam["covariances"], cam["directions"], cam["kappas"],
                                                  cam["weights"], cam["timestamps"], cam["colors"])


def _clone_map(P, amap):
    c = P.AtlasMap(m_tile=amap.m_tile, n_tiles_cap=amap.n_tiles_cap, device=amap.device)
    for k, v in amap.fields.items():
        c.fields[k].copy_(v)
    c.tiles = dict(amap.tiles)
    c.next_global_id, c.total_count = amap.next_global_id, amap.total_count
    return c


def _one_by_one(P, ops, amap, pts, t, w, xi, pose, cam, scan_seq, t0, t1):
    """One hypothesis with the stand-alone operators on a COPY of the map (the reference's functional inflation)."""
    m = _clone_map(P, amap)
    active = P.ma_hex_stencil_tile_ids(pose[:3])
    dk, c_dk, _ = ops.deskew_constant_twist(pts, t, w, t0, t1, xi, 1.0, "GC-RIGHT-01", "a")
    batch, c_sf, _ = P.extract_lidar_surfels(dk.points, dk.timestamps, dk.weights, None, _base(P, cam))
    m, _, _, inf = P.primitive_map_recency_inflate(m, active, scan_seq)
    view = P.extract_atlas_map_view(m, active, 1024)
    assoc, c_as, e_as = P.associate_primitives_ot(batch, view, P.AssociationConfig(scan_seq=scan_seq))
    vpe, c_pe, _ = P.visual_pose_evidence(assoc, batch, view, pose, z_lin_pose=pose)
    return dict(dk=dk, c_dk=c_dk, batch=batch, c_sf=c_sf, inf=inf, view=view, assoc=assoc, c_as=c_as, e_as=e_as, vpe=vpe, c_pe=c_pe)


@pytest.mark.parametrize("update_map", [False, True])
def test_batched_hypotheses_bit_identical_to_the_loop(mods, update_map):
    P, ops, HB = mods
    from gc_slam_b200 import synth
    n, H, scan_seq = 30000, 6, 23
    pts, t, w, xis, poses, cam = _inputs(n, H, 51, spread=True)
    t0, t1 = synth.EPOCH_T0, synth.EPOCH_T0 + 0.1
    atlas_np = synth.synthetic_atlas(150000, 50000, 10, scan_seq=20)
    amap = P.AtlasMap.from_numpy(atlas_np, n_tiles_cap=len(atlas_np["tiles"]) + 24)
    amap_ref = _clone_map(P, amap)

    out = HB.lidar_evidence_primitives_batched(pts, t, w, t0, t1, xis, amap, poses, scan_seq, base_batch=_base(P, cam),
                                               update_map=update_map)
    assert out.L_pose.shape == (H, 22, 22) and out.h_pose.shape == (H, 22)
    n_groups = len({tuple(P.ma_hex_stencil_tile_ids(p[:3])) for p in poses[(1 if update_map else 0):]})
    assert len(out.groups) == n_groups + (1 if update_map else 0) and n_groups >= 2   # hypothesis 0 runs ahead of the map update

    # the reference's order with the stand-alone operators
    h_first = 0
    if update_map:
        active0 = P.ma_hex_stencil_tile_ids(poses[0, :3])
        f = P.lidar_evidence_primitives(pts, t, w, t0, t1, xis[0], amap_ref, active0, poses[0], scan_seq, base_batch=_base(P, cam))
        assert torch.equal(out.L_pose[0], f["pose_evidence"][0].L_pose) and torch.equal(out.h_pose[0], f["pose_evidence"][0].h_pose)
        assert out.map_update[0].n_inserted == f["map_update"][0].n_inserted
        u0 = out.unit(0)
        assert torch.equal(u0["association"][0].responsibilities, f["association"][0].responsibilities)
        assert torch.equal(u0["surfels"][0].Lambdas, f["surfels"][0].Lambdas) and u0["surfels"][0].n_lidar_valid == f["surfels"][0].n_lidar_valid
        assert torch.equal(u0["map_view"].positions, f["map_view"].positions) and u0["map_view"].n_valid == f["map_view"].n_valid
        assert u0["recency_inflate"][3].stale_precision_downscale_total == f["recency_inflate"][3].stale_precision_downscale_total
        mu_a, mu_b = u0["map_update"][1].map_update, f["map_update"][1].map_update
        assert (mu_a.fused_count, mu_a.insert_count_total, mu_a.evicted_count) == (mu_b.fused_count, mu_b.insert_count_total, mu_b.evicted_count)
        assert mu_a.fused_mass_total == mu_b.fused_mass_total and mu_a.staleness_inflation_strength == mu_b.staleness_inflation_strength
        assert torch.equal(u0["map_update"][0].new_ids, f["map_update"][0].new_ids)
        assert amap.next_global_id == amap_ref.next_global_id and amap.total_count == amap_ref.total_count
        for name in amap.fields:                       # the batch left the updated map untouched
            assert torch.equal(amap.fields[name], amap_ref.fields[name]), name
        h_first = 1
    for h in range(h_first, H):
        ref = _one_by_one(P, ops, amap_ref, pts, t, w, xis[h], poses[h], cam, scan_seq, t0, t1)
        u = out.unit(h)
        dk, c_dk, _ = u["deskew"]
        assert torch.equal(dk.points, ref["dk"].points) and torch.equal(dk.weights, ref["dk"].weights)
        assert c_dk.support.support_frac == ref["c_dk"].support.support_frac
        b, c_sf, _ = u["surfels"]
        assert b.n_lidar_valid == ref["batch"].n_lidar_valid and c_sf.support.ess_total == ref["c_sf"].support.ess_total
        for fld in ("Lambdas", "thetas", "etas", "weights", "timestamps", "colors", "sources", "source_indices", "valid_mask"):
            assert torch.equal(getattr(b, fld), getattr(ref["batch"], fld)), fld
        inf = u["recency_inflate"][3]
        assert inf.stale_precision_downscale_total == ref["inf"].stale_precision_downscale_total
        assert inf.staleness_cov_inflation_trace == ref["inf"].staleness_cov_inflation_trace
        v = u["map_view"]
        assert v.n_valid == ref["view"].n_valid and v.tile_ids == ref["view"].tile_ids
        for fld in ("candidate_slots", "candidate_tile_ids", "valid_mask", "positions", "covariances", "directions", "kappas",
                    "weights", "primitive_ids", "last_supported_scan_seq", "etas", "colors"):
            assert torch.equal(getattr(v, fld), getattr(ref["view"], fld)), fld
        a, c_as, e_as = u["association"]
        for fld in ("candidate_pool_indices", "candidate_tile_ids", "candidate_slots", "cost_matrix", "responsibilities", "row_masses"):
            assert torch.equal(getattr(a, fld), getattr(ref["assoc"], fld)), fld
        assert c_as.ot.transport_mass_total == ref["c_as"].ot.transport_mass_total and c_as.ot.b_recency_p95 == ref["c_as"].ot.b_recency_p95
        assert e_as.predicted == ref["e_as"].predicted
        vpe, c_pe, _ = u["pose_evidence"]
        assert torch.equal(vpe.L_pose, ref["vpe"].L_pose) and torch.equal(vpe.h_pose, ref["vpe"].h_pose)
        assert torch.equal(out.L_pose[h], ref["vpe"].L_pose) and torch.equal(out.h_pose[h], ref["vpe"].h_pose)
        assert vpe.total_weighted_cost == ref["vpe"].total_weighted_cost and vpe.n_associations == ref["vpe"].n_associations
        assert c_pe.support.ess_total == ref["c_pe"].support.ess_total
    for name in amap.fields:                           # read-only hypotheses never write the map
        assert torch.equal(amap.fields[name], amap_ref.fields[name]), name


def test_batched_hypotheses_vs_oracle(mods):
    """Reference budgets, 65,536 points, 4 hypotheses: every unit against the NumPy oracle of the same steps."""
    P, ops, HB = mods
    from gc_slam_b200 import synth
    from oracle import bin_path as ob
    from oracle import prim_path as op
    n, H, scan_seq = 65536, 4, 21
    pts, t, w, xis, poses, cam = _inputs(n, H, 61)
    t0, t1 = synth.EPOCH_T0, synth.EPOCH_T0 + 0.1
    atlas_np = synth.synthetic_atlas(400000, 50000, 9, scan_seq=20)
    amap = P.AtlasMap.from_numpy(atlas_np)
    out = HB.lidar_evidence_primitives_batched(pts, t, w, t0, t1, xis, amap, poses, scan_seq, base_batch=_base(P, cam), update_map=False)
    ob_base = op.batch_from_camera_splats(cam["positions"], cam["covariances"], cam["directions"], cam["kappas"],
                                          cam["weights"], cam["timestamps"], cam["colors"])
    for h in range(H):
        dk, _ = ob.deskew_constant_twist(pts, t, w, t0, t1, xis[h])
        o_batch, _, _ = op.extract_lidar_surfels(dk["points"], t, dk["weights"], ob_base)
        active = op.stencil_tile_ids(poses[h, :3])
        o_atlas, _ = op.recency_inflate(atlas_np, active, scan_seq)
        o_view = op.extract_atlas_map_view(o_atlas, active)
        o_assoc, _ = op.associate_primitives_ot(o_batch, o_view, scan_seq=scan_seq)
        o_vpe, _ = op.visual_pose_evidence(o_assoc, o_batch, o_view, poses[h])
        u = out.unit(h)
        b = u["surfels"][0]
        assert b.n_lidar_valid == o_batch["n_lidar_valid"]
        assert rel_err(_np(b.Lambdas), o_batch["Lambdas"]) < 1e-7 and rel_err(_np(b.etas), o_batch["etas"]) < 1e-8
        v = u["map_view"]
        assert np.array_equal(_np(v.candidate_slots), o_view["candidate_slots"])
        vm = _np(v.valid_mask).astype(bool)
        assert rel_err(_np(v.positions) * vm[:, None], o_view["positions"] * vm[:, None]) < 1e-8
        a = u["association"][0]
        bad = int(np.sum(np.any(_np(a.candidate_pool_indices) != o_assoc["candidate_pool_indices"], axis=1)))
        assert bad == 0, f"hypothesis {h}: {bad} rows differ in their candidate set"
        assert rel_err(_np(a.responsibilities), o_assoc["responsibilities"]) < 1e-7
        assert rel_err(_np(out.L_pose[h]), o_vpe["L_pose"]) < 1e-7 and rel_err(_np(out.h_pose[h])[:3], o_vpe["h_pose"][:3]) < 1e-6
    # a second run is bit-identical
    out2 = HB.lidar_evidence_primitives_batched(pts, t, w, t0, t1, xis, amap, poses, scan_seq, base_batch=_base(P, cam), update_map=False)
    assert torch.equal(out.L_pose, out2.L_pose) and torch.equal(out.h_pose, out2.h_pose)


def test_batched_empty_map_and_errors(mods):
    P, ops, HB = mods
    from gc_slam_b200 import synth
    pts, t, w, xis, poses, cam = _inputs(4096, 3, 71)
    amap = P.create_empty_atlas_map(m_tile=2048, n_tiles_cap=16)
    cfg = P.SurfelExtractionConfig(n_surfel=128, n_feat=16)
    out = HB.lidar_evidence_primitives_batched(pts, t, w, 0.0, 0.1, xis, amap, poses, 1, surfel_config=cfg, m_tile_view=64, update_map=False)
    for h in range(3):
        u = out.unit(h)
        assert u["map_view"].n_valid == 0
        a, c, e = u["association"]
        assert c.exact and float(a.responsibilities.abs().sum().item()) == 0.0 and e.predicted == 0.0
        assert u["pose_evidence"][1].exact and float(u["pose_evidence"][0].h_pose.abs().sum().item()) == 0.0
    # first scan on an empty map with the update: hypothesis 0 inserts, the others then see a map that is no longer empty
    out = HB.lidar_evidence_primitives_batched(pts, t, w, 0.0, 0.1, xis, amap, poses, 1, surfel_config=cfg, m_tile_view=64, update_map=True,
                                               map_update_kwargs=dict(k_insert_tile=16))
    ref_map = P.create_empty_atlas_map(m_tile=2048, n_tiles_cap=16)
    f = P.lidar_evidence_primitives(pts, t, w, 0.0, 0.1, xis[0], ref_map, P.ma_hex_stencil_tile_ids(poses[0, :3]), poses[0], 1,
                                    surfel_config=cfg, m_tile_view=64, map_update_kwargs=dict(k_insert_tile=16))
    assert out.unit(0)["association"][1].exact and out.map_update[0].n_inserted == f["map_update"][0].n_inserted > 0
    assert torch.equal(out.L_pose[0], f["pose_evidence"][0].L_pose) and torch.equal(out.unit(0)["map_update"][0].new_ids, f["map_update"][0].new_ids)
    for name in amap.fields:
        assert torch.equal(amap.fields[name][:len(amap.tiles)], ref_map.fields[name][:len(ref_map.tiles)]), name
    v1 = P.extract_atlas_map_view(amap, P.ma_hex_stencil_tile_ids(poses[1, :3]), 64)      # the map after hypothesis 0's update
    assert out.unit(1)["map_view"].n_valid == v1.n_valid and torch.equal(out.unit(1)["map_view"].valid_mask, v1.valid_mask)
    with pytest.raises(ValueError):
        HB.lidar_evidence_primitives_batched(pts, t, w, 0.0, 0.1, xis, amap, poses[:2], 1, surfel_config=cfg, m_tile_view=64)
    with pytest.raises(ValueError):
        HB.lidar_evidence_primitives_batched(pts[:10], t, w, 0.0, 0.1, xis, amap, poses, 1, surfel_config=cfg, m_tile_view=64)


def test_workspace_is_sized_once_and_can_be_frozen(mods):
    """no per-call allocation in the steady state.  The context starts with 32 MB; a batch that needs more grows
    it once (the outgrown block is retired, not freed under a live stream); frozen, a call that would have to grow raises
    instead of allocating, and after an explicit reserve the same call runs."""
    P, ops, HB = mods
    from gc_slam_b200 import _lib as L
    ctx = L.Context(torch.cuda.current_device())
    try:
        assert ctx.workspace_bytes == 32 << 20
        ctx.freeze_workspace(True)
        with pytest.raises(RuntimeError, match="workspace frozen"):
            ctx.check(_grow_inside_a_call(ctx, L))
        ctx.reserve_workspace(96 << 20)          # explicit sizing is allowed while frozen (it is made outside the steady state)
        assert ctx.workspace_bytes >= 96 << 20
        assert _grow_inside_a_call(ctx, L) == 0
    finally:
        ctx.close()


def _grow_inside_a_call(ctx, L):
    """A deskew of 2^21 points x 8 units needs 8 x blocks x 16 B of partials -- tiny; the surfel batch needs ~64 MB."""
    import ctypes as C
    from gc_slam_b200 import hypothesis_batch as HB, primitives as PR
    n, U = 65536, 16
    z = lambda *s, dt=torch.float64: torch.zeros(*s, dtype=dt, device="cuda")
    pts, t, w = z(U, n, 3), z(n), z(U, n)
    b = PR.create_empty_measurement_batch()
    stacked = PR.MeasurementBatch(**{f: (getattr(b, f).unsqueeze(0).repeat((U,) + (1,) * getattr(b, f).dim()) if isinstance(getattr(b, f), torch.Tensor)
                                         else getattr(b, f)) for f in b.__dataclass_fields__})
    cb, cc = stacked._c(), PR.SurfelExtractionConfig()._c()
    nv = z(U, dt=torch.int32)
    return ctx.lib.gcs_extract_lidar_surfels_batched(ctx.handle, L.stream_ptr(torch.device("cuda")), L.ptr(pts), L.ptr(t), L.ptr(w), n, U, 1,
                                                     C.byref(cc), C.byref(cb), L.ptr(nv))


def test_deferred_scans_with_map_update_equal_the_synchronous_loop(mods):
    """Scans with the map update enqueued back to back (the wait for scan k-1 issued after scan k: the id counter lives on
    the device) leave the map, the ids and the evidence of the loop that waits after every scan."""
    P, ops, HB = mods
    from gc_slam_b200 import synth
    n, H = 20000, 3
    atlas_np = synth.synthetic_atlas(120000, 50000, 12, scan_seq=20)
    t0, t1 = synth.EPOCH_T0, synth.EPOCH_T0 + 0.1
    scans = []
    for k in range(5):
        pts, t, w, xis, poses, cam = _inputs(n, H, 70 + k)
        scans.append((pts, t, w, xis, poses, cam))

    def run(deferred):
        amap = P.AtlasMap.from_numpy(atlas_np, n_tiles_cap=len(atlas_np["tiles"]) + 24)
        outs, prev = [], None
        for k, (pts, t, w, xis, poses, cam) in enumerate(scans):
            cur = HB.lidar_evidence_primitives_batched(pts, t, w, t0, t1, xis, amap, poses, 21 + k, base_batch=_base(P, cam),
                                                       update_map=True, defer=deferred)
            if deferred and prev is not None:
                prev.wait()
            prev = cur
            outs.append(cur)
        prev.wait()
        return amap, outs

    m_sync, o_sync = run(False)
    m_def, o_def = run(True)
    assert m_def.next_global_id == m_sync.next_global_id and m_def.total_count == m_sync.total_count
    assert m_def.next_global_id > int(atlas_np["next_global_id"])
    for name in m_sync.fields:
        assert torch.equal(m_def.fields[name], m_sync.fields[name]), name
    for a, b in zip(o_def, o_sync):
        assert torch.equal(a.L_pose, b.L_pose) and torch.equal(a.h_pose, b.h_pose)
        ra, rb = a.map_update[0], b.map_update[0]
        assert torch.equal(ra.new_ids, rb.new_ids) and (ra.n_fused, ra.n_inserted, ra.n_culled) == (rb.n_fused, rb.n_inserted, rb.n_culled)
        ca, cb = a.map_update[1].map_update, b.map_update[1].map_update
        assert ca.tile_ids_inactive == cb.tile_ids_inactive and ca.insert_mass_total == cb.insert_mass_total
