"""
The per-tile map operators on the device (primitive_map_fuse / insert_masked / cull / forget keep the
reference's signatures; block_associations_for_fuse) against the reference's own outputs (tests/golden/mapops_*.npz, made
by tests/golden/make_golden_mapops.py) and against oracle/prim_path.py at the production tile size (50,000 slots).
Through the C-ABI entries gcs_map_fuse / gcs_map_insert_masked / gcs_map_cull / gcs_map_forget.  Slots, ids, validity and
stamps are bit-exact; fused moments 1e-12 (the segmented sums add in a different, fixed order).
"""
import numpy as np
import pytest

from conftest import golden, rel_err
from test_oracle_mapops_vs_golden import CULL_CASES, EXACT, FLOAT, FUSE_CASES, INSERT_CASES, atlas_of, fuse_args

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")


@pytest.fixture(scope="module")
def P():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    from gc_slam_b200 import primitives
    return primitives


def check_tile(out, ref, tol, prefix="out_"):
    for k in EXACT:
        assert np.array_equal(out[k], ref[prefix + k]), k
    for k in FLOAT:
        assert rel_err(out[k], ref[prefix + k]) < tol, k


@pytest.mark.parametrize("case", FUSE_CASES)
def test_fuse_vs_reference_golden(P, case):
    g = golden(case)
    a = fuse_args(g)
    amap = P.AtlasMap.from_numpy(atlas_of(g))
    kw = dict(valid_mask=a["vm"], colors_meas=a["col"], sources_meas=a["src"]) if bool(g["full"]) else {}
    res, cert, eff = P.primitive_map_fuse(amap, int(g["tile_id"]), a["slots"], a["lam"], a["th"], a["eta"], a["w"], a["resp"],
                                          float(g["timestamp"]), int(g["scan_seq"]), **kw)
    assert res.n_fused == int(g["n_fused"]) and cert.exact
    assert eff.predicted == float(g["predicted"]) and eff.realized == float(g["realized"])
    check_tile(amap.download_tile(int(g["tile_id"])), g, 1e-12)


@pytest.mark.parametrize("case", INSERT_CASES)
def test_insert_masked_vs_reference_golden(P, case):
    g = golden(case)
    full = bool(g["full"])
    amap = P.AtlasMap.from_numpy(atlas_of(g, int(g["next_global_id_in"])))
    kw = dict(colors_new=g["col"], sources_new=g["src"]) if full else {}
    res, cert, eff = P.primitive_map_insert_masked(amap, int(g["tile_id"]), g["lam"], g["th"], g["eta"], g["w"], float(g["timestamp"]),
                                                   g["vnew"], scan_seq=int(g["scan_seq"]), **kw)
    assert res.n_inserted == int(g["n_inserted"]) and np.array_equal(res.new_ids.cpu().numpy(), g["new_ids"])
    assert amap.next_global_id == int(g["next_global_id"]) and amap.total_count == int(g["total_count"])
    assert cert.exact == bool(g["exact"]) and cert.approximation_triggers == [str(x) for x in g["triggers"]]
    assert eff.predicted == float(g["predicted"]) and eff.realized == float(g["realized"])
    out = amap.download_tile(int(g["tile_id"]))
    assert out["count"] == int(g["count_out"])
    check_tile(out, g, 1e-15)


@pytest.mark.parametrize("case", CULL_CASES)
def test_cull_vs_reference_golden(P, case):
    g = golden(case)
    amap = P.AtlasMap.from_numpy(atlas_of(g))
    maxp = None if int(g["maxp"]) < 0 else int(g["maxp"])
    res, cert, eff = P.primitive_map_cull(amap, int(g["tile_id"]), weight_threshold=float(g["thr"]), max_primitives=maxp)
    assert res.n_culled == int(g["n_culled"]) and abs(res.mass_dropped - float(g["mass_dropped"])) <= 1e-13 * max(1.0, res.mass_dropped)
    assert amap.total_count == int(g["total_count"])
    assert cert.exact == bool(g["exact"]) and cert.approximation_triggers == [str(x) for x in g["triggers"]]
    assert abs(cert.influence.mass_epsilon_ratio - float(g["mass_epsilon_ratio"])) < 1e-13
    assert eff.predicted == float(g["predicted"]) and eff.realized == float(g["realized"])
    out = amap.download_tile(int(g["tile_id"]))
    assert out["count"] == int(g["count_out"])
    check_tile(out, g, 1e-15)


def test_forget_vs_reference_golden_and_missing_tile(P):
    g = golden("mapops_forget.npz")
    amap = P.AtlasMap.from_numpy(atlas_of(g))
    res, cert, eff = P.primitive_map_forget(amap, int(g["tile_id"]), forgetting_factor=float(g["gamma"]))
    assert cert.exact and eff.predicted == float(g["predicted"]) and eff.realized == float(g["realized"])
    check_tile(amap.download_tile(int(g["tile_id"])), g, 1e-16)
    # operators on a tile that does not exist: the reference's exact no-ops (:1202-1215, :1337-1346, :1031-1041)
    n_tiles = amap.n_tiles
    _, c1, e1 = P.primitive_map_forget(amap, 123456789)
    _, c2, e2 = P.primitive_map_cull(amap, 123456789)
    r3, c3, e3 = P.primitive_map_fuse(amap, 123456789, np.zeros(0, np.int32), np.zeros((0, 3, 3)), np.zeros((0, 3)), np.zeros((0, 3, 3)),
                                      np.zeros(0), np.zeros(0), 1.0)
    assert c1.exact and c2.exact and c3.exact and e1.realized == 0.0 and e2.realized == 0.0 and r3.n_fused == 0
    assert amap.n_tiles == n_tiles


def test_block_associations_vs_reference_golden(P):
    g = golden("mapops_block_assoc.npz")
    dev = torch.device("cuda", 0)
    n, k = g["responsibilities"].shape
    ar = P.PrimitiveAssociationResult(responsibilities=torch.from_numpy(g["responsibilities"]).to(dev),
                                      candidate_pool_indices=torch.zeros((n, k), dtype=torch.int32, device=dev),
                                      candidate_tile_ids=torch.from_numpy(g["candidate_tile_ids"]).to(dev),
                                      candidate_slots=torch.from_numpy(g["candidate_slots"]).to(dev),
                                      row_masses=torch.zeros(n, dtype=torch.float64, device=dev),
                                      cost_matrix=torch.zeros((n, k), dtype=torch.float64, device=dev))
    mi, ct, cs, rs, vr = P.block_associations_for_fuse(ar, g["valid_mask"], int(g["block"]))
    assert np.array_equal(mi.cpu().numpy(), g["out_meas_idx"]) and np.array_equal(ct.cpu().numpy(), g["out_tile_ids"])
    assert np.array_equal(cs.cpu().numpy(), g["out_slots"]) and np.array_equal(rs.cpu().numpy(), g["out_resp"])
    assert np.array_equal(vr.cpu().numpy(), g["out_valid_rows"])


def _proposals(n, seed):
    rng = np.random.default_rng(seed)
    a = rng.normal(size=(n, 3, 3))
    lam = np.einsum("nij,nkj->nik", a, a) + 0.5 * np.eye(3)[None]
    return (lam, rng.normal(size=(n, 3)) * 3.0, rng.normal(size=(n, 3, 3)), rng.random(n) + 0.05, rng.random((n, 3)) * 1.4 - 0.2,
            (rng.random(n) < 0.6).astype(np.int32))


def test_operator_sequence_at_production_tile_size_vs_oracle(P):
    """50,000-slot tiles (GC_PRIMITIVE_MAP_MAX_SIZE): fuse one 256 x 8 block, insert 64 with eviction, cull, forget -- the order
    of pipeline step 12b -- on the device and in the oracle, compared after every operator."""
    import copy

    from gc_slam_b200 import constants, synth
    from oracle import prim_path as op
    atl = synth.synthetic_atlas(120000, constants.GC_PRIMITIVE_MAP_MAX_SIZE, 81, scan_seq=30)
    tid = max(atl["tiles"], key=lambda t: atl["tiles"][t]["count"])
    one = dict(tiles={tid: atl["tiles"][tid]}, next_global_id=atl["next_global_id"], total_count=int(atl["tiles"][tid]["count"]),
               m_tile=atl["m_tile"])
    ref = copy.deepcopy(one)
    amap = P.AtlasMap.from_numpy(one)
    M = atl["m_tile"]
    rng = np.random.default_rng(82)
    n = 256 * 8
    lam, th, eta, w, col, src = _proposals(n, 83)
    valid_slots = np.flatnonzero(ref["tiles"][tid]["valid_mask"])
    slots = rng.choice(valid_slots[:400] if valid_slots.size >= 400 else np.arange(400), size=n).astype(np.int32)
    resp = rng.random(n) * (rng.random(n) < 0.8)
    vm = rng.random(n) < 0.7
    res, _, _ = P.primitive_map_fuse(amap, tid, slots, lam, th, eta, w, resp, 12.5, 31, valid_mask=vm, colors_meas=col, sources_meas=src)
    n_ref = op.primitive_map_fuse(ref, tid, slots, lam, th, eta, w, resp, 12.5, 31, vm, col, src)
    assert res.n_fused == n_ref
    check_tile(amap.download_tile(tid), ref["tiles"][tid], 1e-12, prefix="")

    k = constants.GC_K_INSERT_TILE
    lam, th, eta, w, col, src = _proposals(k, 84)
    vnew = rng.random(k) < 0.8
    res, _, _ = P.primitive_map_insert_masked(amap, tid, lam, th, eta, w, 12.5, vnew, scan_seq=31, colors_new=col, sources_new=src)
    n_ins, ids, tslots = op.primitive_map_insert_masked(ref, tid, lam, th, eta, w, 12.5, vnew, 31, constants.GC_RECENCY_DECAY_LAMBDA, col, src)
    assert res.n_inserted == n_ins and np.array_equal(res.new_ids.cpu().numpy(), ids)
    assert np.array_equal(res.target_slots.cpu().numpy(), tslots)
    assert amap.next_global_id == ref["next_global_id"] and amap.total_count == ref["total_count"]
    check_tile(amap.download_tile(tid), ref["tiles"][tid], 1e-12, prefix="")

    for thr, maxp in ((0.02, None), (0.0, 1500)):
        res, cert, _ = P.primitive_map_cull(amap, tid, weight_threshold=thr, max_primitives=maxp)
        n_c, mass = op.primitive_map_cull(ref, tid, thr, maxp)
        assert res.n_culled == n_c and n_c > 0 and abs(res.mass_dropped - mass) <= 1e-12 * mass
        assert amap.total_count == ref["total_count"] and amap.tile_count(tid) == ref["tiles"][tid]["count"]
        check_tile(amap.download_tile(tid), ref["tiles"][tid], 1e-12, prefix="")
    assert amap.tile_count(tid) == 1501   # strict "<" against the weight of descending rank 1500 keeps that primitive too

    P.primitive_map_forget(amap, tid)
    op.primitive_map_forget(ref, tid)
    check_tile(amap.download_tile(tid), ref["tiles"][tid], 1e-12, prefix="")


def test_contract_errors_are_value_errors(P):
    """Shape / budget violations fail loudly as ValueError (GCS_EINVAL), as the reference's fail-fast checks do; the map is
    left untouched."""
    g = golden("mapops_forget.npz")
    amap = P.AtlasMap.from_numpy(atlas_of(g))
    tid = int(g["tile_id"])
    before = amap.download_tile(tid)
    n = 16385                                              # one more than the per-call proposal budget of gcs_map_fuse
    with pytest.raises(ValueError):
        P.primitive_map_fuse(amap, tid, np.zeros(n, np.int32), np.zeros((n, 3, 3)), np.zeros((n, 3)), np.zeros((n, 3, 3)), np.ones(n),
                             np.ones(n), 1.0)
    k = amap.m_tile + 1                                    # more proposals than the tile has slots
    with pytest.raises(ValueError):
        P.primitive_map_insert_masked(amap, tid, np.zeros((k, 3, 3)), np.zeros((k, 3)), np.zeros((k, 3, 3)), np.ones(k), 1.0,
                                      np.ones(k, bool))
    with pytest.raises(ValueError):
        P.sinkhorn_unbalanced_fixed_k(np.zeros((4, 33)), np.ones(4) / 4, np.ones(33) / 33, 0.1, 0.5, 0.5, 5)
    with pytest.raises(ValueError):
        P.compute_sparse_cost_matrix(np.zeros((4, 3)), np.zeros((4, 3)), np.ones(4), np.zeros((2, 3)), np.zeros((2, 3)), np.ones(2),
                                     np.zeros((5, 8), np.int32))
    after = amap.download_tile(tid)
    for k_ in EXACT + FLOAT:
        assert np.array_equal(before[k_], after[k_]), k_
