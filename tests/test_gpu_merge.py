"""
primitive_map_merge_reduce on the device (merge half) against the reference's own outputs
(tests/golden/merge_*.npz, made by tests/golden/make_golden_merge.py -- including the tile of the reference's
known-answer test, test/test_primitive_map_merge_reduce.py:77-99) and against oracle/merge.py at the reference's size
cap (2048 slots).  Through the C-ABI entry gcs_map_merge_reduce.  Selected pairs, validity, ids, masses and stamps are
bit-exact; merged moments 1e-9.
"""
import glob
import os

import numpy as np
import pytest

from conftest import GOLDEN, golden, rel_err
from test_oracle_merge_vs_golden import EXACT, FIELDS, kwargs_of, tile_of

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")

MERGE_CASES = sorted(os.path.basename(p) for p in glob.glob(os.path.join(GOLDEN, "merge_*.npz")))


@pytest.fixture(scope="module")
def P():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    from gc_slam_b200 import primitives
    return primitives


def _atlas_from_tile(P, t):
    m = int(np.asarray(t["weights"]).shape[0])
    td = dict(t)
    atl = dict(tiles={int(t["tile_id"]): td}, next_global_id=10 ** 6, total_count=int(t["count"]), m_tile=m)
    return P.AtlasMap.from_numpy(atl)


@pytest.mark.parametrize("case", MERGE_CASES)
def test_merge_vs_reference_golden(P, case):
    g = golden(case)
    t = tile_of(g)
    amap = _atlas_from_tile(P, t)
    res, cert, eff = P.primitive_map_merge_reduce(amap, int(g["tile_id"]), **kwargs_of(g))
    assert res.n_merged == int(g["n_merged"]) and amap.total_count == int(g["total_count"])
    assert cert.exact == bool(g["exact"]) and cert.approximation_triggers == [str(x) for x in g["triggers"]]
    assert cert.frobenius_applied == bool(g["frobenius_applied"])
    assert abs(cert.influence.mass_epsilon_ratio - float(g["mass_epsilon_ratio"])) < 1e-15
    assert eff.predicted == float(g["predicted"]) and eff.realized == float(g["realized"])
    out = amap.download_tile(int(g["tile_id"]))
    for k in EXACT:
        assert np.array_equal(out[k], g["out_" + k]), k
    for k in ("Lambdas", "thetas", "etas", "colors", "rgb"):
        assert rel_err(out[k], g["out_" + k]) < 1e-9, k


def test_merge_at_the_size_cap_vs_oracle_and_noops(P):
    from gc_slam_b200 import synth
    from oracle import merge as om
    atl = synth.synthetic_atlas(120000, 2048, 71, scan_seq=30)
    tid = max(atl["tiles"], key=lambda t: atl["tiles"][t]["count"])
    t_in = {k: (np.array(v, copy=True) if isinstance(v, np.ndarray) else v) for k, v in atl["tiles"][tid].items()}
    amap = P.AtlasMap.from_numpy(atl)
    before = amap.total_count
    res, cert, eff = P.primitive_map_merge_reduce(amap, tid, merge_threshold=0.3, max_pairs=32)
    t_ref, n_ref, _ = om.merge_reduce_tile(t_in, merge_threshold=0.3, max_pairs=32)
    assert res.n_merged == n_ref > 0 and amap.total_count == before - n_ref
    out = amap.download_tile(tid)
    for k in EXACT:
        assert np.array_equal(out[k], t_ref[k]), k
    assert rel_err(out["Lambdas"], t_ref["Lambdas"]) < 1e-9 and rel_err(out["thetas"], t_ref["thetas"]) < 1e-9
    # a second call merges the next-closest pairs; missing tile and zero budget are exact no-ops
    res2, _, _ = P.primitive_map_merge_reduce(amap, tid, merge_threshold=0.3, max_pairs=32)
    t_ref2, n_ref2, _ = om.merge_reduce_tile(t_ref, merge_threshold=0.3, max_pairs=32)
    assert res2.n_merged == n_ref2
    r3, c3, e3 = P.primitive_map_merge_reduce(amap, 123456789)
    assert r3.n_merged == 0 and c3.exact and e3.predicted == 0.0
    r4, c4, _ = P.primitive_map_merge_reduce(amap, tid, max_pairs=0)
    assert r4.n_merged == 0 and c4.exact
