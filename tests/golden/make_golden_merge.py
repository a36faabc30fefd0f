#!/usr/bin/env python
"""
Golden vectors for primitive_map_merge_reduce, produced by the REFERENCE's own operator
(fl/backend/structures/primitive_map.py:1501-2031) on top of oracle/jax_shim, including the tile of the reference's
known-answer test (test/test_primitive_map_merge_reduce.py:12-99).

Run in the build container only (needs /root/reference):   python tests/golden/make_golden_merge.py
Outputs tests/golden/merge_*.npz.  Nothing under /root/reference is written or copied.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path.insert(0, os.path.join(ROOT, "oracle", "jax_shim"))
sys.path.insert(0, os.path.join(REF, "fl_ws", "src", "fl_slam_poc"))
sys.path.insert(0, ROOT)

FIELDS = ("Lambdas", "thetas", "etas", "weights", "timestamps", "created_timestamps", "last_supported_scan_seq",
          "last_update_scan_seq", "primitive_ids", "valid_mask", "colors", "cam_mass", "lidar_mass", "rgb_cam_accum",
          "rgb_cam_denom", "rgb")


def main():
    import jax.numpy as jnp
    from fl_slam_poc.backend.structures import primitive_map as pm

    from gc_slam_b200 import synth

    def run(td, name, **kw):
        tile = pm.PrimitiveMapTile(**{k: (jnp.asarray(v) if isinstance(v, np.ndarray) else v) for k, v in td.items()})
        atlas = pm.AtlasMap(tiles={int(td["tile_id"]): tile}, next_global_id=10 ** 6, total_count=int(td["count"]),
                            m_tile=int(np.asarray(td["weights"]).shape[0]))
        res, cert, eff = pm.primitive_map_merge_reduce(atlas_map=atlas, tile_id=int(td["tile_id"]), **kw)
        nt = res.atlas_map.tiles[int(td["tile_id"])]
        out = {"in_" + k: np.asarray(td[k]) for k in FIELDS}
        out.update({"out_" + k: np.asarray(getattr(nt, k)) for k in FIELDS})
        out.update(n_merged=res.n_merged, total_count=res.atlas_map.total_count, exact=cert.exact,
                   triggers=np.array(cert.approximation_triggers, dtype="U64"), frobenius_applied=cert.frobenius_applied,
                   mass_epsilon_ratio=cert.influence.mass_epsilon_ratio, predicted=eff.predicted, realized=eff.realized,
                   tile_id=int(td["tile_id"]), count_in=int(td["count"]), next_local_id=int(td["next_local_id"]),
                   **{"kw_" + k: v for k, v in kw.items()})
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
        print(name, "n_merged", res.n_merged, "triggers", cert.approximation_triggers)

    # 1: the tile of the reference's own known-answer test
    m, b = 3, 3
    mu = np.array([[0.0, 0.0, 0.0], [0.01, 0.0, 0.0], [10.0, 0.0, 0.0]])
    ref_tile = dict(tile_id=0, Lambdas=np.stack([np.eye(3)] * 3), thetas=mu.copy(), etas=np.zeros((m, b, 3)), weights=np.ones(m),
                    timestamps=np.zeros(m), created_timestamps=np.zeros(m), last_supported_scan_seq=np.zeros(m, dtype=np.int64),
                    last_update_scan_seq=np.zeros(m, dtype=np.int64), primitive_ids=np.arange(3, dtype=np.int64),
                    valid_mask=np.ones(m, dtype=bool), colors=np.zeros((m, 3)), cam_mass=np.array([1.0, 0.0, 0.0]),
                    lidar_mass=np.array([0.0, 1.0, 1.0]), rgb_cam_accum=np.array([[1.0, 0, 0], [0, 0, 0], [0, 0, 0]]),
                    rgb_cam_denom=np.array([1.0, 0.0, 0.0]), rgb=np.array([[1.0, 0, 0], [0.5, 0.5, 0.5], [0.5, 0.5, 0.5]]),
                    next_local_id=3, count=3)
    run(ref_tile, "merge_reference_test_tile", merge_threshold=0.5, max_pairs=1, max_tile_size=10)

    # 2-3: dense synthetic tiles (surfels on the room's surfaces crowd the slots of one tile), reference budgets
    for name, n_surf, m_tile, seed, kw in (("merge_dense_160", 4000, 160, 61, dict(merge_threshold=0.1, max_pairs=4, max_tile_size=2048)),
                                           ("merge_dense_96_pairs16", 3000, 96, 62, dict(merge_threshold=0.6, max_pairs=16, max_tile_size=2048))):
        atl = synth.synthetic_atlas(n_surf, m_tile, seed, scan_seq=30)
        tid = max(atl["tiles"], key=lambda t: atl["tiles"][t]["count"])
        run(atl["tiles"][tid], name, **kw)
    # 4: budget cap (tile larger than max_tile_size): approximate no-op certificate
    atl = synth.synthetic_atlas(500, 64, 63, scan_seq=30)
    tid = max(atl["tiles"], key=lambda t: atl["tiles"][t]["count"])
    run(atl["tiles"][tid], "merge_budget_cap", merge_threshold=0.1, max_pairs=4, max_tile_size=32)


if __name__ == "__main__":
    main()
