#!/usr/bin/env python
"""
Golden vectors for the per-tile map operators, produced by the REFERENCE's own functions on top of oracle/jax_shim:
  primitive_map_fuse / primitive_map_insert_masked / primitive_map_cull / primitive_map_forget
      fl/backend/structures/primitive_map.py:807-1384
  block_associations_for_fuse
      fl/backend/operators/primitive_association.py:561-588

Run in the build container only (needs /root/reference):   python tests/golden/make_golden_mapops.py
Outputs tests/golden/mapops_*.npz.  Nothing under /root/reference is written or copied.
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
    from fl_slam_poc.backend.operators import primitive_association as pa
    from fl_slam_poc.backend.structures import primitive_map as pm

    from gc_slam_b200 import synth

    def atlas_of(td, next_global_id=10 ** 6):
        tile = pm.PrimitiveMapTile(**{k: (jnp.asarray(v) if isinstance(v, np.ndarray) else v) for k, v in td.items()})
        return pm.AtlasMap(tiles={int(td["tile_id"]): tile}, next_global_id=next_global_id, total_count=int(td["count"]),
                           m_tile=int(np.asarray(td["weights"]).shape[0]))

    def save(name, td, res_atlas, cert, eff, extra):
        nt = res_atlas.tiles[int(td["tile_id"])]
        out = {"in_" + k: np.asarray(td[k]) for k in FIELDS}
        out.update({"out_" + k: np.asarray(getattr(nt, k)) for k in FIELDS})
        out.update(tile_id=int(td["tile_id"]), count_in=int(td["count"]), next_local_id=int(td["next_local_id"]),
                   count_out=int(nt.count), total_count=int(res_atlas.total_count), next_global_id=int(res_atlas.next_global_id),
                   exact=cert.exact, triggers=np.array(cert.approximation_triggers, dtype="U64"),
                   mass_epsilon_ratio=cert.influence.mass_epsilon_ratio, predicted=eff.predicted, realized=eff.realized)
        out.update(extra)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
        print(name, {k: v for k, v in extra.items() if np.ndim(v) == 0})

    def dense_tile(n_surf, m_tile, seed, cam_fraction=0.3):
        atl = synth.synthetic_atlas(n_surf, m_tile, seed, scan_seq=30)
        tid = max(atl["tiles"], key=lambda t: atl["tiles"][t]["count"])
        td = dict(atl["tiles"][tid])
        # give some primitives camera support so that the rgb accumulators are exercised
        rng = np.random.default_rng(seed + 1)
        cam = (rng.random(m_tile) < cam_fraction) & td["valid_mask"]
        td["cam_mass"] = np.where(cam, 0.5 * td["weights"], 0.0)
        td["lidar_mass"] = td["weights"] - td["cam_mass"]
        col = rng.random((m_tile, 3))
        td["rgb_cam_accum"] = col * td["cam_mass"][:, None]
        td["rgb_cam_denom"] = td["cam_mass"].copy()
        td["rgb"] = np.where(cam[:, None], col, 0.5)
        td["colors"] = td["rgb"].copy()
        return td

    def proposals(n, seed):
        rng = np.random.default_rng(seed)
        a = rng.normal(size=(n, 3, 3))
        lam = np.einsum("nij,nkj->nik", a, a) + 0.5 * np.eye(3)[None]
        th = rng.normal(size=(n, 3)) * 3.0
        eta = rng.normal(size=(n, 3, 3))
        w = rng.random(n) + 0.05
        col = rng.random((n, 3)) * 1.4 - 0.2     # some outside [0,1]: exercises the clip
        src = (rng.random(n) < 0.6).astype(np.int32)
        return lam, th, eta, w, col, src

    # ---------------- fuse
    for name, m_tile, n, seed, full in (("mapops_fuse_masked_colors", 96, 700, 11, True), ("mapops_fuse_plain", 64, 150, 12, False),
                                        ("mapops_fuse_negative_slots", 80, 300, 13, True)):
        td = dense_tile(3000, m_tile, seed)
        rng = np.random.default_rng(seed + 7)
        lam, th, eta, w, col, src = proposals(n, seed + 3)
        slots = rng.integers(0, m_tile // 2, size=n).astype(np.int32)     # many repeats per slot
        if "negative" in name:     # .at[].add wraps indices in [-m_tile, -1]; jnp.unique counts -1 and m_tile - 1 separately
            slots = rng.integers(-m_tile, m_tile, size=n).astype(np.int32)
            slots[:4] = [-1, m_tile - 1, -m_tile, 0]
        resp = rng.random(n) * (rng.random(n) < 0.8)
        vm = rng.random(n) < 0.7
        kw = dict(valid_mask=jnp.asarray(vm), colors_meas=jnp.asarray(col), sources_meas=jnp.asarray(src)) if full else {}
        res, cert, eff = pm.primitive_map_fuse(atlas_of(td), int(td["tile_id"]), jnp.asarray(slots), jnp.asarray(lam), jnp.asarray(th),
                                               jnp.asarray(eta), jnp.asarray(w), jnp.asarray(resp), 12.5, 31, **kw)
        save(name, td, res.atlas_map, cert, eff, dict(slots=slots, lam=lam, th=th, eta=eta, w=w, resp=resp, vm=vm, col=col, src=src,
                                                      full=full, timestamp=12.5, scan_seq=31, n_fused=res.n_fused))

    # ---------------- insert_masked
    for name, m_tile, k, seed, full, n_surf in (("mapops_insert_evict", 96, 16, 21, True, 3000), ("mapops_insert_defaults", 64, 8, 22, False, 3000),
                                                ("mapops_insert_sparse_tile", 128, 24, 23, True, 150)):
        td = dense_tile(n_surf, m_tile, seed)
        rng = np.random.default_rng(seed + 7)
        lam, th, eta, w, col, src = proposals(k, seed + 3)
        vnew = rng.random(k) < 0.75
        kw = dict(colors_new=jnp.asarray(col), sources_new=jnp.asarray(src)) if full else {}
        res, cert, eff = pm.primitive_map_insert_masked(atlas_of(td, 777), int(td["tile_id"]), jnp.asarray(lam), jnp.asarray(th),
                                                        jnp.asarray(eta), jnp.asarray(w), 13.25, jnp.asarray(vnew), scan_seq=33, **kw)
        save(name, td, res.atlas_map, cert, eff, dict(lam=lam, th=th, eta=eta, w=w, vnew=vnew, col=col, src=src, full=full,
                                                      timestamp=13.25, scan_seq=33, n_inserted=res.n_inserted,
                                                      new_ids=np.asarray(res.new_ids), next_global_id_in=777))

    # ---------------- cull (threshold; max_primitives), forget
    for name, m_tile, seed, thr, maxp in (("mapops_cull_threshold", 96, 31, 0.35, None), ("mapops_cull_max_primitives", 96, 32, 0.05, 20),
                                          ("mapops_cull_nothing", 64, 33, 1e-9, None)):
        td = dense_tile(3000, m_tile, seed)
        rng = np.random.default_rng(seed + 7)
        td["weights"] = np.where(td["valid_mask"], rng.random(m_tile) + 1e-3, rng.random(m_tile) * 0.01)   # stale mass in free slots
        res, cert, eff = pm.primitive_map_cull(atlas_of(td), int(td["tile_id"]), weight_threshold=thr, max_primitives=maxp)
        save(name, td, res.atlas_map, cert, eff, dict(thr=thr, maxp=-1 if maxp is None else maxp, n_culled=res.n_culled,
                                                      mass_dropped=res.mass_dropped))
    td = dense_tile(3000, 96, 41)
    res, cert, eff = pm.primitive_map_forget(atlas_of(td), int(td["tile_id"]), forgetting_factor=0.9)
    save("mapops_forget", td, res.atlas_map, cert, eff, dict(gamma=0.9))

    # ---------------- block_associations_for_fuse
    rng = np.random.default_rng(51)
    n_tot, k_assoc, block = 600, 8, 256
    ar = pa.PrimitiveAssociationResult(responsibilities=jnp.asarray(rng.random((n_tot, k_assoc))),
                                       candidate_pool_indices=jnp.asarray(rng.integers(0, 4000, (n_tot, k_assoc)).astype(np.int32)),
                                       candidate_tile_ids=jnp.asarray(rng.integers(0, 2 ** 40, (n_tot, k_assoc)).astype(np.int64)),
                                       candidate_slots=jnp.asarray(rng.integers(0, 5000, (n_tot, k_assoc)).astype(np.int64)),
                                       row_masses=jnp.asarray(rng.random(n_tot)), cost_matrix=jnp.asarray(rng.random((n_tot, k_assoc))))
    vm = rng.random(n_tot) < 0.8
    mi, ct, cs, rs, vr = pa.block_associations_for_fuse(ar, jnp.asarray(vm), block)
    np.savez_compressed(os.path.join(HERE, "mapops_block_assoc.npz"), responsibilities=np.asarray(ar.responsibilities),
                        candidate_tile_ids=np.asarray(ar.candidate_tile_ids), candidate_slots=np.asarray(ar.candidate_slots),
                        valid_mask=vm, block=block, out_meas_idx=np.asarray(mi), out_tile_ids=np.asarray(ct), out_slots=np.asarray(cs),
                        out_resp=np.asarray(rs), out_valid_rows=np.asarray(vr))
    print("mapops_block_assoc", np.asarray(mi).shape)


if __name__ == "__main__":
    main()
