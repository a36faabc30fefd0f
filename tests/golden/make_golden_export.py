#!/usr/bin/env python
"""
Golden vectors for the map export (per-tile view -> renderable batch -> /gc/map/points payload), produced by the
REFERENCE's own code: extract_primitive_map_view and renderable_batch_from_view are imported from
fl/backend/structures/primitive_map.py on top of oracle/jax_shim; _build_pointcloud2_from_view is compiled from the
source lines of fl/backend/map_publisher.py (its module needs ROS) with stub message classes, as make_golden_pc2.py does
for the parser.  The publisher's concatenation and recency order (map_publisher.py:190-203, plain NumPy: concatenate +
np.lexsort) are applied to those outputs here.

Run in the build container only (needs /root/reference):   python tests/golden/make_golden_export.py
Outputs tests/golden/export_*.npz.  Nothing under /root/reference is written or copied.
"""
import ast
import os
import sys
import types
from typing import Optional  # noqa: F401

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
REF_PKG = os.path.join(REF, "fl_ws", "src", "fl_slam_poc", "fl_slam_poc")
sys.path.insert(0, os.path.join(ROOT, "oracle", "jax_shim"))
sys.path.insert(0, os.path.join(REF, "fl_ws", "src", "fl_slam_poc"))
sys.path.insert(0, ROOT)


class _Bag:
    FLOAT32 = 7

    def __init__(self, **kw):
        self.__dict__.update(kw)


def load_cloud_builder():
    src = open(os.path.join(REF_PKG, "backend", "map_publisher.py")).read()
    tree = ast.parse(src)
    wanted = [n for n in tree.body if isinstance(n, ast.FunctionDef)
              and n.name in ("_pointcloud2_fields_xyz_intensity", "_build_pointcloud2_from_view")]
    assert len(wanted) == 2
    fake = types.ModuleType("builtin_interfaces.msg")
    fake.Time = _Bag
    sys.modules.setdefault("builtin_interfaces", types.ModuleType("builtin_interfaces"))
    sys.modules["builtin_interfaces.msg"] = fake
    ns = {"np": np, "PointCloud2": _Bag, "Header": _Bag, "PointField": _Bag, "Optional": Optional}
    exec(compile(ast.Module(body=wanted, type_ignores=[]), "map_publisher.py", "exec"), ns)
    return ns["_build_pointcloud2_from_view"]


def main():
    import jax.numpy as jnp
    from fl_slam_poc.backend.structures import primitive_map as pm

    from gc_slam_b200 import synth

    build_cloud = load_cloud_builder()
    cases = {"export_tiles": (2500, 4000, 51, None), "export_sparse_with_empty_tile": (700, 1500, 52, None)}
    for name, (n_surf, m_tile, seed, max_prim) in cases.items():
        atl = synth.synthetic_atlas(n_surf, m_tile, seed, scan_seq=30)
        tids = sorted(atl["tiles"].keys())
        if "empty" in name:      # a tile without any valid slot
            t0 = tids[0]
            atl["tiles"][t0]["valid_mask"][:] = False
            atl["tiles"][t0]["count"] = 0
        views, last = [], []
        for tid in tids:
            td = atl["tiles"][tid]
            tile = pm.PrimitiveMapTile(**{k: (jnp.asarray(v) if isinstance(v, np.ndarray) else v) for k, v in td.items()})
            v = pm.extract_primitive_map_view(tile=tile, max_primitives=max_prim)
            views.append(v)
            last.append(np.asarray(np.asarray(td["last_supported_scan_seq"])[np.asarray(v.slot_indices)], dtype=np.int64))
        positions = np.concatenate([np.array(v.positions) for v in views], axis=0)
        weights = np.concatenate([np.array(v.weights) for v in views], axis=0)
        colors = np.concatenate([np.array(v.colors) for v in views], axis=0)
        recency = np.concatenate(last, axis=0)
        pids = np.concatenate([np.array(v.primitive_ids) for v in views], axis=0)
        order = np.lexsort((pids, -recency))
        msg = build_cloud(positions[order], weights[order], colors[order], "odom", 12.5)
        Sigma = np.concatenate([np.array(v.covariances) for v in views], axis=0)[order]
        eta = np.concatenate([np.array(v.etas) for v in views], axis=0)[order]
        rb = [pm.renderable_batch_from_view(v) for v in views]
        Lam_views = np.concatenate([b.Lambda_world for b in rb if b.Lambda_world is not None and b.mu_world.shape[0]], axis=0)[order]
        out = dict(mu_world=positions[order], Sigma_world=Sigma, Lambda_world=Lam_views, eta=eta, mass=weights[order],
                   color=colors[order], primitive_ids=pids[order], last_supported_scan_seq=recency[order],
                   cloud=np.frombuffer(bytes(msg.data), dtype=np.uint8), point_step=msg.point_step, width=msg.width,
                   tile_ids=np.asarray(tids, dtype=np.int64), m_tile=m_tile, n_surf=n_surf, seed=seed,
                   slot_counts=np.asarray([v.count for v in views]))
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
        print(name, "n =", positions.shape[0], "tiles", len(tids), "cloud bytes", len(msg.data))


if __name__ == "__main__":
    main()
