#!/usr/bin/env python
"""
Golden vectors for the PointCloud2 ingest step, produced by the REFERENCE's own parser.

Run in the build container only (needs /root/reference):   python tests/golden/make_golden_pc2.py

`parse_pointcloud2_vlp16` lives in backend_node.py, whose module import needs ROS (rclpy, sensor_msgs), absent here.
The function itself only needs NumPy, the PointField datatype codes and three constants, so this script compiles
the source lines of that one function (and of `_pointfield_to_dtype`) straight from the reference file at run time,
in a namespace holding a stub PointField and the reference's real constants module, and calls it on seeded messages.
Nothing from the reference is copied into the repository; the outputs are tests/golden/pc2_*.npz.
"""
import ast
import importlib.util
import os
import sys
from types import SimpleNamespace
from typing import Tuple  # noqa: F401  (used by the reference's annotations)

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF_PKG = "/root/reference/fl_ws/src/fl_slam_poc/fl_slam_poc"
sys.path.insert(0, ROOT)


def load_reference_parser():
    spec = importlib.util.spec_from_file_location("ref_constants", os.path.join(REF_PKG, "common", "constants.py"))
    constants = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(constants)
    src = open(os.path.join(REF_PKG, "backend", "backend_node.py")).read()
    tree = ast.parse(src)
    wanted = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name in ("_pointfield_to_dtype", "parse_pointcloud2_vlp16")]
    assert len(wanted) == 2
    PointField = SimpleNamespace(INT8=1, UINT8=2, INT16=3, UINT16=4, INT32=5, UINT32=6, FLOAT32=7, FLOAT64=8)
    ns = {"np": np, "constants": constants, "PointField": PointField, "PointCloud2": object, "Tuple": Tuple}
    exec(compile(ast.Module(body=wanted, type_ignores=[]), "backend_node.py", "exec"), ns)
    return ns["parse_pointcloud2_vlp16"]


def make_msg(data, n, fields, point_step, stamp=(1700000000, 250000000)):
    return SimpleNamespace(width=n, height=1, point_step=point_step, data=bytes(data),
                           fields=[SimpleNamespace(name=k, offset=v[0], datatype=v[1]) for k, v in fields.items()],
                           header=SimpleNamespace(stamp=SimpleNamespace(sec=stamp[0], nanosec=stamp[1])))


def main():
    from gc_slam_b200 import synth
    parse = load_reference_parser()
    R, t = synth.base_lidar_extrinsics()
    cases = {}
    # 1: driver layout, per-point time in seconds
    cases["pc2_vlp16_seconds"] = synth.vlp16_pointcloud2(4096, 21, time_unit="s")
    # 2: per-point time in nanoseconds (triggers the 1e-9 rescale), ragged size
    cases["pc2_vlp16_ns_3001"] = synth.vlp16_pointcloud2(3001, 22, time_unit="ns")
    # 3: no time field, uint8 ring, float64 z, padded point_step, NaN / +-inf coordinates
    rng = np.random.default_rng(23)
    n = 1000
    dt = np.dtype({"names": ["x", "y", "z", "ring"], "formats": ["<f4", "<f4", "<f8", "u1"], "offsets": [0, 4, 8, 20], "itemsize": 32})
    rec = np.zeros(n, dtype=dt)
    rec["x"] = rng.normal(0, 8, n); rec["y"] = rng.normal(0, 8, n); rec["z"] = rng.normal(0, 1, n)
    rec["ring"] = rng.integers(0, 16, n)
    rec["x"][5] = np.nan; rec["y"][6] = np.inf; rec["z"][7] = -np.inf; rec["x"][8] = -np.inf; rec["z"][9] = np.nan
    rec["x"][10] = rec["y"][10] = 0.0; rec["z"][10] = 0.0            # zero-range return
    cases["pc2_no_time_nonfinite"] = (np.frombuffer(rec.tobytes(), np.uint8).copy(),
                                      {"x": (0, 7), "y": (4, 7), "z": (8, 8), "ring": (20, 2)}, 32)
    for name, (data, fields, step) in cases.items():
        n = data.size // step
        msg = make_msg(data, n, fields, step)
        pts, ts, w, ring, tag = parse(msg)
        pts_base = (R @ pts.T).T + t[None, :]          # backend_node.py:1682-1684, same expression
        keys = sorted(fields)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), data=data, point_step=step, n_points=n,
                            field_names=np.array(keys), field_offsets=np.array([fields[k][0] for k in keys]),
                            field_types=np.array([fields[k][1] for k in keys]),
                            header_stamp=np.float64(1700000000 + 250000000 * 1e-9),
                            points=pts, points_base=pts_base, t=ts, w=w, ring=ring, tag=tag, R=R, t_base=t)
        print(name, n, "points; t range", ts.min(), ts.max(), "w range", w.min(), w.max())


if __name__ == "__main__":
    main()
