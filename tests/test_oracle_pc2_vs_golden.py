"""oracle/pc2.py against the vectors produced by the reference's own parse_pointcloud2_vlp16
(tests/golden/make_golden_pc2.py; backend_node.py:377-468, :1677-1690).  CPU only."""
import glob
import os

import numpy as np
import pytest

from conftest import GOLDEN

PC2_CASES = sorted(os.path.basename(p) for p in glob.glob(os.path.join(GOLDEN, "pc2_*.npz")))


def load_case(name):
    g = np.load(os.path.join(GOLDEN, name))
    fields = {str(k): (int(o), int(t)) for k, o, t in zip(g["field_names"], g["field_offsets"], g["field_types"])}
    return g, fields


@pytest.mark.parametrize("case", PC2_CASES)
def test_oracle_parse_matches_reference(case):
    from oracle import pc2
    g, fields = load_case(case)
    pts, t, w, ring, tag = pc2.parse_pointcloud2_vlp16(g["data"].tobytes(), int(g["n_points"]), int(g["point_step"]), fields,
                                                       float(g["header_stamp"]))
    assert np.array_equal(pts, g["points"]) and np.array_equal(t, g["t"]) and np.array_equal(w, g["w"])
    assert np.array_equal(ring, g["ring"]) and np.array_equal(tag, g["tag"]) and ring.dtype == np.uint8
    assert np.array_equal(pc2.lidar_to_base(pts, g["R"], g["t_base"]), g["points_base"])


def test_oracle_parse_edge_cases():
    from oracle import pc2
    out = pc2.parse_pointcloud2_vlp16(b"", 0, 22, {"x": (0, 7), "y": (4, 7), "z": (8, 7), "ring": (16, 4)}, 1.0)
    assert out[0].shape == (0, 3) and out[3].dtype == np.uint8
    with pytest.raises(RuntimeError):
        pc2.parse_pointcloud2_vlp16(b"\0" * 22, 1, 22, {"x": (0, 7), "y": (4, 7), "z": (8, 7)}, 1.0)
