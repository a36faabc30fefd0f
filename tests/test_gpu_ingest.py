"""
GPU parity of the PointCloud2 ingest step through the C-ABI: device decode vs the vectors the
reference's own parse_pointcloud2_vlp16 produced (tests/golden/pc2_*.npz), and the bin path fed from wire bytes vs
the oracle fed from the oracle's decode.  Integer outputs bit-exact; coordinates and timestamps bit-exact (exact
float32 -> float64 widening); weights 1e-13 (libdevice exp vs NumPy exp); base-frame points 1e-15.
"""
import numpy as np
import pytest

from conftest import rel_err
from test_oracle_pc2_vs_golden import PC2_CASES, load_case

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")


@pytest.fixture(scope="module")
def ops():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from gc_slam_b200 import operators
    return operators


def _np(t):
    return t.detach().cpu().numpy()


@pytest.mark.parametrize("case", PC2_CASES)
def test_parse_pointcloud2_vs_reference_vectors(ops, case):
    g, fields = load_case(case)
    n, step = int(g["n_points"]), int(g["point_step"])
    pts, t, w, ring, tag, info = ops.parse_pointcloud2_vlp16(g["data"], n, step, fields, float(g["header_stamp"]))
    assert np.array_equal(_np(pts), g["points"]) and np.array_equal(_np(t), g["t"])
    assert np.array_equal(_np(ring), g["ring"]) and np.array_equal(_np(tag), g["tag"])
    assert np.max(np.abs(_np(w) - g["w"])) < 1e-13
    assert info["n_nonfinite"] == int(np.sum(~np.isfinite(np.stack([np.frombuffer(g["data"].tobytes(), dtype=np.dtype(
        {"names": ["v"], "formats": [{7: "<f4", 8: "<f8"}[fields[k][1]]], "offsets": [fields[k][0]], "itemsize": step}), count=n)["v"]
        for k in ("x", "y", "z")]))))
    assert info["time_rescaled"] == ("ns" in case)
    # fused with the base transform
    pb, *_ = ops.parse_pointcloud2_vlp16(g["data"], n, step, fields, float(g["header_stamp"]), g["R"], g["t_base"])
    assert rel_err(_np(pb), g["points_base"]) < 1e-15


def test_parse_errors_and_empty(ops):
    f = {"x": (0, 7), "y": (4, 7), "z": (8, 7), "ring": (16, 4)}
    out = ops.parse_pointcloud2_vlp16(b"", 0, 22, f, 1.0)
    assert out[0].shape == (0, 3) and out[3].dtype == torch.uint8
    with pytest.raises(RuntimeError):
        ops.parse_pointcloud2_vlp16(b"\0" * 22, 1, 22, {"x": (0, 7), "y": (4, 7), "z": (8, 7)}, 1.0)
    with pytest.raises(ValueError):
        ops.parse_pointcloud2_vlp16(b"\0" * 22, 1, 22, {"x": (0, 7), "y": (4, 7), "z": (8, 7), "ring": (21, 4)}, 1.0)  # field past point_step
    with pytest.raises(ValueError):
        ops.parse_pointcloud2_vlp16(b"\0" * 10, 1, 22, f, 1.0)   # short payload


def test_bin_path_from_wire_bytes_vs_oracle(ops):
    """3 messages of 20,000 points (ragged vs the 256-point decode tiles and the 16-byte staging words) through
    BinPathPlan.upload_pointcloud2 -> full bin path, against oracle decode -> oracle bin path."""
    from gc_slam_b200 import synth
    from oracle import bin_path as ob
    from oracle import lie, pc2
    S, n, cap = 3, 20000, 8192
    msgs = [synth.vlp16_pointcloud2(n, 700 + k, time_unit="s") for k in range(S)]
    fields, step = msgs[0][1], msgs[0][2]
    R, tb = synth.base_lidar_extrinsics()
    bins = synth.fibonacci_atlas(48)
    ms = synth.random_map_bin_stats(48, 7, bins)
    xi = np.stack([synth.scan_twist(60 + k) for k in range(S)])
    poses = synth.hypothesis_poses(S, 3)
    t0, t1 = np.zeros(S), np.full(S, 0.1)
    plan = ops.BinPathPlan(S, n, cap, n_hyp=1, n_bins=48, tau=0.1, origin=synth.lidar_origin_base(), want_evidence=True,
                           materialize_deskewed=True, materialize_resampled=True)
    plan.set_bins(bins, 0.1)
    plan.set_map(ms)
    plan.enable_pointcloud2(fields, step, R, tb)
    payload = np.concatenate([m[0] for m in msgs])
    moved = plan.upload_pointcloud2(payload, None, t0, t1, xi, poses, non_blocking=False)
    assert moved == S * n * step + 8 * (S + S + 6 * S + 6 * S)
    plan.run()
    torch.cuda.synchronize()
    out = plan.outputs()
    for k in range(S):
        p, t, w, ring, tag = pc2.parse_pointcloud2_vlp16(msgs[k][0].tobytes(), n, step, fields, 0.0)
        pb = pc2.lidar_to_base(p, R, tb)
        o = ob.lidar_evidence_bins(pb, t, w, ring, tag, cap, xi[k], 0.0, 0.1, synth.lidar_origin_base(), bins, 0.1, ms,
                                   lie.so3_exp(poses[k, 3:]), poses[k, :3])
        assert np.array_equal(_np(out.resampled["ring"][k]), o["resample"]["ring"])
        assert rel_err(_np(out.resampled["points"][k]), o["resample"]["points"]) < 1e-15
        assert rel_err(_np(out.stats["N"][k]), o["stats"]["N"]) < 1e-9
        assert rel_err(_np(out.stats["Sigma_p"][k]), o["stats"]["Sigma_p"]) < 1e-9
        assert rel_err(_np(out.L22[k]), o["L"]) < 1e-7
