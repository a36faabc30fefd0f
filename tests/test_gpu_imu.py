"""
IMU scan-twist prologue on the device against the reference's own outputs (tests/golden/imu_*.npz,
made by tests/golden/make_golden_imu.py) and against oracle/imu.py on fresh seeded inputs.  All calls go through the
C-ABI entry gcs_imu_scan_twist.  Tolerances: weights 1e-13 (libdevice vs NumPy exp), integrated quantities 1e-11
relative (the device combines per-thread runs in a fixed order instead of one sequential recurrence), twist 1e-10.
"""
import glob
import os

import numpy as np
import pytest

from conftest import GOLDEN, golden, rel_err

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")

IMU_CASES = sorted(os.path.basename(p) for p in glob.glob(os.path.join(GOLDEN, "imu_*.npz")))
KEYS = ("delta_pose", "delta_R", "delta_p", "delta_v", "ess", "a_body_mean", "a_world_nog_mean", "a_world_mean", "dt_eff_sum")


def _np(t):
    return t.detach().cpu().numpy()


@pytest.fixture(scope="module")
def imu():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    from gc_slam_b200 import imu as m
    return m


@pytest.mark.parametrize("case", IMU_CASES)
def test_imu_scan_twist_vs_reference_golden(imu, case):
    g = golden(case)
    H = g["hp_sigma"].shape[0]
    res = imu.imu_scan_twist(g["stamps"], g["gyro"], g["accel"], float(g["t0"]), float(g["t1"]), g["hp_sigma"],
                             g["hp_rotvec0"], g["hp_gyro_bias"], g["hp_accel_bias"], g["gravity"],
                             deskew_rotation_only=bool(g["rotation_only"]), want_weights=True)
    assert rel_err(_np(res.weights), g["weights"]) < 1e-13
    for k in KEYS:
        assert rel_err(_np(getattr(res, k)).reshape(g[k].shape), g[k]) < 1e-11, k
    assert rel_err(_np(res.xi_body), g["xi_body"]) < 1e-10
    assert res.xi_body.shape == (H, 6) and res.xi_body.is_contiguous()
    # two runs are bit-identical (fixed-order combine)
    res2 = imu.imu_scan_twist(g["stamps"], g["gyro"], g["accel"], float(g["t0"]), float(g["t1"]), g["hp_sigma"],
                              g["hp_rotvec0"], g["hp_gyro_bias"], g["hp_accel_bias"], g["gravity"],
                              deskew_rotation_only=bool(g["rotation_only"]))
    assert torch.equal(res.xi_body, res2.xi_body) and torch.equal(res.delta_R, res2.delta_R)


@pytest.mark.parametrize("case", IMU_CASES[:2])
def test_reference_signature_operators(imu, case):
    """smooth_window_weights and preintegrate_imu_relative_pose with the reference's own argument lists."""
    g = golden(case)
    w = imu.smooth_window_weights(g["stamps"], float(g["t0"]), float(g["t1"]), float(g["hp_sigma"][1]))
    assert rel_err(_np(w), g["weights"][1]) < 1e-13
    res = imu.preintegrate_imu_relative_pose(g["stamps"], g["gyro"], g["accel"], g["weights"][1], g["hp_rotvec0"][1],
                                             g["hp_gyro_bias"][1], g["hp_accel_bias"][1], g["gravity"])
    for k in KEYS:
        assert rel_err(_np(getattr(res, k)).reshape(g[k][1].shape), g[k][1]) < 1e-11, k


@pytest.mark.parametrize("M,n_valid,H", [(1, 1, 1), (2, 2, 2), (127, 100, 1), (512, 21, 64), (4096, 4000, 3)])
def test_imu_scan_twist_vs_oracle_sizes(imu, M, n_valid, H):
    from gc_slam_b200 import synth
    from oracle import imu as oimu
    stamps, gyro, accel = synth.imu_window(M, n_valid, 70 + M, t_start=synth.EPOCH_T0)
    hp = synth.imu_hypothesis_params(H, 80 + M)
    t0, t1 = synth.EPOCH_T0, synth.EPOCH_T0 + 0.1
    g = np.array([0.0, 0.0, -9.81])
    res = imu.imu_scan_twist(stamps, gyro, accel, t0, t1, hp["sigma"], hp["rotvec0"], hp["gyro_bias"], hp["accel_bias"], g)
    for h in sorted({0, H // 2, H - 1}):
        o = oimu.imu_scan_twist(stamps, gyro, accel, t0, t1, float(hp["sigma"][h]), hp["rotvec0"][h], hp["gyro_bias"][h],
                                hp["accel_bias"][h], g)
        for k in KEYS:
            assert rel_err(_np(getattr(res, k))[h].reshape(np.shape(o[k])), o[k]) < 1e-11, (k, h)
        assert np.max(np.abs(_np(res.xi_body)[h] - o["xi_body"])) < 1e-10 * (1.0 + np.max(np.abs(o["xi_body"])))


def test_twist_feeds_bin_plan_on_device(imu):
    """xi_body written by the prologue into a BinPathPlan's xi rows == uploading the oracle's twist from the host."""
    from gc_slam_b200 import operators as ops, synth
    from oracle import imu as oimu
    H, n = 4, 4096
    bins = synth.fibonacci_atlas(48)
    pts, t, w, ring, tag = synth.vlp16_scan(n, 11, t0=synth.EPOCH_T0)
    stamps, gyro, accel = synth.imu_window(512, 40, 12, t_start=synth.EPOCH_T0)
    hp = synth.imu_hypothesis_params(H, 13)
    t0, t1 = synth.EPOCH_T0, synth.EPOCH_T0 + 0.1
    g = np.array([0.0, 0.0, -9.81])
    poses = synth.hypothesis_poses(H, 3)
    outs = []
    for mode in ("device", "host"):
        plan = ops.BinPathPlan(1, n, n, n_hyp=H, n_bins=48, tau=0.1, origin=synth.lidar_origin_base(), want_evidence=True)
        plan.set_bins(bins, 0.1)
        plan.set_map(synth.random_map_bin_stats(48, 7, bins))
        if mode == "device":
            plan.upload(pts[None], t[None], w[None], ring[None], tag[None], np.array([t0]), np.array([t1]), None, poses,
                        non_blocking=False)
            plan.set_twist_from_imu(0, stamps, gyro, accel, t0, t1, hp["sigma"], hp["rotvec0"], hp["gyro_bias"],
                                    hp["accel_bias"], g)
        else:
            xi = np.stack([oimu.imu_scan_twist(stamps, gyro, accel, t0, t1, float(hp["sigma"][h]), hp["rotvec0"][h],
                                               hp["gyro_bias"][h], hp["accel_bias"][h], g)["xi_body"] for h in range(H)])
            plan.upload(pts[None], t[None], w[None], ring[None], tag[None], np.array([t0]), np.array([t1]), xi, poses,
                        non_blocking=False)
        plan.run()
        torch.cuda.synchronize()
        outs.append(plan.outputs())
    a, b = outs
    assert rel_err(_np(a.deskewed["points"]), _np(b.deskewed["points"])) < 1e-10
    assert rel_err(_np(a.L22), _np(b.L22)) < 1e-7


def test_imu_argument_errors(imu):
    with pytest.raises(ValueError):
        imu.smooth_window_weights(np.zeros(0), 0.0, 1.0, 0.01)
    with pytest.raises(ValueError):
        imu.imu_scan_twist(np.zeros(8), np.zeros((7, 3)), np.zeros((8, 3)), 0.0, 1.0, 0.01, np.zeros(3), np.zeros(3), np.zeros(3))
