"""
Sensor inputs of the pipeline-chain golden (make_golden_pipeline.py), regenerated from seeds: shared by the generator and
by the tests that chain the same scans through the oracle / the CUDA path.  No reference or shim imports here.
"""
import numpy as np

N_RAW, N_IMU, N_CAM = 12000, 512, 60
SCANS = [dict(seed=3101, scan_seq=1), dict(seed=3102, scan_seq=2), dict(seed=3103, scan_seq=3), dict(seed=3104, scan_seq=4)]
T_FIRST = 1.6e9 + 12.0          # epoch stamps, like the bags
FULL_STRIDE = 53
X_ANCHOR = [0.6, 0.5, 0.3, 0.0, 0.0, 0.0]


def scan_inputs(k):
    """Sensor inputs of scan k (0-based), all from seeds: shared by the generator and the tests."""
    from gc_slam_b200 import synth
    sc = SCANS[k]
    t0 = T_FIRST + 0.1 * k
    pts, t, w, ring, tag = synth.vlp16_scan(N_RAW, sc["seed"], t0=t0)
    rng = np.random.default_rng(sc["seed"] + 17)
    imu_t = np.linspace(t0 - 0.1, t0 + 0.1, N_IMU)
    gyro = np.array([0.01, -0.02, 0.15])[None] + 0.01 * rng.standard_normal((N_IMU, 3))
    accel = np.array([0.05, -0.03, 9.81])[None] + 0.05 * rng.standard_normal((N_IMU, 3))
    cam = synth.camera_splats(N_CAM, sc["seed"] + 3)
    odom_twist = np.array([0.3, 0.0, 0.0, 0.0, 0.0, 0.15])
    odom_pose = np.array([X_ANCHOR[0] + 0.03 * (k + 1), X_ANCHOR[1] + 0.002 * (k + 1), X_ANCHOR[2], 0.0, 0.0, 0.015 * (k + 1)])
    return dict(points=pts, timestamps=t, weights=w, ring=ring, tag=tag, t0=t0, t1=t0 + 0.1, imu_t=imu_t, gyro=gyro, accel=accel,
                cam=cam, odom_twist=odom_twist, odom_pose=odom_pose, scan_seq=sc["scan_seq"])


