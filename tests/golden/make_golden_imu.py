#!/usr/bin/env python
"""
Golden vectors for the IMU twist prologue (window weights -> preintegration -> xi_body), produced by executing the
REFERENCE's own sources (fl/backend/operators/imu_preintegration.py, fl/common/geometry/se3_jax.py) on top of
oracle/jax_shim (NumPy stand-in for the JAX runtime, which is not installed here).

Run in the build container only (needs /root/reference):   python tests/golden/make_golden_imu.py
Outputs tests/golden/imu_*.npz (small, committed).  Nothing under /root/reference is written or copied.
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


def main():
    from fl_slam_poc.backend.operators.imu_preintegration import preintegrate_imu_relative_pose_jax, smooth_window_weights
    from fl_slam_poc.common import constants
    from fl_slam_poc.common.geometry import se3_jax
    from gc_slam_b200 import synth

    gravity = np.asarray(constants.GC_GRAVITY_W, dtype=np.float64)
    cases = {
        # name: (M, n_valid, seed, t_start, scan_len, rotation_only, gyro_scale)
        "imu_c1_512_epoch": (512, 40, 31, synth.EPOCH_T0, 0.1, False, 0.4),        # pipeline shape: zero-padded buffer
        "imu_c2_512_full_relative": (512, 512, 32, 0.0, 2.5, False, 0.2),           # every slot valid, relative stamps
        "imu_c3_64_fast_rotonly": (64, 30, 33, synth.EPOCH_T0, 0.1, True, 3.0),     # fast rotation, rotation-only deskew
        "imu_c4_777_ragged": (777, 500, 34, 10.0, 2.0, False, 0.8),                 # length not a multiple of anything
    }
    for name, (M, n_valid, seed, t_start, scan_len, rot_only, gscale) in cases.items():
        stamps, gyro, accel = synth.imu_window(M, n_valid, seed, t_start=t_start, gyro_scale=gscale)
        hp = synth.imu_hypothesis_params(3, seed + 100)
        out = {"stamps": stamps, "gyro": gyro, "accel": accel, "t0": t_start, "t1": t_start + scan_len,
               "gravity": gravity, "rotation_only": rot_only, **{"hp_" + k: v for k, v in hp.items()}}
        keys = ("delta_pose", "delta_R", "delta_p", "delta_v", "ess", "a_body_mean", "a_world_nog_mean", "a_world_mean",
                "dt_eff_sum")
        acc = {k: [] for k in keys + ("weights", "xi_body")}
        for h in range(3):
            w = smooth_window_weights(stamps, t_start, t_start + scan_len, float(hp["sigma"][h]))
            res = preintegrate_imu_relative_pose_jax(stamps, gyro, accel, w, hp["rotvec0"][h], hp["gyro_bias"][h],
                                                     hp["accel_bias"][h], gravity)
            xi = np.array(se3_jax.se3_log(res[0]))
            xi[:3] = xi[:3] * (0.0 if rot_only else 1.0)      # pipeline.py:479-483
            for k, v in zip(keys, res):
                acc[k].append(np.asarray(v, dtype=np.float64))
            acc["weights"].append(np.asarray(w, dtype=np.float64))
            acc["xi_body"].append(xi)
        out.update({k: np.stack(v) for k, v in acc.items()})
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
        print(name, "xi_body[0] =", out["xi_body"][0])


if __name__ == "__main__":
    main()
