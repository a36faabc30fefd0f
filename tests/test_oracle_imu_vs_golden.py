"""oracle/imu.py against the vectors produced by the reference's own IMU preintegration and se3_log
(tests/golden/make_golden_imu.py; fl/backend/operators/imu_preintegration.py:19-146, fl/common/geometry/se3_jax.py:178-256,
fl/backend/pipeline.py:436-483).  CPU only."""
import glob
import os

import numpy as np
import pytest

from conftest import GOLDEN, golden, rel_err

IMU_CASES = sorted(os.path.basename(p) for p in glob.glob(os.path.join(GOLDEN, "imu_*.npz")))
KEYS = ("delta_pose", "delta_R", "delta_p", "delta_v", "ess", "a_body_mean", "a_world_nog_mean", "a_world_mean", "dt_eff_sum")


def test_cases_present():
    assert len(IMU_CASES) >= 4


@pytest.mark.parametrize("case", IMU_CASES)
def test_oracle_imu_matches_reference(case):
    from oracle import imu
    g = golden(case)
    for h in range(g["hp_sigma"].shape[0]):
        out = imu.imu_scan_twist(g["stamps"], g["gyro"], g["accel"], float(g["t0"]), float(g["t1"]), float(g["hp_sigma"][h]),
                                 g["hp_rotvec0"][h], g["hp_gyro_bias"][h], g["hp_accel_bias"][h], g["gravity"],
                                 deskew_rotation_only=bool(g["rotation_only"]))
        assert np.array_equal(out["weights"], g["weights"][h])
        for k in KEYS:
            assert rel_err(out[k], g[k][h]) < 1e-13, (k, h)
        assert rel_err(out["xi_body"], g["xi_body"][h]) < 1e-12, h


def test_se3_log_inverts_se3_exp():
    """Property the reference pins (test/test_audit_invariants.py:221-330 style): exp(log(T)) == T."""
    from oracle import imu, lie
    rng = np.random.default_rng(5)
    for _ in range(20):
        T = np.concatenate([rng.uniform(-2, 2, 3), rng.uniform(-1.0, 1.0, 3)])
        assert np.allclose(lie.se3_exp(imu.se3_log(T)), T, atol=1e-10)
    T = np.array([0.3, -0.2, 0.1, 1e-9, -2e-9, 5e-10])      # small-angle branch
    assert np.allclose(lie.se3_exp(imu.se3_log(T)), T, atol=1e-12)
