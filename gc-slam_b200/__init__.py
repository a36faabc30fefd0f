"""
gc_slam_b200 -- B200 (sm_100a) implementation of GC-SLAM v2's per-scan LiDAR evidence path.

Layout
  csrc/          CUDA kernels + the C-ABI (include/gcs_b200.h) -> lib/libgcs_b200.so
  _lib.py        ctypes binding of the C-ABI (raises if the library is missing: no CPU fallback)
  certs.py       CertBundle / ExpectedEffect contract types (same field names as the reference)
  constants.py   budgets and epsilons of the path
  operators.py   bin family: point_budget_resample, deskew_constant_twist, bin_soft_assign, ...
  primitives.py  primitive family: extract_lidar_surfels, associate_primitives_ot, map update, ...
  imu.py         IMU window weights + preintegration -> scan twist xi_body (the step in front of the deskew)
  hypothesis_batch.py  the per-scan hypothesis loop, batched: H hypotheses of one scan through the primitive family at once
  fusion.py      evidence fusion for all hypotheses in one launch (tempering, prior scaling, fusion scale, additive fusion)
  sharding.py    scan / point sharding over ranks, hypothesis combine
  synth.py       seeded synthetic scans / maps
Sub-modules are imported on demand so that ``import gc_slam_b200`` itself never touches CUDA.
"""

__version__ = "0.1.0"

_LAZY = ("constants", "certs", "synth", "_lib", "operators", "primitives", "imu", "fusion", "manifest", "sharding", "build", "hypothesis_batch")


def __getattr__(name):
    if name in _LAZY:
        import importlib

        return importlib.import_module(f"{__name__}.{name}")
    raise AttributeError(name)
