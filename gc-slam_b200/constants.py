"""
Budgets and epsilons of the LiDAR evidence path.  Values are the reference's
(fl_ws/src/fl_slam_poc/fl_slam_poc/common/constants.py, line cited per name); they are part of
the operator contract, so the drop-in must use the same numbers.
"""

GC_CHART_ID = "GC-RIGHT-01"  # :55
GC_D_Z = 22  # :58
GC_K_HYP = 4  # :62
GC_N_POINTS_CAP = 8192  # :64
GC_EPS_PSD = 1e-12  # :70
GC_EPS_LIFT = 1e-9  # :71
GC_EPS_MASS = 1e-12  # :72
GC_EPS_R = 1e-6  # :73
GC_KAPPA_BLEND_R0 = 0.8  # :95
GC_KAPPA_BLEND_TAU = 0.03  # :96
GC_TIME_WARP_SIGMA_FRAC = 0.1  # :141
GC_WEIGHT_FLOOR = 1e-12  # :237
GC_MAX_IMU_PREINT_LEN = 512  # :67  IMU buffer length handed to the preintegration (zero-padded)
GC_GRAVITY_W = (0.0, 0.0, -9.81)  # :80  Z-up world, gravity along -Z
GC_NONFINITE_SENTINEL = 1e6  # :238
GC_RANGE_WEIGHT_SIGMA = 0.25  # :241
GC_RANGE_WEIGHT_MIN_R = 0.5  # :242
GC_RANGE_WEIGHT_MAX_R = 50.0  # :243
GC_N_FEAT = 512  # :350
GC_N_SURFEL = 1024  # :353
GC_K_ASSOC = 8  # :356
GC_K_SINKHORN = 50  # :357
GC_PRIMITIVE_MAP_MAX_SIZE = 50000  # :392
GC_H_TILE = 2.0  # :408
GC_R_ACTIVE_TILES_XY = 1  # :411
GC_R_ACTIVE_TILES_Z = 0  # :412
GC_R_STENCIL_TILES_XY = 1  # :415
GC_R_STENCIL_TILES_Z = 0  # :416
GC_RECENCY_DECAY_LAMBDA = 0.02  # :419
GC_RECENCY_MIN_SCALE = 0.05  # :420
GC_N_ACTIVE_TILES = 7  # :430  (2*Rz+1)*(1+3r(r+1))
GC_N_STENCIL_TILES = 7  # :433
GC_M_TILE_VIEW = 1024  # :436
GC_M_TILE = GC_PRIMITIVE_MAP_MAX_SIZE  # :439
GC_PRIMITIVE_FORGETTING_FACTOR = 0.995  # :442
GC_PRIMITIVE_MERGE_THRESHOLD = 0.1  # :445
GC_K_MERGE_PAIRS_PER_TILE = 4  # :448
GC_PRIMITIVE_MERGE_MAX_TILE_SIZE = 2048  # :450
GC_PRIMITIVE_CULL_WEIGHT_THRESHOLD = 1e-4  # :453
GC_PRIMITIVE_KAPPA_MIN = 1e-3  # :456
GC_PRIMITIVE_KAPPA_MAX = 1e4  # :457
GC_VMF_N_LOBES = 3  # :463
GC_FUSE_CHUNK_SIZE = 1024  # :470
GC_ASSOC_BLOCK_SIZE = 256  # :473
GC_K_INSERT = 64  # :476
GC_K_INSERT_TILE = GC_K_INSERT  # :477

# Bin family.  Both were moved to a never-committed archive/legacy_common/constants_legacy.py
# (constants.py:4-5).  48 is recoverable from prose (CHANGELOG.md:492); tau is not, so every entry
# point takes tau explicitly and 0.1 is only the harness default.
GC_B_BINS = 48
GC_TAU_SOFT_ASSIGN = 0.1

# evidence fusion (pipeline steps 9-11; fl/common/constants.py and PipelineConfig, fl/backend/pipeline.py:104-125)
GC_EXC_EPS = 1e-12  # constants.py:75
GC_ALPHA_MIN = 1.0  # :89
GC_ALPHA_MAX = 1.0  # :90
GC_KAPPA_SCALE = 1.0  # :91
GC_C0_COND = 1e6  # :92
GC_POWER_BETA_MIN = 0.25  # pipeline.py:119
GC_POWER_BETA_EXC_C = 50.0  # pipeline.py:120
GC_POWER_BETA_Z_C = 1.0  # pipeline.py:121
GC_D_Z = 22
