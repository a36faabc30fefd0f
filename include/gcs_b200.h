/*
 * gcs_b200.h -- C ABI of libgcs_b200.so: GC-SLAM v2's per-scan LiDAR evidence path on NVIDIA B200 (sm_100a).
 *
 * This is the drop-in boundary.  The reference (whabacivch/GC-SLAM) is Python + JAX; the functions below are
 * what a Python binding for this path binds (ctypes).  Each entry point names the reference
 * operator it replaces (paths relative to the reference root; fl/ = fl_ws/src/fl_slam_poc/fl_slam_poc/).
 *
 * Conventions
 *   - plain C types only; every pointer marked (dev) is a device pointer owned by the caller, every pointer
 *     marked (host) is ordinary host memory read during the call.  The library never keeps a pointer after the
 *     stream work it enqueued has finished.
 *   - `stream` is a cudaStream_t passed as void*; all work is enqueued on it, nothing synchronises the host.
 *   - all floating arrays are IEEE float64, C-contiguous; ring/tag are uint8; indices int32/int64 as stated.
 *   - return value: GCS_OK (0) or a negative gcs_status; the message is available from gcs_last_error().
 *     The Python layer maps GCS_EINVAL -> ValueError and everything else -> RuntimeError (the reference is
 *     fail-fast: fl/backend/pipeline.py:824-833, archive/legacy_operators/binning.py:253-264).
 *   - certificate scalars are written to a caller-provided device (or device-visible pinned) array of doubles,
 *     valid once the caller has synchronised the stream; index enums below give the layout.
 *   - a gcs_ctx is bound to one device and is not re-entrant: one ctx per calling thread/stream.
 *   - there is no CPU fallback anywhere in this library.
 */
#ifndef GCS_B200_H
#define GCS_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct gcs_ctx gcs_ctx;

typedef enum {
  GCS_OK = 0,
  GCS_EINVAL = -1, /* shape / contract violation */
  GCS_ECUDA = -2,  /* CUDA runtime error */
  GCS_ENOMEM = -3, /* workspace allocation failed */
  GCS_ECOMM = -4   /* peer / IPC exchange failure */
} gcs_status;

/* ---- library / context -------------------------------------------------------------------------------- */
int gcs_version(void);                  /* MAJOR*10000 + MINOR*100 + PATCH */
const char* gcs_version_string(void);   /* "gcs_sm100a <semver>" : used for RuntimeManifest.backends ids   */
int gcs_create(gcs_ctx** out, int device);
int gcs_destroy(gcs_ctx* ctx);
const char* gcs_last_error(gcs_ctx* ctx); /* ctx may be NULL: returns the last create() failure             */
/* Workspace: gcs_create allocates 32 MB (a single scan of the reference budgets through either family); larger batches
 * grow it on first use (doubling; the outgrown block is retired, not freed, so nothing synchronises a live stream).
 * gcs_reserve_workspace sizes it up front (and releases retired blocks: call it outside the steady state);
 * gcs_workspace_freeze(ctx, 1) turns any later in-call growth into GCS_ENOMEM -- the steady state then provably never
 * allocates.  gcs_workspace_bytes reports the current size (run the workload once, read it, reserve that next time).  */
int gcs_reserve_workspace(gcs_ctx* ctx, uint64_t bytes);
int gcs_workspace_freeze(gcs_ctx* ctx, int frozen);
uint64_t gcs_workspace_bytes(gcs_ctx* ctx);

/* Side stream of the context.  gcs_lidar_evidence_primitives_batched prepares the map view on it while the scan's surfels
 * are extracted.  While gcs_side_route is on, gcs_map_recency_inflate and gcs_map_update run there too (after everything
 * enqueued on the caller's stream so far), so that hypothesis 0's map update (pipeline.py:1233-1447) overlaps the
 * scan-side work of the remaining hypotheses; gcs_side_join makes `stream` wait for everything enqueued on the side
 * stream (call it before reading what those calls wrote).  */
int gcs_side_route(gcs_ctx* ctx, int on);
int gcs_side_join(gcs_ctx* ctx, void* stream);
int gcs_device_sm_count(gcs_ctx* ctx);
uint64_t gcs_kernel_launches(gcs_ctx* ctx); /* number of kernels this ctx has launched so far              */
/* Measurement hook: when enabled, the dominant kernels of each path are bracketed by CUDA events on the launching
 * stream (up to 256 launches between collects).  on = 1: every bracketed kernel; on = 100 + GCS_TIME_x: only that
 * kernel; 0: off.  collect() waits for the events and returns the summed device time and the number of launches,
 * then resets.                                                                                                */
enum { GCS_TIME_BIN_SCAN = 0, GCS_TIME_SURFEL_FIT, GCS_TIME_MAP_VIEW, GCS_TIME_TOPK, GCS_TIME_FUSE, GCS_TIME_INFLATE,
       GCS_TIME_SINKHORN };
int gcs_timing_enable(gcs_ctx* ctx, int on);
int gcs_timing_collect(gcs_ctx* ctx, double* total_ms, int* count);

/* Softmax / moment arithmetic of the fused bin kernel. */
typedef enum {
  GCS_PREC_F64 = 0,   /* everything float64 (matches the reference dtype; default)                          */
  GCS_PREC_MIXED = 1, /* exp() of the soft-assign via f32 MUFU ex2 with an f64 residual correction,         */
                      /* moments accumulated in f64; responsibilities within 3e-7 relative of the f64 path  */
  GCS_PREC_TC = 2     /* MIXED + the 48x19 moment contraction on tcgen05 (3xTF32 split, f64 chunk flushes)   */
} gcs_precision;

/* ---- (8f-1) PointCloud2 ingest : fl/backend/backend_node.py:377-468 (parse_pointcloud2_vlp16) and :1682-1684 --
 * Wire payload of n_msgs messages of n_points points each (same layout, back to back) -> SoA arrays in the base
 * frame.  Field datatypes are the sensor_msgs/PointField codes (1 INT8 .. 6 UINT32, 7 FLOAT32, 8 FLOAT64).
 * Non-finite coordinates become +-1e6 (GC_NONFINITE_SENTINEL); ring is taken modulo 256; the per-point time is
 * multiplied by 1e-9 for a message in which any value exceeds 1e6 (the reference's unit switch); without a time field
 * every point gets header_stamp[m]; weights are the range sigmoid window; tag = 0.  `data` must be readable up to
 * the next multiple of 16 bytes past its end.  x, y, z or ring missing (offset < 0) -> GCS_EINVAL.              */
typedef struct {
  int32_t point_step;
  int32_t off_x, type_x, off_y, type_y, off_z, type_z;
  int32_t off_ring, type_ring;
  int32_t off_time, type_time; /* off_time < 0: no per-point time field ("t" or "time") */
} gcs_pc2_layout;
enum { GCS_PC_N_NONFINITE = 0, GCS_PC_TIME_RESCALED, GCS_PC_NCERT = 4 };
int gcs_parse_pointcloud2_vlp16(gcs_ctx* ctx, void* stream, const uint8_t* data /*dev*/, int n_msgs, int64_t n_points,
                                const gcs_pc2_layout* layout /*host*/, const double* header_stamp /*dev (n_msgs) or NULL*/,
                                const double* R_base_lidar /*host [9] row-major or NULL*/,
                                const double* t_base_lidar /*host [3] or NULL*/, double* pts /*dev (n_msgs,n,3)*/,
                                double* t /*dev (n_msgs,n)*/, double* w /*dev (n_msgs,n)*/, uint8_t* ring /*dev*/,
                                uint8_t* tag /*dev*/, double* cert /*dev (n_msgs, GCS_PC_NCERT)*/);

/* ---- (8f-2) IMU scan twist : fl/backend/operators/imu_preintegration.py:19-146 (smooth_window_weights,
 * preintegrate_imu_relative_pose_jax), fl/common/geometry/se3_jax.py:178-256 (se3_log), glue fl/backend/pipeline.py:436-483.
 * One launch serves n_hyp hypotheses that share the IMU buffer (stamps may be zero-padded, as the pipeline pads to
 * GC_MAX_IMU_PREINT_LEN = 512) and differ in the belief-derived parameters.  params row (GCS_IMU_NPARAM doubles):
 * rotvec_start_WB[3], gyro_bias[3], accel_bias[3], gravity_W[3], sigma, scan_start_time, scan_end_time, trans_scale
 * (0 for rotation-only deskew, else 1).  out row: the reference's return tuple at the GCS_IMU_* offsets plus
 * xi_body = se3_log(delta_pose) with its translation scaled.  xi_out (n_hyp, 6), if given, receives xi_body
 * contiguously -- the layout gcs_bins_args.xi and gcs_deskew_constant_twist read.  gyro == accel == NULL: weights only. */
enum { GCS_IMU_NPARAM = 16 };
enum { GCS_IMU_DELTA_POSE = 0, GCS_IMU_DELTA_R = 6, GCS_IMU_DELTA_P = 15, GCS_IMU_DELTA_V = 18, GCS_IMU_ESS = 21,
       GCS_IMU_A_BODY_MEAN = 22, GCS_IMU_A_WORLD_NOG_MEAN = 25, GCS_IMU_A_WORLD_MEAN = 28, GCS_IMU_DT_EFF_SUM = 31,
       GCS_IMU_XI_BODY = 32, GCS_IMU_NOUT = 40 };
int gcs_imu_scan_twist(gcs_ctx* ctx, void* stream, const double* stamps /*dev (M)*/, const double* gyro /*dev (M,3) or NULL*/,
                       const double* accel /*dev (M,3) or NULL*/, int64_t n_samples, const double* params /*dev (n_hyp,16)*/,
                       const double* weights_in /*dev (n_hyp, M) or NULL: use these instead of the smooth window*/,
                       int n_hyp, double* out /*dev (n_hyp, GCS_IMU_NOUT)*/, double* xi_out /*dev (n_hyp,6) or NULL*/,
                       double* weights_out /*dev (n_hyp, M) or NULL*/);

/* ---- (8f-3) hypothesis combine : fl/backend/operators/hypothesis.py:51-115 (_hypothesis_barycenter_core) with
 * domain_projection_psd_core / spd_cholesky_solve_lifted_core (fl/common/primitives.py:80-123, 141-165).
 * K information pairs (L_k (D,D), h_k (D)) + linearisation points + weights -> floored, renormalised weights, their
 * barycenter in information form, PSD projection of the barycenter (symmetrise, eigendecompose, clamp at eps_psd,
 * rebuild), and the spread proxy sum_k w_k |mu_k - mean mu|^2 with mu_k = (L_k + eps_lift I)^-1 h_k.  D <= 32 (the
 * reference state has D_Z = 22).  means_out (K, D) doubles as the scratch the spread needs.                          */
enum { GCS_HB_FLOOR_ADJUSTMENT = 0, GCS_HB_SPREAD_PROXY, GCS_HB_PSD_PROJECTION_DELTA, GCS_HB_PSD_SYM_DELTA,
       GCS_HB_PSD_EIG_MIN, GCS_HB_PSD_EIG_MAX, GCS_HB_PSD_COND, GCS_HB_PSD_NEAR_NULL, GCS_HB_NCERT = 8 };
int gcs_hypothesis_barycenter(gcs_ctx* ctx, void* stream, const double* L_stack /*dev (K,D,D)*/, const double* h_stack /*dev (K,D)*/,
                              const double* z_lin_stack /*dev (K,D) or NULL*/, const double* weights /*dev (K)*/, int n_hyp,
                              int dim, double weight_floor, double eps_psd, double eps_lift, double* L_out /*dev (D,D)*/,
                              double* h_out /*dev (D)*/, double* z_lin_out /*dev (D) or NULL*/,
                              double* weights_norm_out /*dev (K)*/, double* means_out /*dev (K,D)*/,
                              double* cert /*dev (GCS_HB_NCERT)*/);

/* ---- a1 PointBudgetResample : fl/backend/operators/point_budget.py:50-109,117-221 --------------------- */
enum { GCS_RS_MASS_IN = 0, GCS_RS_MASS_SEL, GCS_RS_SUMSQ_SEL, GCS_RS_ESS, GCS_RS_MASS_SCALE, GCS_RS_NCERT = 8 };
int gcs_point_budget_resample(gcs_ctx* ctx, void* stream,
                              const double* pts /*dev (n_raw,3)*/, const double* t /*dev (n_raw)*/,
                              const double* w /*dev (n_raw)*/, const uint8_t* ring /*dev or NULL*/,
                              const uint8_t* tag /*dev or NULL*/, int64_t n_raw, int64_t cap, double eps_mass,
                              double* out_pts /*dev (cap,3)*/, double* out_t /*dev (cap)*/,
                              double* out_w /*dev (cap)*/, uint8_t* out_ring /*dev (cap)*/,
                              uint8_t* out_tag /*dev (cap)*/, double* cert /*dev [GCS_RS_NCERT]*/);

/* ---- a2 DeskewConstantTwist : fl/backend/operators/deskew_constant_twist.py:31-117 -------------------- */
enum { GCS_DK_SUM_W_OUT = 0, GCS_DK_SUM_W_IN, GCS_DK_NCERT = 4 };
int gcs_deskew_constant_twist(gcs_ctx* ctx, void* stream, const double* pts /*dev (n,3)*/,
                              const double* t /*dev (n)*/, const double* w /*dev (n)*/, int64_t n,
                              const double* xi_body /*host [6] = [rho, phi]*/, double scan_start_time,
                              double scan_end_time, double* out_pts /*dev (n,3)*/, double* out_w /*dev (n)*/,
                              double* cert /*dev [GCS_DK_NCERT]*/);

/* Hypothesis-batched form (the per-hypothesis loop of fl/backend/backend_node.py:2036-2066 around pipeline.py:569-587):
 * n_units twists over the SAME raw points.  xi_body: dev (n_units, 6) -- the layout gcs_imu_scan_twist writes; outputs
 * stacked per unit: out_pts (n_units, n, 3), out_w (n_units, n), cert (n_units, GCS_DK_NCERT).                      */
int gcs_deskew_constant_twist_batched(gcs_ctx* ctx, void* stream, const double* pts /*dev (n,3)*/, const double* t,
                                      const double* w, int64_t n, const double* xi_body /*dev (n_units,6)*/,
                                      int32_t n_units, double scan_start_time, double scan_end_time, double* out_pts,
                                      double* out_w, double* cert);

/* ---- a3 ray directions : fl/backend/pipeline.py:589-593 ----------------------------------------------- */
int gcs_ray_directions(gcs_ctx* ctx, void* stream, const double* pts /*dev (n,3)*/, int64_t n,
                       const double* origin /*host [3]*/, double eps, double* out_dirs /*dev (n,3)*/);

/* ---- a4 BinSoftAssign : archive/legacy_operators/binning.py:56-131 ------------------------------------ */
enum { GCS_SA_ENTROPY_SUM = 0, GCS_SA_MAX_RESP, GCS_SA_NCERT = 4 };
int gcs_bin_soft_assign(gcs_ctx* ctx, void* stream, const double* dirs /*dev (n,3)*/, int64_t n,
                        const double* bin_dirs /*dev (n_bins,3)*/, int n_bins, double tau, double eps_mass,
                        int precision, double* out_resp /*dev (n,n_bins)*/, double* cert /*dev [GCS_SA_NCERT]*/);

/* ---- a5 ScanBinMomentMatch (+a6 kappa) : archive/legacy_operators/binning.py:139-324 ------------------ */
/* Per-unit statistics block, all (dev).  Any output pointer may be NULL (not written). */
typedef struct {
  double* N;         /* (U, B)      */
  double* s_dir;     /* (U, B, 3)   */
  double* S_scatter; /* (U, B, 3,3) */
  double* p_bar;     /* (U, B, 3)   */
  double* Sigma_p;   /* (U, B, 3,3) */
  double* kappa;     /* (U, B)      */
  double* sum_p;     /* (U, B, 3)    raw  sum w r p      (additive; feeds the bin-map update) */
  double* sum_ppT;   /* (U, B, 3,3)  raw  sum w r p p^T                                         */
} gcs_bin_stats;
enum { GCS_ST_ESS = 0, GCS_ST_SUPPORT_FRAC, GCS_ST_PSD_DELTA, GCS_ST_MASS_EPS_RATIO, GCS_ST_NCERT = 8 };
int gcs_scan_bin_moment_match(gcs_ctx* ctx, void* stream, const double* pts /*dev (n,3)*/,
                              const double* point_cov /*dev (n,3,3) or NULL = zeros*/,
                              const double* w /*dev (n)*/, const double* resp /*dev (n,n_bins)*/,
                              const double* point_lambda /*dev (n) or NULL = ones*/,
                              const double* origin /*host [3]*/, int64_t n, int n_bins, double eps_psd,
                              double eps_mass, const gcs_bin_stats* out /*host struct of dev ptrs, U=1*/,
                              double* cert /*dev [GCS_ST_NCERT]*/);

/* ---- a6 KappaFromResultant : fl/backend/operators/kappa.py:130-169 ------------------------------------ */
int gcs_kappa_from_resultant_batch(gcs_ctx* ctx, void* stream, const double* R_bar /*dev (n)*/, int64_t n,
                                   double eps_r, double d, double r0, double tau, double* out_kappa /*dev (n)*/);

/* ---- a9 MapBinStats : archive/bin_atlas.py:83-257 ----------------------------------------------------- */
typedef struct {
  double* S_dir;     /* (B,3)   */
  double* S_scatter; /* (B,3,3) */
  double* N_dir;     /* (B)     */
  double* N_pos;     /* (B)     */
  double* sum_p;     /* (B,3)   */
  double* sum_ppT;   /* (B,3,3) */
} gcs_map_bin_stats;
/* map <- forgetting * (map + pushforward(scan stats, R, t))  with t[2] forced to 0 when planar_z != 0.
 * (update_map_stats + apply_forgetting: archive/bin_atlas.py:137-257; rigid pushforward with the start-of-scan
 * pose: CHANGELOG.md:575-578,684-721 -- the operator itself, PoseCovInflationPushforward, was deleted upstream.) */
int gcs_map_bin_update(gcs_ctx* ctx, void* stream, const gcs_map_bin_stats* map /*host struct, dev ptrs, in/out*/,
                       const double* scan_N, const double* scan_s_dir, const double* scan_S_scatter,
                       const double* scan_sum_p, const double* scan_sum_ppT /*all dev, one unit*/, int n_bins,
                       const double* pose6 /*host [t, rotvec]*/, int planar_z, double forgetting);
/* derived stats: mu_dir (B,3), kappa (B), centroid (B,3), Sigma_c (B,3,3)  (archive/bin_atlas.py:159-221) */
int gcs_map_bin_derived(gcs_ctx* ctx, void* stream, const gcs_map_bin_stats* map, int n_bins, double eps_mass,
                        double eps_psd, double* mu_dir, double* kappa, double* centroid, double* Sigma_c);

/* ---- a7/a8 MatrixFisherRotation + PlanarTranslationEvidence + 22-D embed ------------------------------ */
/* archive/legacy_operators/matrix_fisher_evidence.py:83-394, :413-671, :729-756.  One record per unit.     */
enum {
  GCS_EV_R_MF = 0,          /* 9  */
  GCS_EV_L_ROT = 9,         /* 9  (PSD-projected) */
  GCS_EV_H_ROT = 18,        /* 3  */
  GCS_EV_DELTA_ROT = 21,    /* 3  */
  GCS_EV_SVD_S = 24,        /* 3  */
  GCS_EV_SCAN_METRICS = 27, /* 17: eig desc(3), eigvec cols(9), linearity, planarity, sphericity, anisotropy, eff_rank */
  GCS_EV_MAP_METRICS = 44,  /* 17 */
  GCS_EV_T_WLS = 61,        /* 3  */
  GCS_EV_L_TRANS = 64,      /* 9  (PSD-projected) */
  GCS_EV_H_TRANS = 73,      /* 3  */
  GCS_EV_DELTA_TRANS = 76,  /* 3  */
  GCS_EV_XY_INFO = 79,
  GCS_EV_Z_INFO = 80,
  GCS_EV_Z_SCALE = 81,
  /* certificate scalars */
  GCS_EV_MF_EIG_MIN = 82, GCS_EV_MF_EIG_MAX, GCS_EV_MF_COND, GCS_EV_MF_NEAR_NULL, GCS_EV_MF_NLL_PER_ESS,
  GCS_EV_MF_DIR_SCORE, GCS_EV_MF_PSD_DELTA, GCS_EV_MF_MASS_EPS, GCS_EV_MF_ROT_NLL, GCS_EV_MF_N_EFF,
  GCS_EV_PT_EIG_MIN = 92, GCS_EV_PT_EIG_MAX, GCS_EV_PT_COND, GCS_EV_PT_NEAR_NULL, GCS_EV_PT_NLL_PER_ESS,
  GCS_EV_PT_PSD_DELTA, GCS_EV_PT_MASS_EPS, GCS_EV_PT_TRANS_NLL, GCS_EV_PT_N_EFF,
  GCS_EV_NREC = 104
};
/* Evidence from already-computed scan statistics (U units) against one shared map.  poses: dev (U,6) [t,rotvec];
 * out_evidence: dev (U, GCS_EV_NREC); out_L22: dev (U,22,22) or NULL; out_h22: dev (U,22) or NULL.           */
int gcs_bin_evidence(gcs_ctx* ctx, void* stream, const gcs_bin_stats* scan /*host struct, dev ptrs*/, int n_units,
                     int n_bins, const gcs_map_bin_stats* map, const double* poses, double eps_psd,
                     double eps_mass, double* out_evidence, double* out_L22, double* out_h22);

/* Stand-alone operators with the reference's argument lists.  pose6 / R_hat / t_pred are (host).  Only the
 * MatrixFisher (resp. planar-translation) fields of the GCS_EV_* record are written; the rest are zero.      */
int gcs_matrix_fisher_rotation(gcs_ctx* ctx, void* stream, const double* scan_s_dir, const double* scan_S_scatter,
                               const double* scan_N, const double* map_S_dir, const double* map_S_scatter,
                               const double* map_N_dir, int n_bins, const double* pose6 /*host [t,rotvec]*/,
                               double eps_psd, double eps_mass, double* out_evidence /*dev [GCS_EV_NREC]*/);
int gcs_planar_translation(gcs_ctx* ctx, void* stream, const double* scan_p_bar, const double* scan_Sigma_p,
                           const double* scan_N, const double* map_centroid, const double* map_Sigma_c,
                           const double* map_N_pos, const double* map_S_scatter, const double* map_N_dir, int n_bins,
                           const double* R_hat /*host [9]*/, const double* t_pred /*host [3]*/, double eps_psd,
                           double eps_mass, double* out_evidence /*dev [GCS_EV_NREC]*/);

/* ---- fused bin path: README.md:105-120 steps 1,3,4,5,6,7,8 and the LiDAR term of 9 -------------------- */
/* Units: U = n_scans * n_hyp.  Scan s, hypothesis h -> unit u = s*n_hyp + h.  All hypotheses of a scan share
 * the raw scan and differ in twist and predicted pose.                                                    */
enum {
  GCS_BC_RS_MASS_IN = 0, GCS_BC_RS_ESS, GCS_BC_RS_MASS_SCALE, GCS_BC_DK_SUM_W_OUT, GCS_BC_DK_SUM_W_IN,
  GCS_BC_SA_ENTROPY_SUM, GCS_BC_SA_MAX_RESP, GCS_BC_ST_ESS, GCS_BC_ST_SUPPORT_FRAC, GCS_BC_ST_PSD_DELTA,
  GCS_BC_ST_MASS_EPS_RATIO, GCS_BC_NCERT = 16
};
typedef struct {
  /* raw scans (dev) */
  const double* pts;   /* (S, n_raw, 3) */
  const double* t;     /* (S, n_raw)    */
  const double* w;     /* (S, n_raw)    */
  const uint8_t* ring; /* (S, n_raw) or NULL */
  const uint8_t* tag;  /* (S, n_raw) or NULL */
  int64_t n_raw;
  int64_t cap; /* n_points_cap: output rows per scan; stride = max(1, ceil(n_raw/cap)) */
  int32_t n_scans;
  int32_t n_hyp;
  /* per scan / per unit parameters (dev) */
  const double* scan_t0; /* (S) */
  const double* scan_t1; /* (S) */
  const double* xi;      /* (U, 6) body twist over the scan [rho, phi] */
  const double* poses;   /* (U, 6) predicted START-of-scan pose [t, rotvec]; may be NULL if evidence == NULL */
  /* shared (dev) */
  const double* bin_dirs; /* (B,3) */
  int32_t n_bins;
  int32_t precision; /* gcs_precision */
  double origin[3];  /* lidar_origin_base */
  double tau;
  double eps_mass;
  double eps_psd;
  const gcs_map_bin_stats* map; /* host struct of dev ptrs; may be NULL if evidence == NULL */
  /* point-shard support (multi-GPU): this rank holds raw rows [shard_row0, shard_row0 + n_raw) of
   * scans that are n_raw_total long.  Single GPU: shard_row0 = 0, n_raw_total = n_raw.                       */
  int64_t shard_row0;
  int64_t n_raw_total; /* 0 = n_raw */
  int64_t cap_total;   /* 0 = cap; global n_points_cap when `cap` is this rank's share of the output rows      */
  double bin_norm_max; /* max_b |bin_dirs[b]| (1 for the Fibonacci atlas); 0 = 1.  Bounds the softmax shift.    */
  /* optional per-point contract outputs (dev, NULL = not materialised) */
  double* rs_pts;   /* (S, cap_local, 3)   PointBudgetResult */
  double* rs_t;     /* (S, cap_local)      */
  double* rs_w;     /* (S, cap_local)      */
  uint8_t* rs_ring; /* (S, cap_local)      */
  uint8_t* rs_tag;  /* (S, cap_local)      */
  double* dk_pts;   /* (U, cap_local, 3)   DeskewConstantTwistResult */
  double* dk_w;     /* (U, cap_local)      */
  double* resp;     /* (U, cap_local, B)   BinSoftAssignResult ("contract-materialised" mode) */
  /* per-unit outputs (dev) */
  gcs_bin_stats stats;
  double* evidence; /* (U, GCS_EV_NREC) or NULL */
  double* L22;      /* (U, 22, 22) or NULL      */
  double* h22;      /* (U, 22) or NULL          */
  double* cert;     /* (U, GCS_BC_NCERT)        */
} gcs_bins_args;

/* Number of doubles per unit in the raw additive accumulator block used by the two-phase API below.        */
int gcs_bins_raw_sums_len(int n_bins);
/* Whole path in one call (phase 1 + phase 2). */
int gcs_lidar_evidence_bins(gcs_ctx* ctx, void* stream, const gcs_bins_args* args);
/* Phase 1: per-point work; leaves additive raw sums (U, raw_sums_len) in `raw_sums` (dev) and running maxima
 * (U, 2) in `raw_max` (dev).  When points of one scan are sharded over ranks the caller all-reduces raw_sums
 * with SUM and raw_max with MAX (NCCL, ~7.5 KB per unit) before phase 2.  `mass` (dev, (S, 4)): additive
 * resample masses; for sharded clouds run gcs_bins_mass first, all-reduce(SUM) and pass the result here;
 * pass NULL to have them computed locally.                                                                 */
int gcs_bins_mass(gcs_ctx* ctx, void* stream, const gcs_bins_args* args, double* mass /*dev (S,4)*/);
int gcs_bins_accumulate(gcs_ctx* ctx, void* stream, const gcs_bins_args* args, const double* mass,
                        double* raw_sums, double* raw_max);
/* Phase 2: statistics, kappa, PSD projection, Matrix-Fisher rotation, planar translation, 22-D embed, certs. */
int gcs_bins_finalize(gcs_ctx* ctx, void* stream, const gcs_bins_args* args, const double* mass,
                      const double* raw_sums, const double* raw_max);
/* The exchange step of a point-sharded cloud: every rank packs its partial statistics into ONE
 * buffer [additive block of n_sum doubles | running maxima, n_max doubles], the buffers are all-gathered (one collective,
 * ncclAllGather through torch.distributed), and this entry reduces the gathered (world, n_sum + n_max) array in rank
 * order -- SUM for the additive block, MAX for the maxima.  Same data, same order on every rank: bit-identical
 * statistics everywhere.  Used for both exchanges of the path (resample masses; raw sums + maxima).                  */
int gcs_bins_reduce_gathered(gcs_ctx* ctx, void* stream, const double* gathered /*dev (world, n_sum + n_max)*/,
                             int32_t world, int64_t n_sum, int64_t n_max, double* out_sum /*dev (n_sum)*/,
                             double* out_max /*dev (n_max)*/);

/* The same exchange WITHOUT a library collective (SURVEY.md 8b `gcs_allreduce_bin_stats`; ranks = processes of one node,
 * one GPU each): every rank owns a receive window in its HBM, exported through CUDA IPC and mapped by all peers
 * (NVLink / NVSwitch peer access).  gcs_peer_xchg_reduce launches ONE kernel that pushes the rank's packed block
 * [n_sum additive doubles | n_max maxima] into every rank's window, signals with a release store, acquire-polls for all
 * peers (bounded: ~2 s, then the time-out is recorded and reported by gcs_peer_xchg_status), and reduces the window in rank
 * order in place in `pack` -- bit-identical results on every rank.  Every rank must call it the same number of times.
 *   create:  allocates the window for `bytes_per_rank` per block and writes the 64-byte IPC handle to ipc_handle_out (host)
 *   connect: all_handles = the `world` handles in rank order (host, world * 64 bytes; exchange them with any host channel) */
typedef struct gcs_peer_xchg gcs_peer_xchg;
int gcs_peer_xchg_create(gcs_ctx* ctx, int32_t rank, int32_t world, uint64_t bytes_per_rank, gcs_peer_xchg** out,
                         void* ipc_handle_out /*host, 64 bytes*/);
int gcs_peer_xchg_connect(gcs_ctx* ctx, gcs_peer_xchg* x, const void* all_handles /*host, world * 64 bytes*/);
int gcs_peer_xchg_reduce(gcs_ctx* ctx, gcs_peer_xchg* x, void* stream, double* pack /*dev, in place*/, int64_t n_sum,
                         int64_t n_max);
int gcs_peer_xchg_status(gcs_ctx* ctx, gcs_peer_xchg* x, uint32_t* out_epoch /*host: 0 = no time-out*/);
int gcs_peer_xchg_destroy(gcs_ctx* ctx, gcs_peer_xchg* x);


/* ================================================================================================
 * Primitive family (BASELINE config 3): LiDAR surfels -> map view -> OT association -> pose evidence -> map update
 * ================================================================================================ */

/* MeasurementBatch (fl/backend/structures/measurement_batch.py:68-134): fixed (n_feat + n_surfel) rows, camera
 * splats first.  All pointers (dev).  etas are (N_total, 3 lobes, 3).                                        */
typedef struct {
  double* Lambdas;         /* (Nt,3,3) */
  double* thetas;          /* (Nt,3)   */
  double* etas;            /* (Nt,3,3) */
  double* weights;         /* (Nt)     */
  int32_t* sources;        /* (Nt) 0 = camera, 1 = lidar */
  int32_t* source_indices; /* (Nt)     */
  uint8_t* valid;          /* (Nt)     */
  double* timestamps;      /* (Nt)     */
  double* colors;          /* (Nt,3)   */
  int32_t n_feat;
  int32_t n_surfel;
} gcs_meas_batch;

/* measurement_batch_from_camera_splats (measurement_batch.py:174-259): fills rows [0, n) of an all-zero batch. */
int gcs_batch_from_camera_splats(gcs_ctx* ctx, void* stream, const double* positions, const double* covariances,
                                 const double* directions, const double* kappas, const double* weights,
                                 const double* timestamps, const double* colors /*dev (n,3) or NULL = gray*/,
                                 int32_t n, double eps_lift, const gcs_meas_batch* batch);

/* ---- a10 extract_lidar_surfels : fl/backend/operators/lidar_surfel_extraction.py:84-431,
 *      bin_points_3d fl/common/ma_hex_web.py:243-303 --------------------------------------------------------- */
typedef struct {
  int32_t n_cells_1, n_cells_2, n_cells_z, max_occupants, min_points_per_voxel;
  double voxel_size_m, sensor_noise_var_per_axis, wishart_nu, wishart_psi_scale, kappa_main_scale, kappa_min,
      kappa_max, eig_min, eps_lift;
} gcs_surfel_cfg;
/* Writes the first n_valid rows of the LiDAR slice of `batch` (rows beyond stay as the caller left them, exactly
 * as the reference's .at[start:end].set).  out_n_valid: dev int32[1].  Optional debug outputs (dev or NULL):
 * out_bucket (n_cells, max_occ) int32 with -1 = empty, out_count (n_cells) int32 clipped to max_occ.           */
int gcs_extract_lidar_surfels(gcs_ctx* ctx, void* stream, const double* pts /*dev (n,3)*/,
                              const double* timestamps /*dev (n)*/, const double* weights /*dev (n)*/, int64_t n,
                              const gcs_surfel_cfg* cfg /*host*/, const gcs_meas_batch* batch /*host struct*/,
                              int32_t* out_n_valid, int32_t* out_bucket, int32_t* out_count);

/* Hypothesis-batched form: unit u reads pts + u*3n, weights + u*n (timestamps shared when timestamps_shared != 0) and
 * writes unit u of a STACKED measurement batch -- `batch` holds the pointers of unit 0, unit u lies u * (n_feat +
 * n_surfel) rows further in every array.  out_n_valid: dev int32[n_units].                                        */
int gcs_extract_lidar_surfels_batched(gcs_ctx* ctx, void* stream, const double* pts /*dev (n_units,n,3)*/,
                                      const double* timestamps, const double* weights /*dev (n_units,n)*/, int64_t n,
                                      int32_t n_units, int32_t timestamps_shared, const gcs_surfel_cfg* cfg,
                                      const gcs_meas_batch* batch, int32_t* out_n_valid);

/* ---- a11 tiles: PrimitiveMapTile / AtlasMap (fl/backend/structures/primitive_map.py:98-211) ----------------
 * The atlas is a pool of n_tiles_cap tiles of m_tile slots each, one SoA array per field, all (dev).           */
typedef struct {
  double* Lambdas;            /* (T,M,3,3) */
  double* thetas;             /* (T,M,3)   */
  double* etas;               /* (T,M,3,3) */
  double* weights;            /* (T,M)     */
  double* timestamps;         /* (T,M)     */
  double* created_timestamps; /* (T,M)     */
  int64_t* last_supported_scan_seq; /* (T,M) */
  int64_t* last_update_scan_seq;    /* (T,M) */
  int64_t* primitive_ids;     /* (T,M)     */
  uint8_t* valid;             /* (T,M)     */
  double* colors;             /* (T,M,3)   */
  double* cam_mass;           /* (T,M)     */
  double* lidar_mass;         /* (T,M)     */
  double* rgb_cam_accum;      /* (T,M,3)   */
  double* rgb_cam_denom;      /* (T,M)     */
  double* rgb;                /* (T,M,3)   */
  int32_t m_tile;
  int32_t n_tiles_cap;
} gcs_atlas;

/* primitive_map_recency_inflate (primitive_map.py:1400-1484).  tile_index: host int32[n_tiles], pool index of each
 * active tile or -1 if the tile does not exist (skipped).  stats: dev double[4] =
 * [downscale_total, cov_inflation_trace, n_valid_total, -].                                                    */
int gcs_map_recency_inflate(gcs_ctx* ctx, void* stream, const gcs_atlas* atlas, const int32_t* tile_index,
                            int32_t n_tiles, int64_t scan_seq, double recency_decay_lambda, double min_scale,
                            double* stats);

/* AtlasMapView (primitive_map.py:269-300), pool of n_tiles * m_tile_view rows, all (dev). */
typedef struct {
  int64_t* candidate_tile_ids; /* (P) */
  int32_t* candidate_slots;    /* (P) */
  uint8_t* valid;              /* (P) */
  double* positions;           /* (P,3)   */
  double* covariances;         /* (P,3,3) */
  double* directions;          /* (P,3)   */
  double* kappas;              /* (P)     */
  double* weights;             /* (P)     */
  int64_t* primitive_ids;      /* (P)     */
  int64_t* last_supported_scan_seq; /* (P) */
  double* etas;                /* (P,3,3) */
  double* colors;              /* (P,3)   */
} gcs_map_view;
/* extract_atlas_map_view (primitive_map.py:356-450, :303-322, :474-498): per stencil tile the first m_tile_view
 * slots of a stable descending sort by weight (invalid -> -1e30), then mu = solve(Lambda+eps I, theta),
 * Sigma = inv(.), kappa = |sum eta|, dir = sum eta / (kappa + eps).  tile_index[i] = -1: tile missing = empty tile.
 * out_n_valid: dev int32[1] = number of valid pool rows.                                                       */
int gcs_extract_atlas_map_view(gcs_ctx* ctx, void* stream, const gcs_atlas* atlas, const int32_t* tile_index /*host*/,
                               const int64_t* tile_ids /*host*/, int32_t n_tiles, int32_t m_tile_view,
                               double eps_lift, double eps_mass, const gcs_map_view* view, int32_t* out_n_valid);

/* extract_atlas_map_view(primitive_map_recency_inflate(map)) without touching the map: the hypotheses after the first
 * see an inflated COPY of the map in the reference (pipeline.py:835-853; only hypothesis 0's map is kept,
 * backend_node.py:2079-2083).  The gathered Lambda / theta of every valid view entry are scaled by
 * clip(exp(-lambda * max(scan_seq - last_supported_scan_seq, 0)), min_scale, 1) as primitive_map.py:1400-1484 scales
 * them; selection and all other fields do not depend on the inflation.  out_inflate_stats (dev double[4], may be NULL):
 * the statistics gcs_map_recency_inflate reports.                                                                  */
int gcs_extract_atlas_map_view_inflated(gcs_ctx* ctx, void* stream, const gcs_atlas* atlas, const int32_t* tile_index,
                                        const int64_t* tile_ids, int32_t n_tiles, int32_t m_tile_view, double eps_lift,
                                        double eps_mass, int64_t scan_seq, double recency_decay_lambda, double min_scale,
                                        const gcs_map_view* view, int32_t* out_n_valid, double* out_inflate_stats);

/* ---- (8f-4, merge half) primitive_map_merge_reduce : fl/backend/structures/primitive_map.py:1501-2031 ----------
 * All-pairs Bhattacharyya distance of the tile's Gaussians (mu = solve(Lambda + eps_lift I, theta), Sigma = inv(.)),
 * greedy disjoint selection of at most max_pairs pairs in stable ascending order of distance (finite, < threshold),
 * moment-matched merge of each pair into its first slot (second slot: weight 0, invalid), in place.  m_tile <= 2048
 * (the reference's GC_PRIMITIVE_MERGE_MAX_TILE_SIZE: above it the operator is a declared no-op, which the host layer
 * reports without calling this).  stats (dev): number of merged pairs, status (0 nothing to merge, 1 merged).       */
enum { GCS_MR_N_MERGED = 0, GCS_MR_STATUS, GCS_MR_NSTATS = 4 };
int gcs_map_merge_reduce(gcs_ctx* ctx, void* stream, const gcs_atlas* atlas, int32_t tile_index /*pool row*/,
                         double merge_threshold, int32_t max_pairs, double eps_psd, double eps_lift,
                         double* stats /*dev [GCS_MR_NSTATS]*/);

/* ---- (8f-4, export half) map export : extract_primitive_map_view (fl/backend/structures/primitive_map.py:474-576),
 * renderable_batch_from_view (:580-616), PrimitiveMapPublisher.publish + _build_pointcloud2_from_view
 * (fl/backend/map_publisher.py:44-90, 131-258).  Every valid slot of the listed tiles (publishing order = the order of
 * tile_index; -1 = missing tile), moments in the world frame, sorted newest first (last_supported_scan_seq descending,
 * ties by primitive id ascending = np.lexsort((ids, -recency))), and the /gc/map/points payload: 16-byte records
 * x, y, z, intensity (float32 LE), intensity = clip(weight, 0, 1e6).  The per-tile down-selection max_primitives of the
 * reference is its publisher's default (None).  capacity >= n_tiles * m_tile rows in every output array; *out_count
 * (dev) = rows written.  cloud must be 16-byte aligned.                                                             */
typedef struct {
  double* mu_world;                  /* (N,3)   */
  double* Sigma_world;               /* (N,3,3) */
  double* Lambda_world;              /* (N,3,3) inv(Sigma + eps_lift I) */
  double* eta;                       /* (N,3,3) */
  double* mass;                      /* (N)     */
  double* color;                     /* (N,3)   tile.rgb */
  int64_t* primitive_ids;            /* (N)     */
  int64_t* last_supported_scan_seq;  /* (N)     */
  uint8_t* cloud;                    /* (N,16)  */
} gcs_map_export;
int gcs_export_map_points(gcs_ctx* ctx, void* stream, const gcs_atlas* atlas, const int32_t* tile_index /*host*/,
                          int32_t n_tiles, double eps_lift, const gcs_map_export* out, int64_t capacity,
                          int32_t* out_count /*dev int32[1]*/);

/* ---- a12 associate_primitives_ot : fl/backend/operators/primitive_association.py:105-553 -------------------- */
typedef struct {
  int32_t k_assoc, k_sinkhorn, r_stencil_xy, r_stencil_z;   /* k_assoc: 4, 8 (reference default) or 16 */
  int32_t a_policy;   /* measurement marginal (primitive_association.py:412-424): 0 UNIFORM a = valid / sum(valid),
                         1 WEIGHT_PROPORTIONAL a = valid * weight / sum(valid * weight)                              */
  int32_t reserved_;
  double beta, epsilon, tau_a, tau_b, eps_mass, eps_lift, h_tile, recency_decay_lambda;
  int64_t scan_seq;
} gcs_assoc_cfg;
typedef struct {
  double* responsibilities;        /* (Nt,K) */
  int32_t* candidate_pool_indices; /* (Nt,K) */
  int64_t* candidate_tile_ids;     /* (Nt,K) */
  int64_t* candidate_slots;        /* (Nt,K) */
  double* row_masses;              /* (Nt)   */
  double* cost_matrix;             /* (Nt,K) */
} gcs_assoc_result;
enum { GCS_OT_MARGINAL_A = 0, GCS_OT_MARGINAL_B, GCS_OT_MASS_TOTAL, GCS_OT_SUM_A, GCS_OT_SUM_M, GCS_OT_SUM_NOVEL,
       GCS_OT_P95_A, GCS_OT_NONZERO_A, GCS_OT_B_RECENCY_P95, GCS_OT_ESS, GCS_OT_TOTAL_COST, GCS_OT_SUM_M2,
       GCS_OT_NCERT = 16 };
/* view_tile_ids: host int64[n_tiles] (tile list of the view, in view order).  Measurement rows use UNIFORM
 * a-marginal, b = 1/K (the only policies the reference implements).  cert: dev double[GCS_OT_NCERT].          */
int gcs_associate_primitives_ot(gcs_ctx* ctx, void* stream, const gcs_meas_batch* batch, const gcs_map_view* view,
                                const int64_t* view_tile_ids /*host*/, int32_t n_tiles, int32_t m_tile_view,
                                const gcs_assoc_cfg* cfg, const gcs_assoc_result* out, double* cert);

/* Hypothesis-batched form: n_units stacked measurement batches (see gcs_extract_lidar_surfels_batched) against ONE
 * read-only view; `out` holds the pointers of unit 0 of stacked (n_units, Nt, K) results, cert is (n_units, GCS_OT_NCERT).
 * One thread-block cluster runs the Sinkhorn iterations of each unit.                                              */
int gcs_associate_primitives_ot_batched(gcs_ctx* ctx, void* stream, const gcs_meas_batch* batch, int32_t n_units,
                                        const gcs_map_view* view, const int64_t* view_tile_ids /*host*/, int32_t n_tiles,
                                        int32_t m_tile_view, const gcs_assoc_cfg* cfg, const gcs_assoc_result* out,
                                        double* cert);

/* ---- a13 visual_pose_evidence : fl/backend/operators/visual_pose_evidence.py:74-436 ------------------------ */
enum { GCS_VP_L_TRANS = 0 /*9*/, GCS_VP_H_TRANS = 9 /*3*/, GCS_VP_L_ROT = 12 /*9*/, GCS_VP_H_ROT = 21 /*3*/,
       GCS_VP_TRANS_COST = 24, GCS_VP_ROT_COST, GCS_VP_SUM_ROW_MASS, GCS_VP_N_VALID_ROWS, GCS_VP_SVD_S /*3*/ = 28,
       GCS_VP_DELTA_ROT /*3*/ = 31, GCS_VP_R_SCATTER /*9*/ = 34, GCS_VP_NREC = 48 };
int gcs_visual_pose_evidence(gcs_ctx* ctx, void* stream, const gcs_meas_batch* batch, const gcs_map_view* view,
                             const gcs_assoc_result* assoc, int32_t k_assoc, const double* pose6 /*host [t,rotvec]*/,
                             double eps_lift, double eps_mass, double* out_L22 /*dev (22,22)*/,
                             double* out_h22 /*dev (22)*/, double* out_rec /*dev [GCS_VP_NREC]*/);

/* Hypothesis-batched form: poses dev (n_units, 6); outputs stacked (n_units, 22, 22), (n_units, 22), (n_units, NREC).
 * n_lidar_valid (dev int32[n_units]) / view_n_valid (dev int32[1]), if both given: the reference's early exits taken on
 * the device -- a unit with n_camera_valid + n_lidar_valid[u] == 0, or every unit when the view is empty, gets the
 * all-zero association (`assoc` is overwritten) and the evidence eps_lift * I, h = 0.                              */
int gcs_visual_pose_evidence_batched(gcs_ctx* ctx, void* stream, const gcs_meas_batch* batch, int32_t n_units,
                                     const gcs_map_view* view, const gcs_assoc_result* assoc, int32_t k_assoc,
                                     const double* poses /*dev (n_units,6) [t,rotvec]*/, double eps_lift, double eps_mass,
                                     double* out_L22, double* out_h22, double* out_rec,
                                     const int32_t* n_lidar_valid /*dev or NULL*/, int32_t n_camera_valid,
                                     const int32_t* view_n_valid /*dev or NULL*/);

/* ---- the hypothesis loop of one scan in ONE call (fl/backend/backend_node.py:2036-2066 around the primitive-family
 *      steps of process_scan_single_hypothesis, fl/backend/pipeline.py:569-587, 780-877, 998-1010): n_units hypotheses
 *      that share one stencil -- deskew with every twist, surfels of every deskewed cloud, ONE read-only view (recency
 *      inflation applied functionally when `inflate` is set), association, pose evidence.  Runs
 *      gcs_deskew_constant_twist_batched, gcs_extract_lidar_surfels_batched, gcs_extract_atlas_map_view[_inflated],
 *      gcs_associate_primitives_ot_batched, gcs_visual_pose_evidence_batched in this order on `stream`; results are
 *      bit-identical to calling them (or the single-hypothesis entries, hypothesis by hypothesis).  No host
 *      synchronisation, no host -> device copy: the sequence can be captured in a CUDA graph.  All pointers (dev) unless
 *      noted; stacked arrays hold unit u at offset u * (rows of one unit).                                          */
typedef struct {
  const double* pts;      /* (n,3) raw points, shared by all units */
  const double* t;        /* (n)   */
  const double* w;        /* (n)   */
  int64_t n;
  int32_t n_units;
  int32_t inflate;        /* 1: view of the recency-inflated map (map itself untouched), 0: view of the map as it is */
  const double* xi;       /* (n_units,6) twists   */
  const double* poses;    /* (n_units,6) [t, rotvec] predicted poses (linearisation points) */
  double scan_start_time, scan_end_time;
  double* dk_pts;         /* out (n_units,n,3) */
  double* dk_w;           /* out (n_units,n)   */
  double* dk_cert;        /* out (n_units, GCS_DK_NCERT) */
  gcs_surfel_cfg surfel_cfg;
  gcs_meas_batch base;    /* camera slice + zero LiDAR rows, (Nt) rows: copied into every unit before the surfels are written;
                             base.Lambdas == NULL: the stacked batch is taken as the caller prepared it                 */
  gcs_meas_batch batch;   /* out, stacked: unit 0's pointers */
  int32_t* n_lidar_valid; /* out (n_units) */
  int32_t n_camera_valid; /* valid rows of the camera slice (host knowledge of the base batch), for the early exits */
  int32_t reserved_;
  const gcs_atlas* atlas; /* host struct */
  int32_t tile_index[16]; /* pool rows of the stencil tiles (-1: tile missing) */
  int64_t tile_ids[16];
  int32_t n_tiles, m_tile_view;
  double eps_lift, eps_mass, recency_min_scale;
  gcs_map_view view;      /* out (n_tiles * m_tile_view rows) */
  int32_t* view_n_valid;  /* out int32[1] */
  double* inflate_stats;  /* out double[4] (written when inflate != 0) */
  gcs_assoc_cfg assoc_cfg;
  gcs_assoc_result assoc; /* out, stacked */
  double* ot_cert;        /* out (n_units, GCS_OT_NCERT) */
  double* L22;            /* out (n_units,22,22) */
  double* h22;            /* out (n_units,22)    */
  double* rec;            /* out (n_units, GCS_VP_NREC) */
} gcs_prim_batch_args;
int gcs_lidar_evidence_primitives_batched(gcs_ctx* ctx, void* stream, const gcs_prim_batch_args* args /*host*/);

/* ---- a14 map update = pipeline step 12b : fl/backend/pipeline.py:1233-1447 with primitive_map_fuse /
 *      insert_masked / cull / forget (fl/backend/structures/primitive_map.py:807-1384).  This is what survives of
 *      "PoseCovInflationPushforward" (README.md:119).  merge_reduce is not run: it is a no-op for tiles larger than
 *      GC_PRIMITIVE_MERGE_MAX_TILE_SIZE = 2048 (primitive_map.py:1881-1890) and M_TILE is 50,000.               */
typedef struct {
  int32_t k_insert_tile, k_assoc, assoc_block_size, strict_tile_state; /* strict: reproduce the unmasked timestamp
                                                                           stamping of quirk Q7                    */
  double recency_decay_lambda, eps_lift, eps_mass, h_tile, cull_weight_threshold, forgetting_factor;
  int64_t scan_seq, next_global_id;
  double timestamp;
  int64_t* next_global_id_dev; /* optional (NULL: next_global_id above is used): device int64[1] holding the next id; read by
                                  this update and advanced by the ids it assigns, so that a caller can enqueue the next scan's
                                  update before it has read this one's statistics back                                   */
} gcs_map_update_cfg;
enum { GCS_MU_FUSED_COUNT = 0, GCS_MU_FUSED_MASS, GCS_MU_INSERT_COUNT, GCS_MU_INSERT_MASS, GCS_MU_INSERT_MASS_P95,
       GCS_MU_EVICTED_COUNT, GCS_MU_EVICTED_MASS, GCS_MU_NEXT_GLOBAL_ID, GCS_MU_TILE_COUNT0 /* 16 tile counts */ = 8,
       GCS_MU_NSTATS = 24 };
/* tile_index / tile_ids: host arrays of the n_tiles active tiles (all must exist in the pool: the host layer creates
 * empty tiles first, as the reference's fuse/insert do).  out_new_ids: dev int64 (n_tiles, k_insert) (-1 = not
 * inserted); out_insert_slots: dev int32 (n_tiles, k_insert); stats: dev double[GCS_MU_NSTATS].                 */
int gcs_map_update(gcs_ctx* ctx, void* stream, const gcs_atlas* atlas, const int32_t* tile_index,
                   const int64_t* tile_ids, int32_t n_tiles, const gcs_meas_batch* batch,
                   const gcs_assoc_result* assoc, const double* pose6 /*host z_t*/, const gcs_map_update_cfg* cfg,
                   int64_t* out_new_ids, int32_t* out_insert_slots, double* stats);

/* ---- a14, one operator at a time: the per-tile map operators behind gcs_map_update, with the reference's own
 *      granularity (primitive_map_fuse / insert_masked / cull / forget keep their Python signatures).
 *      tile_index = pool row of the tile (the host layer creates the empty tile first, as the reference does).
 *      All array arguments are device pointers.                                                                    */

/* primitive_map_fuse (fl/backend/structures/primitive_map.py:992-1163), one (block, tile) call: Product-of-Experts
 * scatter-add of n proposals into target_slots (int32, proposals with a slot outside [0, m_tile) are dropped as a JAX
 * scatter drops them), r = responsibilities * valid_mask (valid_mask NULL: all valid); camera / lidar masses only when
 * sources_meas is given, camera colour accumulators only when colors_meas is given too; rgb and colors of the whole
 * tile recomputed; timestamps stamped at every named slot, masked or not (:1112); scan-seq stamps where sum r > 0.
 * stats (dev double[4]): [0] n_fused = number of distinct slots.  n <= 16384.                                      */
int gcs_map_fuse(gcs_ctx* ctx, void* stream, const gcs_atlas* atlas, int32_t tile_index, const int32_t* target_slots,
                 const double* Lambdas_meas /*(n,3,3)*/, const double* thetas_meas /*(n,3)*/,
                 const double* etas_meas /*(n,3,3)*/, const double* weights_meas /*(n)*/,
                 const double* responsibilities /*(n)*/, const uint8_t* valid_mask /*(n) or NULL*/,
                 const double* colors_meas /*(n,3) or NULL*/, const int32_t* sources_meas /*(n) or NULL*/, int32_t n,
                 double timestamp, int64_t scan_seq, double eps_mass, double* stats);

/* primitive_map_insert_masked (primitive_map.py:807-981): the k lowest-retention slots of the tile (empty first,
 * retention = weight * exp(-lambda * max(0, scan_seq - last_supported)), stable by slot: _select_lowest_mass_slots_fixed
 * :326-353) receive the proposals whose valid_new_mask is set; ids next_global_id + (running count of set masks).
 * colors_new NULL = zeros, sources_new NULL = all lidar (:884-893).  out_new_ids (k) int64, -1 = not inserted;
 * out_target_slots (k) int32.  stats (dev double[4]): [0] n_inserted, [1] masked-out proposals, [2] valid slots after.
 * k <= min(1024, m_tile).                                                                                          */
int gcs_map_insert_masked(gcs_ctx* ctx, void* stream, const gcs_atlas* atlas, int32_t tile_index,
                          const double* Lambdas_new, const double* thetas_new, const double* etas_new,
                          const double* weights_new, const uint8_t* valid_new_mask, const double* colors_new,
                          const int32_t* sources_new, int32_t k, double timestamp, int64_t scan_seq,
                          double recency_decay_lambda, int64_t next_global_id, int64_t* out_new_ids,
                          int32_t* out_target_slots, double* stats);

/* primitive_map_cull (primitive_map.py:1175-1304): clear valid slots with weight < weight_threshold; with
 * max_primitives >= 0 and more survivors than that, the threshold becomes the weight of rank max_primitives in the
 * descending order of weights * valid (:1226-1232).  max_primitives < 0 = None.
 * stats (dev double[4]): [0] n_culled, [1] mass_dropped, [2] sum of all slot weights (the certificate's denominator),
 * [3] valid slots left.                                                                                            */
enum { GCS_CULL_N = 0, GCS_CULL_MASS, GCS_CULL_SUM_W, GCS_CULL_N_LEFT, GCS_CULL_NSTATS = 4 };
int gcs_map_cull(gcs_ctx* ctx, void* stream, const gcs_atlas* atlas, int32_t tile_index, double weight_threshold,
                 int32_t max_primitives, double* stats);

/* primitive_map_forget (primitive_map.py:1314-1384): weights *= forgetting_factor over every slot of the tile.      */
int gcs_map_forget(gcs_ctx* ctx, void* stream, const gcs_atlas* atlas, int32_t tile_index, double forgetting_factor);

/* ---- (8f-3, fusion half) evidence fusion for all hypotheses of a scan in one launch: pipeline steps 9-11
 *      fl/backend/pipeline.py:1038-1193 (raw evidence, observability sentinels, power tempering beta, pose-block
 *      conditioning), fl/backend/operators/excitation.py:15-64 (Fisher-derived scaling of the prior),
 *      fl/backend/operators/fusion.py:46-230 (fusion scale alpha, InfoFusionAdditive with DomainProjectionPSD,
 *      fl/common/primitives.py:80-123).  Stacks are (K, 22, 22) / (K, 22), all (dev).
 *      cert_scalars (K, 4): the certificate-level inputs of the two control laws, per hypothesis:
 *      support.ess_total, excitation.dt_effect + extrinsic_effect, mismatch.nll_per_ess of the aggregated evidence
 *      certificate (host scalars in the reference as well).  L_other / h_other = the IMU + odometry evidence (NULL: none).
 *      Optional outputs (NULL to skip): tempered evidence, scaled prior.  rec: (K, GCS_FU_NREC).                   */
enum { GCS_FU_SKIP_TEMPERING = 1,       /* beta = 1, evidence taken as given (info_fusion_additive on its own)      */
       GCS_FU_SKIP_PRIOR_SCALING = 2,   /* prior taken as given                                                     */
       GCS_FU_ALPHA_GIVEN = 4 };        /* alpha = cfg.alpha_override instead of fusion_scale_from_certificates     */
enum { GCS_FU_IN_ESS_TOTAL = 0, GCS_FU_IN_EXC_TOTAL, GCS_FU_IN_NLL_PER_ESS, GCS_FU_IN_RESERVED };
typedef struct {
  double power_beta_min, power_beta_z_c, power_beta_exc_c;   /* PipelineConfig, pipeline.py:119-121 */
  double alpha_min, alpha_max, c0_cond;                      /* fusion.py:49-52                     */
  double eps_mass, eps_psd, exc_eps;
  double alpha_override;
  int32_t flags, reserved;
} gcs_fusion_cfg;
enum { GCS_FU_BETA = 0, GCS_FU_DT_ASYMMETRY, GCS_FU_Z_TO_XY, GCS_FU_ESS_TO_EXC, GCS_FU_S_DT, GCS_FU_S_EX,
       GCS_FU_POSE_EIG_MIN, GCS_FU_POSE_EIG_MAX, GCS_FU_POSE_COND, GCS_FU_POSE_NEAR_NULL, GCS_FU_ALPHA, GCS_FU_QUALITY,
       GCS_FU_PSD_PROJECTION_DELTA, GCS_FU_PSD_SYM_DELTA, GCS_FU_POST_EIG_MIN, GCS_FU_POST_EIG_MAX, GCS_FU_POST_COND,
       GCS_FU_POST_NEAR_NULL, GCS_FU_TRACE_INCREASE, GCS_FU_NREC = 24 };
int gcs_evidence_fusion(gcs_ctx* ctx, void* stream, const double* L_lidar, const double* h_lidar, const double* L_other,
                        const double* h_other, const double* L_prior, const double* h_prior, const double* cert_scalars,
                        int n_hyp, int dim, const gcs_fusion_cfg* cfg, double* L_post, double* h_post, double* L_evidence,
                        double* h_evidence, double* L_prior_scaled, double* h_prior_scaled, double* rec);

/* ---- a12, the two inner functions of the association on their own (the reference calls them directly from its
 *      start-up warm-up, fl/backend/backend_node.py:884-905).  All pointers (dev).
 * _compute_sparse_cost_matrix_jax (fl/backend/operators/primitive_association.py:152-197):
 *   out_cost[i,k] = |x_i - x_j|^2 + beta * H^2_vMF(dir_i kappa_i, dir_j kappa_j), j = candidate_indices[i,k]
 *   (negative indices wrap, out-of-range indices clamp, as the gather there).                                       */
int gcs_sparse_cost_matrix(gcs_ctx* ctx, void* stream, const double* meas_positions /*(N,3)*/,
                           const double* meas_directions /*(N,3)*/, const double* meas_kappas /*(N)*/, int32_t n_meas,
                           const double* map_positions /*(M,3)*/, const double* map_directions /*(M,3)*/,
                           const double* map_kappas /*(M)*/, int32_t n_map, const int32_t* candidate_indices /*(N,K)*/,
                           int32_t k_cand, double beta, double eig_min, double* out_cost /*(N,K)*/);
/* _sinkhorn_unbalanced_fixed_k_jax (primitive_association.py:105-138): K = exp(-C / max(epsilon, 1e-12)), n_iters
 * fixed iterations u = (a / (K v + 1e-12))^(1 / (1 + tau_a / eps)), v = (b / (K^T u + 1e-12))^(1 / (1 + tau_b / eps)),
 * out_pi = diag(u) K diag(v).  cost (n_rows, n_cols) row-major, n_cols <= 32.                                         */
int gcs_sinkhorn_unbalanced_fixed_k(gcs_ctx* ctx, void* stream, const double* cost, const double* a, const double* b,
                                    int32_t n_rows, int32_t n_cols, double epsilon, double tau_a, double tau_b,
                                    int32_t n_iters, double* out_pi);

#ifdef __cplusplus
}
#endif
#endif /* GCS_B200_H */
