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
int gcs_reserve_workspace(gcs_ctx* ctx, uint64_t bytes); /* optional: pre-size so later calls never allocate */
int gcs_device_sm_count(gcs_ctx* ctx);
uint64_t gcs_kernel_launches(gcs_ctx* ctx); /* number of kernels this ctx has launched so far              */
/* Measurement hook: when enabled, the dominant kernel of each path (bin_scan_kernel, ...) is bracketed by CUDA
 * events on the launching stream (up to 256 launches between collects).  collect() waits for them and returns
 * the summed device time and the number of launches, then resets.                                           */
int gcs_timing_enable(gcs_ctx* ctx, int on);
int gcs_timing_collect(gcs_ctx* ctx, double* total_ms, int* count);

/* Softmax / moment arithmetic of the fused bin kernel. */
typedef enum {
  GCS_PREC_F64 = 0,   /* everything float64 (matches the reference dtype; default)                          */
  GCS_PREC_MIXED = 1, /* exp() of the soft-assign via f32 MUFU ex2 with an f64 residual correction,         */
                      /* moments accumulated in f64; responsibilities within 3e-7 relative of the f64 path  */
  GCS_PREC_TC = 2     /* MIXED + the 48x19 moment contraction on tcgen05 (3xTF32 split, f64 chunk flushes)   */
} gcs_precision;

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

#ifdef __cplusplus
}
#endif
#endif /* GCS_B200_H */
