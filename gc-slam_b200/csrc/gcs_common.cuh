// gcs_common.cuh -- context, error handling and small-matrix float64 math shared by all kernels.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/gcs_b200.h"

#define GCS_VERSION_MAJOR 0
#define GCS_VERSION_MINOR 1
#define GCS_VERSION_PATCH 0

struct gcs_ctx {
  int device;
  int sm_count;
  void* ws;  // grow-only device workspace
  uint64_t ws_bytes;
  void* ws_retired[32];  // outgrown blocks: freed in gcs_destroy / gcs_reserve_workspace, never while streams may use them
  int n_retired;
  int ws_frozen;         // gcs_workspace_freeze: growth inside a call is an error (steady state never allocates)
  uint64_t launches;
  // optional CUDA-event timing of the dominant kernel (gcs_timing_*): pairs recorded on the launching stream
  int timing_on;   // 0 off, 1 every bracketed kernel, 100 + tag: only the kernel with that tag (GCS_TIME_*)
  int timing_n;
  cudaEvent_t timing_ev[2 * 256];
  // side stream of the fused primitive entry (the map view is prepared while the scan's surfels are extracted): created
  // on first use, with a workspace of its own (the two streams' scratch must not alias)
  cudaStream_t side_stream;
  cudaEvent_t ev_fork, ev_join;
  void* ws_side;
  uint64_t ws_side_bytes;
  int route_side;   // gcs_side_route: map maintenance calls (recency inflate in place, map update) run on the side stream
  char err[512];
};

// Bracket the dominant kernel of a path with events when timing is enabled (no-ops otherwise).
static inline bool gcs_timing_wants(gcs_ctx* ctx, int tag) {
  return ctx->timing_n < 256 && (ctx->timing_on == 1 || ctx->timing_on == 100 + tag);
}
static inline void gcs_timing_begin(gcs_ctx* ctx, cudaStream_t st, int tag) {
  if (gcs_timing_wants(ctx, tag)) cudaEventRecord(ctx->timing_ev[2 * ctx->timing_n], st);
}
static inline void gcs_timing_end(gcs_ctx* ctx, cudaStream_t st, int tag) {
  if (gcs_timing_wants(ctx, tag)) { cudaEventRecord(ctx->timing_ev[2 * ctx->timing_n + 1], st); ctx->timing_n++; }
}

int gcs_set_error(gcs_ctx* ctx, int code, const char* fmt, ...);
int gcs_ws_reserve(gcs_ctx* ctx, uint64_t bytes);
int gcs_side_reserve(gcs_ctx* ctx, uint64_t bytes);   // side stream + events (first call) and its workspace (grow-only)
// Stream and scratch of a map-maintenance call: the caller's stream and the context workspace, or -- while gcs_side_route is
// on -- the side stream (made to wait for everything enqueued on the caller's stream so far) and the side workspace.
int gcs_maint_stream(gcs_ctx* ctx, cudaStream_t caller, uint64_t ws_bytes, cudaStream_t* st, char** ws);

#define GCS_CHECK_CUDA(ctx, expr)                                                                 \
  do {                                                                                            \
    cudaError_t _e = (expr);                                                                      \
    if (_e != cudaSuccess)                                                                        \
      return gcs_set_error((ctx), GCS_ECUDA, "%s:%d %s -> %s", __FILE__, __LINE__, #expr,         \
                           cudaGetErrorString(_e));                                               \
  } while (0)

#define GCS_REQUIRE(ctx, cond, ...)                                      \
  do {                                                                   \
    if (!(cond)) return gcs_set_error((ctx), GCS_EINVAL, __VA_ARGS__);   \
  } while (0)

// Opt a kernel in to `bytes` of dynamic shared memory.  Function attributes belong to the device's context, so a process
// that drives several GPUs (one gcs_ctx per device, possibly one thread each) must set them on each: the table keeps the
// largest size set so far PER (kernel, device) and is guarded by a mutex -- contexts are per thread, this table is per
// process.  Defined once in gcs_context.cu.
cudaError_t gcs_smem_attr_once(const void* kern, int bytes);

#define GCS_LAUNCH_CHECK(ctx)                        \
  do {                                               \
    (ctx)->launches++;                               \
    GCS_CHECK_CUDA((ctx), cudaGetLastError());       \
  } while (0)

// ------------------------------------------------------------------------------------------------
// Device math.  All float64; operation order follows the reference where it matters for parity.
// ------------------------------------------------------------------------------------------------
namespace gcs {

constexpr double kSmallAngle = 1e-7;  // fl/common/geometry/se3_jax.py:30
constexpr double kNearPi = 1e-7;      // :34
constexpr double kWeightFloor = 1e-12;      // fl/common/constants.py:237
constexpr double kTimeWarpSigmaFrac = 0.1;  // :141
constexpr double kEpsR = 1e-6;              // :73
constexpr double kKappaR0 = 0.8;            // :95
constexpr double kKappaTau = 0.03;          // :96
constexpr double kF64Eps = 2.220446049250313e-16;

// Unit axis of the batched primitive path: unit u (hypothesis) of a stacked (n_units, N_total, ...) structure
__device__ __forceinline__ gcs_meas_batch meas_batch_unit(gcs_meas_batch B, int64_t u) {
  const int64_t o = u * (B.n_feat + B.n_surfel);
  B.Lambdas += 9 * o; B.thetas += 3 * o; B.etas += 9 * o; B.weights += o; B.sources += o; B.source_indices += o;
  B.valid += o; B.timestamps += o; B.colors += 3 * o;
  return B;
}
__device__ __forceinline__ gcs_assoc_result assoc_result_unit(gcs_assoc_result R, int64_t u, int N, int K) {
  const int64_t o = u * N;
  R.responsibilities += o * K; R.candidate_pool_indices += o * K; R.candidate_tile_ids += o * K; R.candidate_slots += o * K;
  R.row_masses += o; R.cost_matrix += o * K;
  return R;
}

struct Mat3 {
  double m[9];  // row-major
  __host__ __device__ double& operator()(int r, int c) { return m[3 * r + c]; }
  __host__ __device__ double operator()(int r, int c) const { return m[3 * r + c]; }
};

__host__ __device__ inline Mat3 mat3_identity() {
  Mat3 a;
  for (int i = 0; i < 9; ++i) a.m[i] = 0.0;
  a.m[0] = a.m[4] = a.m[8] = 1.0;
  return a;
}
__host__ __device__ inline Mat3 mat3_mul(const Mat3& a, const Mat3& b) {
  Mat3 c;
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) {
      double s = 0.0;
      for (int k = 0; k < 3; ++k) s += a(i, k) * b(k, j);
      c(i, j) = s;
    }
  return c;
}
__host__ __device__ inline Mat3 mat3_T(const Mat3& a) {
  Mat3 c;
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) c(i, j) = a(j, i);
  return c;
}
__host__ __device__ inline void mat3_vec(const Mat3& a, const double* v, double* o) {
  for (int i = 0; i < 3; ++i) o[i] = a(i, 0) * v[0] + a(i, 1) * v[1] + a(i, 2) * v[2];
}
__host__ __device__ inline double mat3_det(const Mat3& a) {
  return a(0, 0) * (a(1, 1) * a(2, 2) - a(1, 2) * a(2, 1)) - a(0, 1) * (a(1, 0) * a(2, 2) - a(1, 2) * a(2, 0)) +
         a(0, 2) * (a(1, 0) * a(2, 1) - a(1, 1) * a(2, 0));
}
// General 3x3 inverse by cofactors (jnp.linalg.inv call sites: matrix_fisher_evidence.py:470, pipeline.py:1251).
__host__ __device__ inline Mat3 mat3_inv(const Mat3& a) {
  Mat3 c;
  double c00 = a(1, 1) * a(2, 2) - a(1, 2) * a(2, 1);
  double c01 = a(1, 2) * a(2, 0) - a(1, 0) * a(2, 2);
  double c02 = a(1, 0) * a(2, 1) - a(1, 1) * a(2, 0);
  double det = a(0, 0) * c00 + a(0, 1) * c01 + a(0, 2) * c02;
  double id = 1.0 / det;
  c(0, 0) = c00 * id;
  c(0, 1) = (a(0, 2) * a(2, 1) - a(0, 1) * a(2, 2)) * id;
  c(0, 2) = (a(0, 1) * a(1, 2) - a(0, 2) * a(1, 1)) * id;
  c(1, 0) = c01 * id;
  c(1, 1) = (a(0, 0) * a(2, 2) - a(0, 2) * a(2, 0)) * id;
  c(1, 2) = (a(0, 2) * a(1, 0) - a(0, 0) * a(1, 2)) * id;
  c(2, 0) = c02 * id;
  c(2, 1) = (a(0, 1) * a(2, 0) - a(0, 0) * a(2, 1)) * id;
  c(2, 2) = (a(0, 0) * a(1, 1) - a(0, 1) * a(1, 0)) * id;
  return c;
}
// Solve A x = b (3x3) by Gaussian elimination with partial pivoting (jnp.linalg.solve call sites).
__host__ __device__ inline void mat3_solve(const Mat3& A, const double* b, double* x) {
  double M[3][4];
  for (int i = 0; i < 3; ++i) {
    for (int j = 0; j < 3; ++j) M[i][j] = A(i, j);
    M[i][3] = b[i];
  }
  for (int c = 0; c < 3; ++c) {
    int piv = c;
    double best = fabs(M[c][c]);
    for (int r = c + 1; r < 3; ++r)
      if (fabs(M[r][c]) > best) { best = fabs(M[r][c]); piv = r; }
    if (piv != c)
      for (int j = 0; j < 4; ++j) { double tmp = M[c][j]; M[c][j] = M[piv][j]; M[piv][j] = tmp; }
    double inv = 1.0 / M[c][c];
    for (int r = c + 1; r < 3; ++r) {
      double f = M[r][c] * inv;
      for (int j = c; j < 4; ++j) M[r][j] -= f * M[c][j];
    }
  }
  for (int i = 2; i >= 0; --i) {
    double s = M[i][3];
    for (int j = i + 1; j < 3; ++j) s -= M[i][j] * x[j];
    x[i] = s / M[i][i];
  }
}

// ---- division-free float64 helpers (MUFU seed + Newton; ~1 ulp, no denormal/branchy slow path) -------------------
// x must be finite and > 0 (or +inf: returns 0).
__device__ __forceinline__ double fast_rcp(double x) {
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
  double e = fma(-x, r, 1.0);
  r = fma(r, e, r);
  e = fma(-x, r, 1.0);
  return fma(r, e, r);
}
// 1/sqrt(x) for finite x > 0
__device__ __forceinline__ double fast_rsqrt(double x) {
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  const double hx = 0.5 * x;
  double e = fma(-hx * y, y, 0.5);
  y = fma(y, e, y);
  e = fma(-hx * y, y, 0.5);
  return fma(y, e, y);
}
// One Jacobi rotation annihilating a_pq of a symmetric 3x3; (arp, arq) are the two entries coupling the third
// index r to p and q; vp / vq are the eigenvector columns p and q.  Everything stays in registers.
__host__ __device__ inline void jacobi_rot(double& app, double& aqq, double& apq, double& arp, double& arq, double* vp,
                                           double* vq) {
  if (apq == 0.0) return;
#ifdef __CUDA_ARCH__
  // device: reciprocal / reciprocal-sqrt by MUFU seed + Newton (the IEEE division and sqrt sequences are ~3x longer
  // dependent chains, and this rotation sits on the serial path of every 3x3 eigen-decomposition of the epilogue)
  if (fabs(apq) < 1e-290) return;
  const double theta = (aqq - app) * (0.5 * fast_rcp(fabs(apq))) * (apq < 0.0 ? -1.0 : 1.0);
  const double th1 = fma(theta, theta, 1.0);
  double t, c;
  if (th1 < 1e300) {
    t = (theta >= 0.0 ? 1.0 : -1.0) * fast_rcp(fabs(theta) + th1 * fast_rsqrt(th1));
    c = fast_rsqrt(fma(t, t, 1.0));
  } else {   // |theta| > 1e150: the rotation angle is below double precision
    t = 0.5 / theta;
    c = 1.0;
  }
#else
  const double theta = (aqq - app) / (2.0 * apq);
  const double t = (theta >= 0.0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
  const double c = 1.0 / sqrt(t * t + 1.0);
#endif
  const double sn = t * c;
  app -= t * apq;
  aqq += t * apq;
  apq = 0.0;
  const double rp = arp, rq = arq;
  arp = c * rp - sn * rq;
  arq = sn * rp + c * rq;
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    const double a = vp[k], b = vq[k];
    vp[k] = c * a - sn * b;
    vq[k] = sn * a + c * b;
  }
}

#define GCS_CSWAP3(da, db, va, vb)                                                     \
  if ((da) > (db)) {                                                                   \
    double _t = (da); (da) = (db); (db) = _t;                                          \
    for (int _k = 0; _k < 3; ++_k) { _t = (va)[_k]; (va)[_k] = (vb)[_k]; (vb)[_k] = _t; } \
  }

// Cyclic Jacobi eigen-decomposition of a symmetric 3x3.  Eigenvalues ascending in w[], eigenvectors in the
// columns of V (same order).  Replaces jnp.linalg.eigh (fl/common/primitives.py:101 and others).
__host__ __device__ inline void eigh3(const Mat3& Ain, double* w, Mat3& V) {
  double a00 = Ain(0, 0), a11 = Ain(1, 1), a22 = Ain(2, 2);
  double a01 = 0.5 * (Ain(0, 1) + Ain(1, 0)), a02 = 0.5 * (Ain(0, 2) + Ain(2, 0)), a12 = 0.5 * (Ain(1, 2) + Ain(2, 1));
  double v0[3] = {1, 0, 0}, v1[3] = {0, 1, 0}, v2[3] = {0, 0, 1};
  for (int sweep = 0; sweep < 16; ++sweep) {
    const double off = fabs(a01) + fabs(a02) + fabs(a12);
    const double diag = fabs(a00) + fabs(a11) + fabs(a22);
    if (off <= 1e-300 || off <= 1e-19 * diag) break;
    jacobi_rot(a00, a11, a01, a02, a12, v0, v1);  // (p,q,r) = (0,1,2)
    jacobi_rot(a00, a22, a02, a01, a12, v0, v2);  // (0,2,1)
    jacobi_rot(a11, a22, a12, a01, a02, v1, v2);  // (1,2,0)
  }
  GCS_CSWAP3(a00, a11, v0, v1);
  GCS_CSWAP3(a11, a22, v1, v2);
  GCS_CSWAP3(a00, a11, v0, v1);
  w[0] = a00; w[1] = a11; w[2] = a22;
#pragma unroll
  for (int r = 0; r < 3; ++r) { V(r, 0) = v0[r]; V(r, 1) = v1[r]; V(r, 2) = v2[r]; }
}

// DomainProjectionPSD (fl/common/primitives.py:80-123): symmetrise, eigh, clamp >= eps, rebuild.
// cert6 = [projection_delta, sym_delta, eig_min, eig_max, cond, near_null_count].
__host__ __device__ inline Mat3 psd_project3(const Mat3& M, double eps_psd, double* cert6) {
  Mat3 S;
  double sym2 = 0.0;
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) {
      S(i, j) = 0.5 * (M(i, j) + M(j, i));
      double d = S(i, j) - M(i, j);
      sym2 += d * d;
    }
  double w[3];
  Mat3 V;
  eigh3(S, w, V);
  double wc[3];
  for (int k = 0; k < 3; ++k) wc[k] = w[k] > eps_psd ? w[k] : eps_psd;
  Mat3 P;
  double d2 = 0.0;
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) {
      double s = 0.0;
      for (int k = 0; k < 3; ++k) s += V(i, k) * wc[k] * V(j, k);
      P(i, j) = s;
      double d = s - S(i, j);
      d2 += d * d;
    }
  if (cert6) {
    double mn = fmin(wc[0], fmin(wc[1], wc[2])), mx = fmax(wc[0], fmax(wc[1], wc[2]));
    cert6[0] = sqrt(d2);
    cert6[1] = sqrt(sym2);
    cert6[2] = mn;
    cert6[3] = mx;
    cert6[4] = mx / mn;
    cert6[5] = (double)((wc[0] < 10.0 * eps_psd) + (wc[1] < 10.0 * eps_psd) + (wc[2] < 10.0 * eps_psd));
  }
  return P;
}

// One-sided (Hestenes) rotation making columns gp, gq of G = H V orthogonal.
__host__ __device__ inline double hestenes_rot(double* gp, double* gq, double* vp, double* vq) {
  const double alpha = gp[0] * gp[0] + gp[1] * gp[1] + gp[2] * gp[2];
  const double beta = gq[0] * gq[0] + gq[1] * gq[1] + gq[2] * gq[2];
  const double gamma = gp[0] * gq[0] + gp[1] * gq[1] + gp[2] * gq[2];
  if (gamma == 0.0) return 0.0;
  const double denom = sqrt(alpha * beta);
  const double rel = denom > 0.0 ? fabs(gamma) / denom : 0.0;
  if (rel < 1e-17) return rel;
#ifdef __CUDA_ARCH__
  if (fabs(gamma) < 1e-290) return 0.0;
  double t, c;
  {
    const double zeta = (beta - alpha) * (0.5 * fast_rcp(fabs(gamma))) * (gamma < 0.0 ? -1.0 : 1.0);
    const double z1 = fma(zeta, zeta, 1.0);
    if (z1 < 1e300) {
      t = (zeta >= 0.0 ? 1.0 : -1.0) * fast_rcp(fabs(zeta) + z1 * fast_rsqrt(z1));
      c = fast_rsqrt(fma(t, t, 1.0));
    } else {
      t = 0.5 / zeta;
      c = 1.0;
    }
  }
  const double sn = c * t;
#else
  const double zeta = (beta - alpha) / (2.0 * gamma);
  const double t = (zeta >= 0.0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
  const double c = 1.0 / sqrt(1.0 + t * t), sn = c * t;
#endif
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    const double a = gp[k], b = gq[k];
    gp[k] = c * a - sn * b;
    gq[k] = sn * a + c * b;
    const double x = vp[k], y = vq[k];
    vp[k] = c * x - sn * y;
    vq[k] = sn * x + c * y;
  }
  return rel;
}

#define GCS_CSWAP_DESC(na, nb, ga, gb, va, vb)                                         \
  if ((na) < (nb)) {                                                                   \
    double _t = (na); (na) = (nb); (nb) = _t;                                          \
    for (int _k = 0; _k < 3; ++_k) {                                                   \
      _t = (ga)[_k]; (ga)[_k] = (gb)[_k]; (gb)[_k] = _t;                               \
      _t = (va)[_k]; (va)[_k] = (vb)[_k]; (vb)[_k] = _t;                               \
    }                                                                                  \
  }

// 3x3 SVD  H = U diag(s) V^T, s descending, by one-sided Jacobi on the columns of H V.
// Replaces jnp.linalg.svd (matrix_fisher_evidence.py:215, visual_pose_evidence.py:223).  Null directions of U
// are completed by cross products so that U is always orthonormal.
__host__ __device__ inline void svd3(const Mat3& H, Mat3& U, double* s, Mat3& V) {
  double g0[3] = {H(0, 0), H(1, 0), H(2, 0)}, g1[3] = {H(0, 1), H(1, 1), H(2, 1)}, g2[3] = {H(0, 2), H(1, 2), H(2, 2)};
  double v0[3] = {1, 0, 0}, v1[3] = {0, 1, 0}, v2[3] = {0, 0, 1};
  for (int sweep = 0; sweep < 24; ++sweep) {
    double m = hestenes_rot(g0, g1, v0, v1);
    m = fmax(m, hestenes_rot(g0, g2, v0, v2));
    m = fmax(m, hestenes_rot(g1, g2, v1, v2));
    if (m < 1e-16) break;
  }
  double n0 = sqrt(g0[0] * g0[0] + g0[1] * g0[1] + g0[2] * g0[2]);
  double n1 = sqrt(g1[0] * g1[0] + g1[1] * g1[1] + g1[2] * g1[2]);
  double n2 = sqrt(g2[0] * g2[0] + g2[1] * g2[1] + g2[2] * g2[2]);
  GCS_CSWAP_DESC(n0, n1, g0, g1, v0, v1);
  GCS_CSWAP_DESC(n1, n2, g1, g2, v1, v2);
  GCS_CSWAP_DESC(n0, n1, g0, g1, v0, v1);
  s[0] = n0; s[1] = n1; s[2] = n2;
  const double tiny = fmax(1e-300, 1e-15 * n0);
  double u0[3], u1[3], u2[3];
  const bool ok0 = n0 > tiny, ok1 = n1 > tiny, ok2 = n2 > tiny;
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    u0[k] = ok0 ? g0[k] / n0 : (k == 0 ? 1.0 : 0.0);
    u1[k] = ok1 ? g1[k] / n1 : 0.0;
    u2[k] = ok2 ? g2[k] / n2 : 0.0;
  }
  if (!ok1) {
    // any unit vector orthogonal to u0: remove u0 from the axis it is least aligned with
    const double a0 = fabs(u0[0]), a1 = fabs(u0[1]), a2 = fabs(u0[2]);
    double e[3] = {0, 0, 0};
    if (a0 <= a1 && a0 <= a2) e[0] = 1.0; else if (a1 <= a2) e[1] = 1.0; else e[2] = 1.0;
    const double d = u0[0] * e[0] + u0[1] * e[1] + u0[2] * e[2];
    double y[3] = {e[0] - d * u0[0], e[1] - d * u0[1], e[2] - d * u0[2]};
    const double ny = sqrt(y[0] * y[0] + y[1] * y[1] + y[2] * y[2]);
    for (int k = 0; k < 3; ++k) u1[k] = y[k] / ny;
  }
  if (!ok2) {
    u2[0] = u0[1] * u1[2] - u0[2] * u1[1];
    u2[1] = u0[2] * u1[0] - u0[0] * u1[2];
    u2[2] = u0[0] * u1[1] - u0[1] * u1[0];
  }
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    U(r, 0) = u0[r]; U(r, 1) = u1[r]; U(r, 2) = u2[r];
    V(r, 0) = v0[r]; V(r, 1) = v1[r]; V(r, 2) = v2[r];
  }
}

// skew-symmetric K and K^2 exactly as the reference forms them (K @ K with the zero terms dropped).
__device__ inline void skew_and_square(const double* p, Mat3& K, Mat3& K2) {
  K(0, 0) = 0.0;  K(0, 1) = -p[2]; K(0, 2) = p[1];
  K(1, 0) = p[2]; K(1, 1) = 0.0;   K(1, 2) = -p[0];
  K(2, 0) = -p[1]; K(2, 1) = p[0]; K(2, 2) = 0.0;
  K2 = mat3_mul(K, K);
}

// so3_exp (se3_jax.py:259-301)
__device__ inline Mat3 so3_exp(const double* omega) {
  double theta_sq = omega[0] * omega[0] + omega[1] * omega[1] + omega[2] * omega[2];
  double theta = sqrt(theta_sq);
  bool small = theta < kSmallAngle;
  double st = small ? 1.0 : theta;
  double stsq = (theta_sq < kSmallAngle * kSmallAngle) ? 1.0 : theta_sq;
  double sn, cs;
  sincos(st, &sn, &cs);
  double sin_coeff = small ? 1.0 : sn / st;
  double cos_coeff = small ? 0.5 : (1.0 - cs) / stsq;
  Mat3 K, K2, R;
  skew_and_square(omega, K, K2);
  for (int i = 0; i < 9; ++i) R.m[i] = ((i % 4 == 0) ? 1.0 : 0.0) + sin_coeff * K.m[i] + cos_coeff * K2.m[i];
  return R;
}

__device__ inline void softmax3(const double* x, double* w) {
  double m = fmax(x[0], fmax(x[1], x[2]));
  double e0 = exp(x[0] - m), e1 = exp(x[1] - m), e2 = exp(x[2] - m);
  double s = e0 + e1 + e2;
  w[0] = e0 / s; w[1] = e1 / s; w[2] = e2 / s;
}

// so3_log with the reference's softmax-blended near-pi axis (se3_jax.py:304-366)
__device__ inline void so3_log(const Mat3& R, double* omega) {
  double cos_theta = 0.5 * ((R(0, 0) + R(1, 1) + R(2, 2)) - 1.0);
  cos_theta = fmin(1.0, fmax(-1.0, cos_theta));
  double theta = acos(cos_theta);
  double vex[3] = {0.5 * (R(2, 1) - R(1, 2)), 0.5 * (R(0, 2) - R(2, 0)), 0.5 * (R(1, 0) - R(0, 1))};
  double sin_theta = sin(theta);
  double safe_sin = fabs(sin_theta) < kSmallAngle ? 1.0 : sin_theta;
  double f = theta / (2.0 * safe_sin);
  if (theta < kSmallAngle) {
    omega[0] = vex[0]; omega[1] = vex[1]; omega[2] = vex[2];
  } else if (fabs(theta - 3.141592653589793) < kNearPi) {
    double x[3] = {50.0 * (R(0, 0) + 1.0), 50.0 * (R(1, 1) + 1.0), 50.0 * (R(2, 2) + 1.0)};
    double w[3];
    softmax3(x, w);
    double ax[3];
    for (int r = 0; r < 3; ++r)
      ax[r] = w[0] * (R(r, 0) + (r == 0)) + w[1] * (R(r, 1) + (r == 1)) + w[2] * (R(r, 2) + (r == 2));
    double n = sqrt(ax[0] * ax[0] + ax[1] * ax[1] + ax[2] * ax[2]);
    double sn = n < kSmallAngle ? 1.0 : n;
    for (int r = 0; r < 3; ++r) omega[r] = ax[r] / sn * theta;
  } else {
    for (int r = 0; r < 3; ++r) omega[r] = f * (2.0 * vex[r]);
  }
}

__device__ inline double sigmoid(double x) { return 1.0 / (1.0 + exp(-x)); }

// kappa blend (fl/backend/operators/kappa.py:130-169)
__device__ inline double kappa_from_resultant(double Rbar, double eps_r, double d, double r0, double tau) {
  double R = fmin(fmax(Rbar, 0.0), 1.0 - eps_r);
  double R2 = R * R;
  double k_low = (R * (d - R2)) / (1.0 - R2 + eps_r);
  double k_high = -log(fmax(1.0 - R2, eps_r));
  double s = sigmoid((R - r0) / fmax(tau, 1e-6));
  return (1.0 - s) * k_low + s * k_high;
}

// Per-point constant-twist deskew  p0 = R(a*phi)^T (p - V(a*phi) a*rho)
// (deskew_constant_twist.py:48-56 with se3_exp se3_jax.py:473-504 and so3_exp :259-301).
// For |theta| < 0.5 rad (every real LiDAR sweep) the three Rodrigues coefficients sin(t)/t, (1-cos t)/t^2 and
// (t - sin t)/t^3 are evaluated from their Maclaurin series in theta^2 (8 terms, truncation < 5e-20): no sqrt, no
// sincos, no division, and no cancellation -- the reference's closed forms lose ~1e-16/theta^2 relative accuracy at
// small theta, so the two agree to ~1e-16 * |p| absolute.  Larger angles (e.g. the garbage alpha of zero-stamped
// padded rows under epoch time) take the reference's closed forms with sincos.
__device__ inline void deskew_point(const double* p, double alpha, const double* xi, double* p0) {
  const double rho[3] = {alpha * xi[0], alpha * xi[1], alpha * xi[2]};
  const double phi[3] = {alpha * xi[3], alpha * xi[4], alpha * xi[5]};
  const double th2 = phi[0] * phi[0] + phi[1] * phi[1] + phi[2] * phi[2];
  double Bc, Cc, sin_coeff, cos_coeff;
  if (th2 < 0.25) {
    double s = 2.8114572543455206e-15, c = 1.5619206968586225e-16, g = 8.22063524662433e-18;
    s = fma(s, th2, -7.647163731819816e-13); c = fma(c, th2, -4.779477332387385e-14); g = fma(g, th2, -2.8114572543455206e-15);
    s = fma(s, th2, 1.6059043836821613e-10); c = fma(c, th2, 1.1470745597729725e-11); g = fma(g, th2, 7.647163731819816e-13);
    s = fma(s, th2, -2.505210838544172e-08); c = fma(c, th2, -2.08767569878681e-09);  g = fma(g, th2, -1.6059043836821613e-10);
    s = fma(s, th2, 2.7557319223985893e-06); c = fma(c, th2, 2.755731922398589e-07);  g = fma(g, th2, 2.505210838544172e-08);
    s = fma(s, th2, -0.0001984126984126984); c = fma(c, th2, -2.48015873015873e-05);  g = fma(g, th2, -2.7557319223985893e-06);
    s = fma(s, th2, 0.008333333333333333);   c = fma(c, th2, 0.001388888888888889);   g = fma(g, th2, 0.0001984126984126984);
    s = fma(s, th2, -0.16666666666666666);   c = fma(c, th2, -0.041666666666666664);  g = fma(g, th2, -0.008333333333333333);
    s = fma(s, th2, 1.0);                    c = fma(c, th2, 0.5);                    g = fma(g, th2, 0.16666666666666666);
    sin_coeff = s; cos_coeff = c; Bc = c; Cc = g;
  } else {
    const double theta = sqrt(th2);
    double sn, cs;
    sincos(theta, &sn, &cs);
    Bc = (1.0 - cs) / th2;
    Cc = (theta - sn) / (th2 * theta);
    sin_coeff = sn / theta;
    cos_coeff = Bc;
  }
  // K = [phi]x, K2 = K K = phi phi^T - th2 I  (entries written out; same sums the reference's K @ K forms)
  const double x = phi[0], y = phi[1], z = phi[2];
  const double k2_00 = -(z * z) - y * y, k2_11 = -(z * z) - x * x, k2_22 = -(y * y) - x * x;
  const double k2_01 = x * y, k2_02 = x * z, k2_12 = y * z;
  // t = (I + B K + C K2) rho
  const double t0 = rho[0] + Bc * (-z * rho[1] + y * rho[2]) + Cc * (k2_00 * rho[0] + k2_01 * rho[1] + k2_02 * rho[2]);
  const double t1 = rho[1] + Bc * (z * rho[0] - x * rho[2]) + Cc * (k2_01 * rho[0] + k2_11 * rho[1] + k2_12 * rho[2]);
  const double t2 = rho[2] + Bc * (-y * rho[0] + x * rho[1]) + Cc * (k2_02 * rho[0] + k2_12 * rho[1] + k2_22 * rho[2]);
  const double q0 = p[0] - t0, q1 = p[1] - t1, q2 = p[2] - t2;
  // p0 = R^T q,  R = I + s K + c K2  =>  R^T = I - s K + c K2
  p0[0] = q0 - sin_coeff * (-z * q1 + y * q2) + cos_coeff * (k2_00 * q0 + k2_01 * q1 + k2_02 * q2);
  p0[1] = q1 - sin_coeff * (z * q0 - x * q2) + cos_coeff * (k2_01 * q0 + k2_11 * q1 + k2_12 * q2);
  p0[2] = q2 - sin_coeff * (-y * q0 + x * q1) + cos_coeff * (k2_02 * q0 + k2_12 * q1 + k2_22 * q2);
}

// 2^t for t <= 0 (clamped at -1000), branch-free: round-to-nearest split t = k + r with the 2^52+2^51 trick (r exact,
// |r| <= 0.5), degree-13 Taylor of 2^r in Horner form (truncation 4e-18), exponent add on the high word.
// Straight-line code so that eight of these interleave in the soft-assign loop (libdevice exp() has range-check
// branches that serialise them).
__device__ __forceinline__ double exp2_nonpos(double t) {
  t = fmax(t, -1000.0);
  const double magic = 6755399441055744.0;
  const double tm = t + magic;
  const int k = __double2loint(tm);
  const double r = t - (tm - magic);
  double q = 1.3691488853904128e-12;
  q = fma(q, r, 2.5678435993488206e-11);
  q = fma(q, r, 4.4455382718708116e-10);
  q = fma(q, r, 7.054911620801123e-09);
  q = fma(q, r, 1.01780860092397e-07);
  q = fma(q, r, 1.321548679014431e-06);
  q = fma(q, r, 1.5252733804059841e-05);
  q = fma(q, r, 0.0001540353039338161);
  q = fma(q, r, 0.0013333558146428443);
  q = fma(q, r, 0.009618129107628477);
  q = fma(q, r, 0.05550410866482158);
  q = fma(q, r, 0.24022650695910072);
  q = fma(q, r, 0.6931471805599453);
  q = fma(q, r, 1.0);
  return __hiloint2double(__double2hiint(q) + (k << 20), __double2loint(q));
}

// Eight independent 2^t evaluations with the Horner steps interleaved by hand (the compiler will not reorder them
// across the shared-memory stores that separate consecutive bins).
__device__ __forceinline__ void exp2_nonpos_x8(const double (&t)[8], double (&e)[8]) {
  const double magic = 6755399441055744.0;
  double r[8], q[8];
  int k[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const double tc = fmax(t[j], -1000.0);
    const double tm = tc + magic;
    k[j] = __double2loint(tm);
    r[j] = tc - (tm - magic);
    q[j] = fma(1.3691488853904128e-12, r[j], 2.5678435993488206e-11);
  }
  const double c[12] = {4.4455382718708116e-10, 7.054911620801123e-09, 1.01780860092397e-07, 1.321548679014431e-06,
                        1.5252733804059841e-05, 0.0001540353039338161, 0.0013333558146428443, 0.009618129107628477,
                        0.05550410866482158, 0.24022650695910072, 0.6931471805599453, 1.0};
#pragma unroll
  for (int n = 0; n < 12; ++n)
#pragma unroll
    for (int j = 0; j < 8; ++j) q[j] = fma(q[j], r[j], c[n]);
#pragma unroll
  for (int j = 0; j < 8; ++j) e[j] = __hiloint2double(__double2hiint(q[j]) + (k[j] << 20), __double2loint(q[j]));
}

// float32 -> float64 of a non-negative, flushed-to-zero float by integer ops (keeps the fp64 pipe free of F2F).
__device__ __forceinline__ double f32_bits_to_f64(float f) {
  const unsigned b = __float_as_uint(f);
  const unsigned hi = b ? (b >> 3) + 0x38000000u : 0u;
  return __hiloint2double((int)hi, (int)(b << 29));
}

// smooth_window_weights (fl/backend/operators/imu_preintegration.py:19-43) with sigma = 0.1 * max(t1-t0,1e-12):
//   w = sigmoid(a) sigmoid(b) (1-floor) + floor,  sigmoid(a) sigmoid(b) = 1 / ((1 + e^-a)(1 + e^-b))   (one division).
// inv_sig = 1 / max(0.1 * max(t1 - t0, 1e-12), 1e-6) is hoisted by the caller.
__device__ inline double window_weight(double t, double t0, double t1, double inv_sig) {
  const double a = (t - t0) * inv_sig, b = (t1 - t) * inv_sig;
  const double den = (1.0 + exp(-a)) * (1.0 + exp(-b));
  return (1.0 / den) * (1.0 - kWeightFloor) + kWeightFloor;
}
__device__ inline double window_inv_sigma(double t0, double t1) {
  return 1.0 / fmax(kTimeWarpSigmaFrac * fmax(t1 - t0, 1e-12), 1e-6);
}

// 2^t for |t| <= 1000 (clamped): same polynomial as exp2_nonpos
__device__ __forceinline__ double exp2_bounded(double t) {
  t = fmin(fmax(t, -1000.0), 1000.0);
  const double magic = 6755399441055744.0;
  const double tm = t + magic;
  const int k = __double2loint(tm);
  const double r = t - (tm - magic);
  double q = 1.3691488853904128e-12;
  q = fma(q, r, 2.5678435993488206e-11);
  q = fma(q, r, 4.4455382718708116e-10);
  q = fma(q, r, 7.054911620801123e-09);
  q = fma(q, r, 1.01780860092397e-07);
  q = fma(q, r, 1.321548679014431e-06);
  q = fma(q, r, 1.5252733804059841e-05);
  q = fma(q, r, 0.0001540353039338161);
  q = fma(q, r, 0.0013333558146428443);
  q = fma(q, r, 0.009618129107628477);
  q = fma(q, r, 0.05550410866482158);
  q = fma(q, r, 0.24022650695910072);
  q = fma(q, r, 0.6931471805599453);
  q = fma(q, r, 1.0);
  return __hiloint2double(__double2hiint(q) + (k << 20), __double2loint(q));
}

// ---- constant-twist deskew with the per-scan invariants hoisted ---------------------------------------------------
// xi = (rho, phi).  For a point at sweep fraction a:  R = Exp(a phi), t = V(a phi) a rho,  p0 = R^T (p - t), with
//   K = [phi]x:   t  = a (rho + a (C c1 + a G c2)),  c1 = phi x rho,  c2 = K^2 rho = phi (phi.rho) - |phi|^2 rho
//                 p0 = q - S a (phi x q) + C a^2 (phi (phi.q) - |phi|^2 q),   q = p - t
// S, C, G = sin(th)/th, (1-cos th)/th^2, (th - sin th)/th^3 at th = a |phi| (Maclaurin series in th^2).
// Same algebra as deskew_point (se3_jax.py:473-504, deskew_constant_twist.py:48-56); ~60 flops per point instead of ~105.
struct TwistCtx {
  double rho[3], phi[3], c1[3], c2[3], th1sq;
  const double* xi;
};
__device__ inline TwistCtx make_twist_ctx(const double* xi) {
  TwistCtx c;
  c.xi = xi;
#pragma unroll
  for (int k = 0; k < 3; ++k) { c.rho[k] = xi[k]; c.phi[k] = xi[3 + k]; }
  c.th1sq = c.phi[0] * c.phi[0] + c.phi[1] * c.phi[1] + c.phi[2] * c.phi[2];
  c.c1[0] = c.phi[1] * c.rho[2] - c.phi[2] * c.rho[1];
  c.c1[1] = c.phi[2] * c.rho[0] - c.phi[0] * c.rho[2];
  c.c1[2] = c.phi[0] * c.rho[1] - c.phi[1] * c.rho[0];
  const double pr = c.phi[0] * c.rho[0] + c.phi[1] * c.rho[1] + c.phi[2] * c.rho[2];
#pragma unroll
  for (int k = 0; k < 3; ++k) c.c2[k] = c.phi[k] * pr - c.th1sq * c.rho[k];
  return c;
}
__device__ __forceinline__ void deskew_point_ctx(const double* p, double a, const TwistCtx& c, double* p0) {
  const double a2 = a * a;
  const double th2 = a2 * c.th1sq;
  if (!(th2 < 0.25)) {   // large angles (garbage sweep fractions of zero-stamped padded rows): reference closed forms
    deskew_point(p, a, c.xi, p0);
    return;
  }
  double S, C, G;
  if (th2 < 0.01) {      // truncation < th2^5 / 11! < 3e-18
    S = -2.505210838544172e-08;              C = -2.08767569878681e-09;              G = -1.6059043836821613e-10;
    S = fma(S, th2, 2.7557319223985893e-06); C = fma(C, th2, 2.755731922398589e-07); G = fma(G, th2, 2.505210838544172e-08);
  } else {
    S = 2.8114572543455206e-15;              C = 1.5619206968586225e-16;              G = 8.22063524662433e-18;
    S = fma(S, th2, -7.647163731819816e-13); C = fma(C, th2, -4.779477332387385e-14); G = fma(G, th2, -2.8114572543455206e-15);
    S = fma(S, th2, 1.6059043836821613e-10); C = fma(C, th2, 1.1470745597729725e-11); G = fma(G, th2, 7.647163731819816e-13);
    S = fma(S, th2, -2.505210838544172e-08); C = fma(C, th2, -2.08767569878681e-09);  G = fma(G, th2, -1.6059043836821613e-10);
    S = fma(S, th2, 2.7557319223985893e-06); C = fma(C, th2, 2.755731922398589e-07);  G = fma(G, th2, 2.505210838544172e-08);
  }
  S = fma(S, th2, -0.0001984126984126984); C = fma(C, th2, -2.48015873015873e-05);  G = fma(G, th2, -2.7557319223985893e-06);
  S = fma(S, th2, 0.008333333333333333);   C = fma(C, th2, 0.001388888888888889);   G = fma(G, th2, 0.0001984126984126984);
  S = fma(S, th2, -0.16666666666666666);   C = fma(C, th2, -0.041666666666666664);  G = fma(G, th2, -0.008333333333333333);
  S = fma(S, th2, 1.0);                    C = fma(C, th2, 0.5);                    G = fma(G, th2, 0.16666666666666666);
  const double aG = a * G;
  double q[3];
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    const double u = fma(aG, c.c2[k], C * c.c1[k]);
    q[k] = fma(-a, fma(a, u, c.rho[k]), p[k]);
  }
  const double x0 = c.phi[1] * q[2] - c.phi[2] * q[1];
  const double x1 = c.phi[2] * q[0] - c.phi[0] * q[2];
  const double x2 = c.phi[0] * q[1] - c.phi[1] * q[0];
  const double pq = fma(c.phi[0], q[0], fma(c.phi[1], q[1], c.phi[2] * q[2]));
  const double sa = S * a, ca = C * a2;
  p0[0] = fma(ca, fma(c.phi[0], pq, -c.th1sq * q[0]), fma(-sa, x0, q[0]));
  p0[1] = fma(ca, fma(c.phi[1], pq, -c.th1sq * q[1]), fma(-sa, x1, q[1]));
  p0[2] = fma(ca, fma(c.phi[2], pq, -c.th1sq * q[2]), fma(-sa, x2, q[2]));
}

// ---- window weight with one exponential and no division -----------------------------------------------------------
// sigmoid(a) sigmoid(b) with a + b = S = (t1 - t0)/sigma constant over the scan:  x = e^-a, c = e^-S
//   1 / ((1 + x)(1 + c/x)) = x / (c + x (1 + c + x)).
struct WindowCtx { double t0, inv_sig, c, one_plus_c; };
__device__ inline WindowCtx make_window_ctx(double t0, double t1) {
  WindowCtx w;
  w.t0 = t0;
  w.inv_sig = window_inv_sigma(t0, t1);
  w.c = exp(-(t1 - t0) * w.inv_sig);
  w.one_plus_c = 1.0 + w.c;
  return w;
}
__device__ __forceinline__ double window_weight_ctx(double t, const WindowCtx& w) {
  const double a = (t - w.t0) * w.inv_sig;
  const double x = exp2_bounded(fmin(a * -1.4426950408889634, 500.0));   // x^2 must stay finite
  const double den = fma(x, w.one_plus_c + x, w.c);
  return fma(x * fast_rcp(den), 1.0 - kWeightFloor, kWeightFloor);
}

__device__ inline double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ inline double warp_max(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

}  // namespace gcs
