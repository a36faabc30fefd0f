// gcs_hypothesis.cu -- the step behind the per-hypothesis LiDAR evidence: combine K
// hypotheses' posterior information (L_k, h_k, z_lin_k) into one on the device, so that 64 hypotheses evaluated on the
// GPU(s) need no per-hypothesis host round trip.
//   _hypothesis_barycenter_core      fl/backend/operators/hypothesis.py:51-115
//   domain_projection_psd_core       fl/common/primitives.py:80-123   (D x D, D = 22: symmetrise, eigh, clamp, rebuild)
//   spd_cholesky_solve_lifted_core   fl/common/primitives.py:141-165  (means of the hypotheses for the spread proxy)
//
// One CTA.  The symmetric eigenproblem runs as a parallel-ordered cyclic Jacobi iteration in shared memory: the D/2
// disjoint (p, q) pairs of a round-robin step get their rotations from the current matrix, then all threads apply the
// column rotations (A <- A J, V <- V J) and the row rotations (A <- J^T A); D - 1 steps make a sweep, a fixed number
// of sweeps bounded by 12, ended as soon as the off-diagonal mass drops below 1e-30 of the squared Frobenius norm
// (quadratic convergence: 6-8 sweeps; the count depends on the data only, so reruns stay bit-identical).  The rebuilt
// matrix V max(lambda, eps) V^T does not depend on the eigenvalue order.  The K lifted Cholesky solves run one warp
// per hypothesis (lane = row) in a kernel of their own, eight hypotheses per CTA, before the combine.
#include "gcs_jacobi.cuh"

namespace gcs {

namespace {

constexpr int kHbCholWarps = kHbThreads / 32;
constexpr size_t kHbDynBytes = (size_t)kHbCholWarps * kHbMaxD * kHbLd * sizeof(double);   // per-warp Cholesky buffers

struct HbParams {
  const double* L; const double* h; const double* z; const double* w;
  int K, D;
  double floor, eps_psd, eps_lift;
  double* L_out; double* h_out; double* z_out; double* wn_out; double* means; double* cert;
};

// Means of the hypotheses: (L_k + eps_lift I) mu_k = h_k by Cholesky (:101-105), one warp per hypothesis (lane = row),
// eight hypotheses per CTA -- this part is independent of the barycenter and spreads over the SMs.
__global__ void __launch_bounds__(kHbThreads) hypothesis_means_kernel(const HbParams P) {
  extern __shared__ double chol_buf[];
  const int tid = threadIdx.x, D = P.D, K = P.K;
  const int warp = tid >> 5, lane = tid & 31;
  const int k = blockIdx.x * kHbCholWarps + warp;
  if (k >= K) return;
  double* Cw = chol_buf + (size_t)warp * kHbMaxD * kHbLd;
  const double* Lk = P.L + (int64_t)k * D * D;
  if (lane < D)
    for (int j = 0; j <= lane; ++j)   // jnp.linalg.cholesky symmetrises its input (symmetrize_input=True)
      Cw[lane * kHbLd + j] = 0.5 * (Lk[lane * D + j] + Lk[j * D + lane]) + (j == lane ? P.eps_lift : 0.0);
  __syncwarp();
  // right-looking factorisation: column j, then the rank-one update of the trailing rows (all lanes busy)
  for (int j = 0; j < D; ++j) {
    const double cjj = sqrt(Cw[j * kHbLd + j]);
    __syncwarp();
    double cij = 0.0;
    if (lane == j) Cw[j * kHbLd + j] = cjj;
    if (lane > j && lane < D) { cij = Cw[lane * kHbLd + j] / cjj; Cw[lane * kHbLd + j] = cij; }
    __syncwarp();
    if (lane > j && lane < D)
      for (int m = j + 1; m <= lane; ++m) Cw[lane * kHbLd + m] -= cij * Cw[m * kHbLd + j];
    __syncwarp();
  }
  // C y = h (forward), C^T mu = y (backward), column-oriented: the solved component is broadcast, every lane updates
  double r = lane < D ? P.h[(int64_t)k * D + lane] : 0.0, y = 0.0;
  for (int j = 0; j < D; ++j) {
    const double yj = __shfl_sync(0xffffffffu, r, j) / Cw[j * kHbLd + j];
    if (lane == j) y = yj;
    if (lane > j && lane < D) r -= Cw[lane * kHbLd + j] * yj;
  }
  r = y;
  double x = 0.0;
  for (int j = D - 1; j >= 0; --j) {
    const double xj = __shfl_sync(0xffffffffu, r, j) / Cw[j * kHbLd + j];
    if (lane == j) x = xj;
    if (lane < j) r -= Cw[j * kHbLd + lane] * xj;
  }
  if (lane < D) P.means[(int64_t)k * D + lane] = x;
}

__global__ void __launch_bounds__(kHbThreads) hypothesis_barycenter_kernel(const HbParams P) {
  __shared__ double A[kHbMaxD * kHbLd], V[kHbMaxD * kHbLd], S[kHbMaxD * kHbLd];
  __shared__ JacobiScratch jsc;
  __shared__ double sred[kHbThreads];
  __shared__ double s_wsum, s_floor_adj;
  const int tid = threadIdx.x, D = P.D, K = P.K;

  // ---- weights: floor, renormalise (hypothesis.py:83-88)
  if (tid == 0) {
    double sum = 0.0, adj = 0.0;
    for (int k = 0; k < K; ++k) {
      const double wf = fmax(P.w[k], P.floor);
      adj += fabs(wf - P.w[k]);
      sum += wf;
    }
    s_wsum = sum; s_floor_adj = adj;
  }
  __syncthreads();
  for (int k = tid; k < K; k += kHbThreads) P.wn_out[k] = fmax(P.w[k], P.floor) / s_wsum;
  __syncthreads();   // wn_out is re-read below by other threads of this CTA (global memory, same block: visible after the barrier)

  // ---- barycenter in information form (:90-97), hypotheses summed in index order
  for (int e = tid; e < D * D; e += kHbThreads) {
    double acc = 0.0;
    for (int k = 0; k < K; ++k) acc += P.wn_out[k] * P.L[(int64_t)k * D * D + e];
    A[(e / D) * kHbLd + (e % D)] = acc;
  }
  for (int e = tid; e < D; e += kHbThreads) {
    double ah = 0.0, az = 0.0;
    for (int k = 0; k < K; ++k) {
      ah += P.wn_out[k] * P.h[(int64_t)k * D + e];
      if (P.z) az += P.wn_out[k] * P.z[(int64_t)k * D + e];
    }
    P.h_out[e] = ah;
    if (P.z_out) P.z_out[e] = az;
  }
  __syncthreads();

  // ---- DomainProjectionPSD: symmetrise
  double part = 0.0;
  for (int e = tid; e < D * D; e += kHbThreads) {
    const int i = e / D, j = e % D;
    const double m = 0.5 * (A[i * kHbLd + j] + A[j * kHbLd + i]);
    const double d = m - A[i * kHbLd + j];
    part += d * d;
    S[i * kHbLd + j] = m;
    V[i * kHbLd + j] = (i == j) ? 1.0 : 0.0;
  }
  const double sym_delta = sqrt(hb_block_sum(part, sred));
  for (int e = tid; e < D * D; e += kHbThreads) A[(e / D) * kHbLd + (e % D)] = S[(e / D) * kHbLd + (e % D)];
  __syncthreads();

  // ---- parallel-ordered cyclic Jacobi (gcs_jacobi.cuh)
  cta_jacobi_eigh(A, V, D, jsc, sred);

  // ---- clamp, rebuild, certificate (primitives.py:104-121)
  part = 0.0;
  for (int e = tid; e < D * D; e += kHbThreads) {
    const int i = e / D, j = e % D;
    double acc = 0.0;
    for (int m = 0; m < D; ++m) acc += V[i * kHbLd + m] * fmax(A[m * kHbLd + m], P.eps_psd) * V[j * kHbLd + m];
    P.L_out[e] = acc;
    const double d = acc - S[i * kHbLd + j];
    part += d * d;
  }
  const double pd = hb_block_sum(part, sred);
  if (tid == 0) {
    double emin = 1e300, emax = -1e300, nn = 0.0;
    for (int m = 0; m < D; ++m) {
      const double v = fmax(A[m * kHbLd + m], P.eps_psd);
      emin = fmin(emin, v); emax = fmax(emax, v);
      nn += (v < 10.0 * P.eps_psd) ? 1.0 : 0.0;
    }
    P.cert[GCS_HB_FLOOR_ADJUSTMENT] = s_floor_adj;
    P.cert[GCS_HB_PSD_PROJECTION_DELTA] = sqrt(pd);
    P.cert[GCS_HB_PSD_SYM_DELTA] = sym_delta;
    P.cert[GCS_HB_PSD_EIG_MIN] = emin;
    P.cert[GCS_HB_PSD_EIG_MAX] = emax;
    P.cert[GCS_HB_PSD_COND] = emax / emin;
    P.cert[GCS_HB_PSD_NEAR_NULL] = nn;
  }
  __syncthreads();

  // ---- spread proxy: sum_k w_k |mu_k - sum_j w_j mu_j|^2 (:107-113); means come from hypothesis_means_kernel
  __shared__ double mom[kHbMaxD];
  if (tid < D) {
    double m = 0.0;
    for (int k = 0; k < K; ++k) m += P.wn_out[k] * P.means[(int64_t)k * D + tid];
    mom[tid] = m;
  }
  __syncthreads();
  double sp = 0.0;
  for (int k = tid; k < K; k += kHbThreads) {
    double dsq = 0.0;
    for (int i = 0; i < D; ++i) {
      const double d = P.means[(int64_t)k * D + i] - mom[i];
      dsq += d * d;
    }
    sp += P.wn_out[k] * dsq;
  }
  const double spread = hb_block_sum(sp, sred);
  if (tid == 0) P.cert[GCS_HB_SPREAD_PROXY] = spread;
}

}  // namespace

}  // namespace gcs

using namespace gcs;

extern "C" int gcs_hypothesis_barycenter(gcs_ctx* ctx, void* stream, const double* L_stack, const double* h_stack,
                                         const double* z_lin_stack, const double* weights, int n_hyp, int dim,
                                         double weight_floor, double eps_psd, double eps_lift, double* L_out, double* h_out,
                                         double* z_lin_out, double* weights_norm_out, double* means_out, double* cert) {
  if (!ctx) return GCS_EINVAL;
  GCS_CHECK_CUDA(ctx, cudaSetDevice(ctx->device));
  GCS_REQUIRE(ctx, L_stack && h_stack && weights, "gcs_hypothesis_barycenter: L_stack / h_stack / weights are NULL");
  GCS_REQUIRE(ctx, L_out && h_out && weights_norm_out && means_out && cert, "gcs_hypothesis_barycenter: an output is NULL");
  GCS_REQUIRE(ctx, n_hyp >= 1, "gcs_hypothesis_barycenter: need at least one hypothesis (got %d)", n_hyp);
  GCS_REQUIRE(ctx, dim >= 1 && dim <= kHbMaxD, "gcs_hypothesis_barycenter: dimension %d outside [1, %d]", dim, kHbMaxD);
  GCS_REQUIRE(ctx, (z_lin_stack == nullptr) == (z_lin_out == nullptr),
              "gcs_hypothesis_barycenter: z_lin_stack and z_lin_out must both be set or both NULL");
  HbParams P;
  P.L = L_stack; P.h = h_stack; P.z = z_lin_stack; P.w = weights; P.K = n_hyp; P.D = dim;
  P.floor = weight_floor; P.eps_psd = eps_psd; P.eps_lift = eps_lift;
  P.L_out = L_out; P.h_out = h_out; P.z_out = z_lin_out; P.wn_out = weights_norm_out; P.means = means_out; P.cert = cert;
  GCS_CHECK_CUDA(ctx, gcs_smem_attr_once((const void*)hypothesis_means_kernel, (int)kHbDynBytes));
  hypothesis_means_kernel<<<(n_hyp + kHbCholWarps - 1) / kHbCholWarps, kHbThreads, kHbDynBytes, (cudaStream_t)stream>>>(P);
  GCS_LAUNCH_CHECK(ctx);
  hypothesis_barycenter_kernel<<<1, kHbThreads, 0, (cudaStream_t)stream>>>(P);
  GCS_LAUNCH_CHECK(ctx);
  return GCS_OK;
}
