// gcs_assoc_ops.cu -- the two inner functions of the OT association as entries of their own.  The reference calls them
// directly from its start-up warm-up (fl/backend/backend_node.py:884-905) and from associate_primitives_ot; the fused
// association (gcs_associate_primitives_ot, gcs_prims_map.cu) keeps its own cluster kernel for the (N, 8) production shape.
//   _compute_sparse_cost_matrix_jax     fl/backend/operators/primitive_association.py:152-197
//   _sinkhorn_unbalanced_fixed_k_jax    fl/backend/operators/primitive_association.py:105-138
#include "gcs_assoc.cuh"

namespace gcs {

namespace {

constexpr int kSkgThreads = 512;   // 128 registers per thread: the per-thread column partials (<= 32) stay in registers
constexpr int kSkgMaxCols = 32;

struct CostArgs {
  const double* mp; const double* md; const double* mk; const double* vp; const double* vd; const double* vk;
  const int32_t* cand;
  int N, K, M;
  double beta, eig_min;
  double* out;
};

__global__ void __launch_bounds__(256) sparse_cost_kernel(const CostArgs a) {
  const int64_t e = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (e >= (int64_t)a.N * a.K) return;
  const int i = (int)(e / a.K);
  int j = a.cand[e];
  j = j < 0 ? j + a.M : j;                       // negative indices wrap, as a jnp gather does
  j = j < 0 ? 0 : (j >= a.M ? a.M - 1 : j);      // out-of-range indices clamp
  const double mk = a.mk[i];
  const double A_k1 = A_vmf(fmax(mk, a.eig_min), a.eig_min);
  a.out[e] = pair_cost(a.mp + 3 * i, a.md + 3 * i, mk, A_k1, a.vp + 3 * (int64_t)j, a.vd + 3 * (int64_t)j, a.vk[j], a.beta, a.eig_min);
}

// One CTA.  Rows are strided over the threads; a thread keeps the partial column sums of its rows in registers, the M
// column totals come from fixed-order block sums (shuffle tree per warp, warps in index order): bit-identical reruns.
__global__ void __launch_bounds__(kSkgThreads) sinkhorn_generic_kernel(const double* __restrict__ Cm, const double* __restrict__ a,
                                                                       const double* __restrict__ b, int N, int M, double epsilon,
                                                                       double tau_a, double tau_b, int iters,
                                                                       double* __restrict__ Kmat, double* __restrict__ u,
                                                                       double* __restrict__ pi) {
  __shared__ double v[kSkgMaxCols], sred[kSkgMaxCols][kSkgThreads / 32];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const double eps = fmax(epsilon, 1e-12);
  const double ua = 1.0 / (1.0 + tau_a / eps), vb = 1.0 / (1.0 + tau_b / eps);
  for (int64_t e = tid; e < (int64_t)N * M; e += kSkgThreads) Kmat[e] = exp(-Cm[e] / eps);
  if (tid < M) v[tid] = 1.0;
  __syncthreads();
  for (int it = 0; it < iters; ++it) {
    double col[kSkgMaxCols];
#pragma unroll
    for (int j = 0; j < kSkgMaxCols; ++j) col[j] = 0.0;
    for (int i = tid; i < N; i += kSkgThreads) {
      const double* kr = Kmat + (int64_t)i * M;
      double kv = 0.0;
      for (int j = 0; j < M; ++j) kv += kr[j] * v[j];
      const double ui = pow_pos(a[i] / (kv + 1e-12), ua);
      u[i] = ui;
#pragma unroll
      for (int j = 0; j < kSkgMaxCols; ++j)
        if (j < M) col[j] += kr[j] * ui;
    }
    __syncthreads();   // every thread has read v
#pragma unroll
    for (int j = 0; j < kSkgMaxCols; ++j) {
      if (j < M) {
        const double r = warp_sum(col[j]);
        if (lane == 0) sred[j][warp] = r;
      }
    }
    __syncthreads();
    if (tid < M) {
      double t = 0.0;
      for (int w = 0; w < kSkgThreads / 32; ++w) t += sred[tid][w];
      v[tid] = pow_pos(b[tid] / (t + 1e-12), vb);
    }
    __syncthreads();
  }
  if (iters == 0)
    for (int i = tid; i < N; i += kSkgThreads) u[i] = 1.0;
  __syncthreads();
  for (int64_t e = tid; e < (int64_t)N * M; e += kSkgThreads) pi[e] = u[e / M] * Kmat[e] * v[e % M];
}

}  // namespace

}  // namespace gcs

using namespace gcs;

extern "C" int gcs_sparse_cost_matrix(gcs_ctx* ctx, void* stream, const double* meas_positions, const double* meas_directions,
                                      const double* meas_kappas, int32_t n_meas, const double* map_positions,
                                      const double* map_directions, const double* map_kappas, int32_t n_map,
                                      const int32_t* candidate_indices, int32_t k_cand, double beta, double eig_min, double* out_cost) {
  if (!ctx) return GCS_EINVAL;
  GCS_CHECK_CUDA(ctx, cudaSetDevice(ctx->device));
  GCS_REQUIRE(ctx, meas_positions && meas_directions && meas_kappas && map_positions && map_directions && map_kappas &&
                       candidate_indices && out_cost, "gcs_sparse_cost_matrix: NULL pointer");
  GCS_REQUIRE(ctx, n_meas >= 1 && n_map >= 1 && k_cand >= 1, "gcs_sparse_cost_matrix: empty shape (%d, %d, %d)", n_meas, n_map, k_cand);
  CostArgs a;
  a.mp = meas_positions; a.md = meas_directions; a.mk = meas_kappas; a.vp = map_positions; a.vd = map_directions; a.vk = map_kappas;
  a.cand = candidate_indices; a.N = n_meas; a.K = k_cand; a.M = n_map; a.beta = beta; a.eig_min = eig_min; a.out = out_cost;
  const int64_t n = (int64_t)n_meas * k_cand;
  sparse_cost_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(a);
  GCS_LAUNCH_CHECK(ctx);
  return GCS_OK;
}

extern "C" int gcs_sinkhorn_unbalanced_fixed_k(gcs_ctx* ctx, void* stream, const double* cost, const double* a, const double* b,
                                               int32_t n_rows, int32_t n_cols, double epsilon, double tau_a, double tau_b,
                                               int32_t n_iters, double* out_pi) {
  if (!ctx) return GCS_EINVAL;
  GCS_CHECK_CUDA(ctx, cudaSetDevice(ctx->device));
  GCS_REQUIRE(ctx, cost && a && b && out_pi, "gcs_sinkhorn_unbalanced_fixed_k: NULL pointer");
  GCS_REQUIRE(ctx, n_rows >= 1 && n_cols >= 1 && n_cols <= kSkgMaxCols,
              "gcs_sinkhorn_unbalanced_fixed_k: shape (%d, %d): need 1 <= columns <= %d (the association's K_ASSOC is 8)", n_rows, n_cols,
              kSkgMaxCols);
  GCS_REQUIRE(ctx, n_iters >= 0, "gcs_sinkhorn_unbalanced_fixed_k: n_iters=%d", n_iters);
  const size_t kb = (((size_t)n_rows * n_cols * 8) + 255) & ~(size_t)255;
  int rc = gcs_ws_reserve(ctx, kb + (size_t)n_rows * 8);
  if (rc) return rc;
  double* Kmat = (double*)ctx->ws;
  double* u = (double*)((char*)ctx->ws + kb);
  sinkhorn_generic_kernel<<<1, kSkgThreads, 0, (cudaStream_t)stream>>>(cost, a, b, n_rows, n_cols, epsilon, tau_a, tau_b, n_iters, Kmat, u,
                                                                       out_pi);
  GCS_LAUNCH_CHECK(ctx);
  return GCS_OK;
}
