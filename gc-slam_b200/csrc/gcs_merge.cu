// gcs_merge.cu -- primitive_map_merge_reduce (merge half):
// fl/backend/structures/primitive_map.py:1501-2031.  All-pairs Bhattacharyya distance of a tile's Gaussians, greedy
// disjoint selection of at most max_pairs pairs in stable ascending order of distance (below merge_threshold), moment-
// matched merge of each pair into its first slot.  The reference caps the operator at max_tile_size = 2048 slots (it is
// O(M^2)); the same cap applies here, so the (M choose 2) <= 2.1 M distances live in the workspace.
//
// "Greedy in sorted order, skipping pairs that touch a used primitive" picks, at every step, the smallest (distance,
// pair index) among the pairs whose two ends are still free -- so no sort is needed: max_pairs (reference: 4) masked
// arg-min sweeps over the stored distances by one CTA, ties to the smaller pair index (the stable order of argsort
// over the triu enumeration).  The selected pairs are disjoint, hence merged in parallel.
#include "gcs_common.cuh"

namespace gcs {

namespace {

constexpr int kMrMaxM = 2048;
constexpr int kMrMaxPairs = 64;
constexpr int kMrThreads = 1024;

struct MergeWs { double* mu; double* Sig; double* det; double* dist; };

__device__ __forceinline__ long long triu_index(int i, int j, int M) {   // position of (i, j), i < j, in np.triu_indices(M, 1)
  return (long long)i * (2 * M - i - 1) / 2 + (j - i - 1);
}

__global__ void merge_moments_kernel(gcs_atlas A, int ti, double eps_lift, MergeWs W) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= A.m_tile) return;
  const int64_t o = (int64_t)ti * A.m_tile + i;
  Mat3 L;
#pragma unroll
  for (int k = 0; k < 9; ++k) L.m[k] = A.Lambdas[o * 9 + k];
  L(0, 0) += eps_lift; L(1, 1) += eps_lift; L(2, 2) += eps_lift;
  const double th[3] = {A.thetas[o * 3], A.thetas[o * 3 + 1], A.thetas[o * 3 + 2]};
  double mu[3];
  mat3_solve(L, th, mu);
  const Mat3 S = mat3_inv(L);
#pragma unroll
  for (int k = 0; k < 3; ++k) W.mu[i * 3 + k] = mu[k];
#pragma unroll
  for (int k = 0; k < 9; ++k) W.Sig[i * 9 + k] = S.m[k];
  W.det[i] = mat3_det(S);
}

// Bhattacharyya distance of pair (i, j) (:1925-1936); invalid pairs -> +inf
__global__ void merge_dist_kernel(gcs_atlas A, int ti, double eps_lift, MergeWs W) {
  const int M = A.m_tile, i = blockIdx.y, j = i + 1 + blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= M) return;
  const int64_t base = (int64_t)ti * M;
  double d = INFINITY;
  if (A.valid[base + i] && A.valid[base + j]) {
    Mat3 S, Sr;
#pragma unroll
    for (int k = 0; k < 9; ++k) S.m[k] = 0.5 * (W.Sig[i * 9 + k] + W.Sig[j * 9 + k]);
    Sr = S;
    Sr(0, 0) += eps_lift; Sr(1, 1) += eps_lift; Sr(2, 2) += eps_lift;
    const Mat3 Si = mat3_inv(Sr);
    const double dm[3] = {W.mu[i * 3] - W.mu[j * 3], W.mu[i * 3 + 1] - W.mu[j * 3 + 1], W.mu[i * 3 + 2] - W.mu[j * 3 + 2]};
    double v[3];
    mat3_vec(Si, dm, v);
    const double quad = 0.125 * (dm[0] * v[0] + dm[1] * v[1] + dm[2] * v[2]);
    const double lt = 0.5 * log(mat3_det(S) / sqrt(W.det[i] * W.det[j] + 1e-24));
    d = quad + lt;
  }
  W.dist[triu_index(i, j, M)] = d;
}

__global__ void __launch_bounds__(kMrThreads) merge_select_apply_kernel(gcs_atlas A, int ti, double thr, int max_pairs,
                                                                        double eps_psd, MergeWs W, double* __restrict__ stats) {
  __shared__ uint8_t used[kMrMaxM];
  __shared__ int sel_i[kMrMaxPairs], sel_j[kMrMaxPairs];
  __shared__ double wd[32];
  __shared__ long long wk[32];
  __shared__ int s_n, s_stop;
  const int M = A.m_tile, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int i = tid; i < M; i += kMrThreads) used[i] = 0;
  if (tid == 0) { s_n = 0; s_stop = 0; }
  __syncthreads();
  for (int r = 0; r < max_pairs; ++r) {
    double bd = INFINITY;
    long long bk = 0x7fffffffffffffffll;
    for (int i = 0; i < M - 1; ++i) {
      if (used[i]) continue;
      const long long row = triu_index(i, i + 1, M);
      for (int j = i + 1 + tid; j < M; j += kMrThreads) {
        if (used[j]) continue;
        const long long k = row + (j - i - 1);
        const double d = W.dist[k];
        if (d < thr && d > -INFINITY && (d < bd || (d == bd && k < bk))) { bd = d; bk = k; }   // finite and below the threshold
      }
    }
    // block arg-min by (distance, pair index)
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const double od = __shfl_xor_sync(0xffffffffu, bd, o);
      const long long ok = __shfl_xor_sync(0xffffffffu, bk, o);
      if (od < bd || (od == bd && ok < bk)) { bd = od; bk = ok; }
    }
    if (lane == 0) { wd[warp] = bd; wk[warp] = bk; }
    __syncthreads();
    if (tid == 0) {
      double d = INFINITY;
      long long k = 0x7fffffffffffffffll;
      for (int w = 0; w < kMrThreads / 32; ++w)
        if (wd[w] < d || (wd[w] == d && wk[w] < k)) { d = wd[w]; k = wk[w]; }
      if (d < INFINITY) {
        // decode k -> (i, j): the row whose first pair index is the largest one <= k
        int i = 0;
        while (i + 1 < M - 1 && triu_index(i + 1, i + 2, M) <= k) ++i;
        const int j = i + 1 + (int)(k - triu_index(i, i + 1, M));
        used[i] = 1; used[j] = 1;
        sel_i[s_n] = i; sel_j[s_n] = j;
        ++s_n;
      } else {
        s_stop = 1;
      }
    }
    __syncthreads();
    if (s_stop) break;
  }
  const int n_sel = s_n;
  // moment-matched merges (:1640-1718), one thread per (disjoint) pair
  if (tid < n_sel) {
    const int i = sel_i[tid], j = sel_j[tid];
    const int64_t oi = (int64_t)ti * M + i, oj = (int64_t)ti * M + j;
    const double w1 = A.weights[oi], w2 = A.weights[oj], wsum = w1 + w2;
    if (wsum > 0.0) {
      double mum[3], d1[3], d2[3];
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        mum[c] = (w1 * W.mu[i * 3 + c] + w2 * W.mu[j * 3 + c]) / wsum;
        d1[c] = W.mu[i * 3 + c] - mum[c];
        d2[c] = W.mu[j * 3 + c] - mum[c];
      }
      Mat3 Sm;
#pragma unroll
      for (int a = 0; a < 3; ++a)
#pragma unroll
        for (int b = 0; b < 3; ++b)
          Sm(a, b) = (w1 * (W.Sig[i * 9 + 3 * a + b] + d1[a] * d1[b]) + w2 * (W.Sig[j * 9 + 3 * a + b] + d2[a] * d2[b])) / wsum +
                     (a == b ? eps_psd : 0.0);
      const Mat3 Lm = mat3_inv(Sm);
      double thm[3];
      mat3_vec(Lm, mum, thm);
#pragma unroll
      for (int k = 0; k < 9; ++k) {
        A.Lambdas[oi * 9 + k] = Lm.m[k];
        A.etas[oi * 9 + k] = (w1 * A.etas[oi * 9 + k] + w2 * A.etas[oj * 9 + k]) / wsum;
      }
      const double cam = A.cam_mass[oi] + A.cam_mass[oj];
      const double den = A.rgb_cam_denom[oi] + A.rgb_cam_denom[oj];
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        A.thetas[oi * 3 + c] = thm[c];
        const double acc = A.rgb_cam_accum[oi * 3 + c] + A.rgb_cam_accum[oj * 3 + c];
        const double est = fmin(fmax(acc / fmax(den, eps_psd), 0.0), 1.0);
        const double rgb = cam > 0.0 ? est : 0.5;
        A.rgb_cam_accum[oi * 3 + c] = acc;
        A.colors[oi * 3 + c] = rgb;
        A.rgb[oi * 3 + c] = rgb;
      }
      A.weights[oi] = wsum;
      A.cam_mass[oi] = cam;
      A.lidar_mass[oi] = A.lidar_mass[oi] + A.lidar_mass[oj];
      A.rgb_cam_denom[oi] = den;
      A.timestamps[oi] = fmax(A.timestamps[oi], A.timestamps[oj]);
      A.created_timestamps[oi] = fmin(A.created_timestamps[oi], A.created_timestamps[oj]);
      const long long ls = A.last_supported_scan_seq[oj], lu = A.last_update_scan_seq[oj];
      if (ls > A.last_supported_scan_seq[oi]) A.last_supported_scan_seq[oi] = ls;
      if (lu > A.last_update_scan_seq[oi]) A.last_update_scan_seq[oi] = lu;
      A.weights[oj] = 0.0;
      A.valid[oj] = 0;
    }
  }
  if (tid == 0) {
    stats[GCS_MR_N_MERGED] = (double)n_sel;
    stats[GCS_MR_STATUS] = n_sel > 0 ? 1.0 : 0.0;
    for (int k = GCS_MR_STATUS + 1; k < GCS_MR_NSTATS; ++k) stats[k] = 0.0;
  }
}

}  // namespace

}  // namespace gcs

using namespace gcs;

extern "C" int gcs_map_merge_reduce(gcs_ctx* ctx, void* stream, const gcs_atlas* atlas, int32_t tile_index,
                                    double merge_threshold, int32_t max_pairs, double eps_psd, double eps_lift, double* stats) {
  if (!ctx) return GCS_EINVAL;
  GCS_CHECK_CUDA(ctx, cudaSetDevice(ctx->device));
  GCS_REQUIRE(ctx, atlas && stats, "gcs_map_merge_reduce: NULL pointer");
  const int M = atlas->m_tile;
  GCS_REQUIRE(ctx, tile_index >= 0 && tile_index < atlas->n_tiles_cap, "gcs_map_merge_reduce: tile_index %d outside the pool", tile_index);
  GCS_REQUIRE(ctx, M >= 2 && M <= kMrMaxM,
              "gcs_map_merge_reduce: m_tile=%d outside [2, %d] (the reference's budget cap GC_PRIMITIVE_MERGE_MAX_TILE_SIZE)", M, kMrMaxM);
  GCS_REQUIRE(ctx, max_pairs >= 1 && max_pairs <= kMrMaxPairs, "gcs_map_merge_reduce: max_pairs=%d outside [1, %d]", max_pairs, kMrMaxPairs);
  cudaStream_t st = (cudaStream_t)stream;
  const size_t n_pairs = (size_t)M * (M - 1) / 2;
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t o = off; off += (bytes + 255) & ~(size_t)255; return o; };
  const size_t o_mu = take((size_t)M * 3 * 8), o_sig = take((size_t)M * 9 * 8), o_det = take((size_t)M * 8), o_d = take(n_pairs * 8);
  int rc = gcs_ws_reserve(ctx, off);
  if (rc) return rc;
  char* ws = (char*)ctx->ws;
  MergeWs W;
  W.mu = (double*)(ws + o_mu); W.Sig = (double*)(ws + o_sig); W.det = (double*)(ws + o_det); W.dist = (double*)(ws + o_d);
  merge_moments_kernel<<<(M + 127) / 128, 128, 0, st>>>(*atlas, tile_index, eps_lift, W);
  GCS_LAUNCH_CHECK(ctx);
  merge_dist_kernel<<<dim3((unsigned)((M + 255) / 256), (unsigned)(M - 1)), 256, 0, st>>>(*atlas, tile_index, eps_lift, W);
  GCS_LAUNCH_CHECK(ctx);
  merge_select_apply_kernel<<<1, kMrThreads, 0, st>>>(*atlas, tile_index, merge_threshold, max_pairs, eps_psd, W, stats);
  GCS_LAUNCH_CHECK(ctx);
  return GCS_OK;
}
