// gcs_export.cu -- the output side of the primitive map (export half): every valid
// primitive of the selected tiles -> world-frame moments -> newest-first order -> the /gc/map/points PointCloud2 payload.
//   extract_primitive_map_view / _extract_primitive_map_view_core   fl/backend/structures/primitive_map.py:474-576
//   renderable_batch_from_view                                       fl/backend/structures/primitive_map.py:580-616
//   PrimitiveMapPublisher.publish (concatenate, np.lexsort((id, -recency)))   fl/backend/map_publisher.py:131-258
//   _build_pointcloud2_from_view (x, y, z, intensity float32; 16-byte records)  fl/backend/map_publisher.py:44-90
// A 1 M-surfel map never leaves the device until its 16 MB wire payload is ready (the reference pulls every tile to
// the host and sorts there).
//
//   export_count_kernel    valid slots per tile (one CTA per tile)
//   export_compact_kernel  index-ordered compaction of the valid slots (tile order = publishing order), sort keys
//   cub::DeviceRadixSort   two stable passes, least significant key first: primitive id, then recency descending --
//                          np.lexsort((ids, -recency)).  CUB is library code used for this export step only.
//   export_gather_kernel   mu = solve(Lambda + eps I, theta), Sigma = inv(.), Lambda_world = inv(Sigma + eps I), eta,
//                          mass, colour, ids in the final order, and the packed cloud record
#include <cub/device/device_radix_sort.cuh>

#include "gcs_common.cuh"

namespace gcs {

namespace {

constexpr int kExpThreads = 1024;

struct ExpTiles { int index[256]; int n; };

__global__ void __launch_bounds__(kExpThreads) export_count_kernel(gcs_atlas A, ExpTiles T, int* __restrict__ counts) {
  __shared__ int sred[32];
  const int a = blockIdx.x, ti = T.index[a];
  int c = 0;
  if (ti >= 0)
    for (int s = threadIdx.x; s < A.m_tile; s += kExpThreads) c += A.valid[(int64_t)ti * A.m_tile + s] ? 1 : 0;
  for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
  if ((threadIdx.x & 31) == 0) sred[threadIdx.x >> 5] = c;
  __syncthreads();
  if (threadIdx.x == 0) {
    int t = 0;
    for (int w = 0; w < kExpThreads / 32; ++w) t += sred[w];
    counts[a] = t;
  }
}

// order-preserving int64 -> uint64 (ascending); ~ of it sorts descending
__device__ __forceinline__ unsigned long long i64_key(long long x) { return (unsigned long long)x ^ 0x8000000000000000ull; }

__global__ void __launch_bounds__(kExpThreads) export_compact_kernel(gcs_atlas A, ExpTiles T, const int* __restrict__ counts,
                                                                     long long* __restrict__ src /* tile row * M + slot */,
                                                                     unsigned long long* __restrict__ key_id,
                                                                     unsigned long long* __restrict__ key_rec,
                                                                     unsigned int* __restrict__ iota, int* __restrict__ total) {
  __shared__ int wsum[32];
  __shared__ int s_base;
  const int a = blockIdx.x, ti = T.index[a], tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) {
    int b = 0;
    for (int q = 0; q < a; ++q) b += counts[q];
    s_base = b;
    if (a == T.n - 1) *total = b + counts[a];
  }
  __syncthreads();
  if (ti < 0) return;
  int base = s_base;
  for (int s0 = 0; s0 < A.m_tile; s0 += kExpThreads) {
    const int s = s0 + tid;
    const int64_t o = (int64_t)ti * A.m_tile + s;
    const bool v = s < A.m_tile && A.valid[o];
    const unsigned m = __ballot_sync(0xffffffffu, v);
    if (lane == 0) wsum[warp] = __popc(m);
    __syncthreads();
    int before = 0, all = 0;
    for (int w = 0; w < kExpThreads / 32; ++w) { const int c = wsum[w]; if (w < warp) before += c; all += c; }
    if (v) {
      const int r = base + before + __popc(m & ((1u << lane) - 1u));
      src[r] = o;
      key_id[r] = i64_key(A.primitive_ids[o]);
      key_rec[r] = ~i64_key(A.last_supported_scan_seq[o]);   // -recency ascending
      iota[r] = (unsigned)r;
    }
    base += all;
    __syncthreads();
  }
}

__global__ void export_tail_kernel(unsigned long long* __restrict__ a, unsigned long long* __restrict__ b,
                                   unsigned int* __restrict__ idx, const int* __restrict__ total, int n) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k < n && k >= *total) { a[k] = ~0ull; b[k] = ~0ull; idx[k] = (unsigned)k; }
}

__global__ void export_permute_keys_kernel(const unsigned long long* __restrict__ key, const unsigned int* __restrict__ idx, int n,
                                           unsigned long long* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = key[idx[i]];
}

__global__ void export_gather_kernel(gcs_atlas A, const long long* __restrict__ src, const unsigned int* __restrict__ order,
                                     const int* __restrict__ total, double eps_lift, gcs_map_export E) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= *total) return;
  const long long o = src[order[i]];
  Mat3 L, Lr;
#pragma unroll
  for (int k = 0; k < 9; ++k) L.m[k] = A.Lambdas[o * 9 + k];
  Lr = L;
  Lr(0, 0) += eps_lift; Lr(1, 1) += eps_lift; Lr(2, 2) += eps_lift;
  const double th[3] = {A.thetas[o * 3], A.thetas[o * 3 + 1], A.thetas[o * 3 + 2]};
  double mu[3];
  mat3_solve(Lr, th, mu);
  const Mat3 Sig = mat3_inv(Lr);
  Mat3 Sr = Sig;
  Sr(0, 0) += eps_lift; Sr(1, 1) += eps_lift; Sr(2, 2) += eps_lift;
  const Mat3 Lw = mat3_inv(Sr);
  const double w = A.weights[o];
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    E.mu_world[(int64_t)i * 3 + k] = mu[k];
    E.color[(int64_t)i * 3 + k] = A.rgb[o * 3 + k];
  }
#pragma unroll
  for (int k = 0; k < 9; ++k) {
    E.Sigma_world[(int64_t)i * 9 + k] = Sig.m[k];
    E.Lambda_world[(int64_t)i * 9 + k] = Lw.m[k];
    E.eta[(int64_t)i * 9 + k] = A.etas[o * 9 + k];
  }
  E.mass[i] = w;
  E.primitive_ids[i] = A.primitive_ids[o];
  E.last_supported_scan_seq[i] = A.last_supported_scan_seq[o];
  // cloud record: float32 x, y, z, intensity = clip(float32(w), 0, 1e6)   (map_publisher.py:82-88)
  float4 rec;
  rec.x = (float)mu[0]; rec.y = (float)mu[1]; rec.z = (float)mu[2];
  rec.w = fminf(fmaxf((float)w, 0.0f), 1.0e6f);
  reinterpret_cast<float4*>(E.cloud)[i] = rec;
}

}  // namespace

}  // namespace gcs

using namespace gcs;

extern "C" int gcs_export_map_points(gcs_ctx* ctx, void* stream, const gcs_atlas* atlas, const int32_t* tile_index, int32_t n_tiles,
                                     double eps_lift, const gcs_map_export* out, int64_t capacity, int32_t* out_count) {
  if (!ctx) return GCS_EINVAL;
  GCS_CHECK_CUDA(ctx, cudaSetDevice(ctx->device));
  GCS_REQUIRE(ctx, atlas && tile_index && out && out_count, "gcs_export_map_points: NULL pointer");
  GCS_REQUIRE(ctx, n_tiles >= 1 && n_tiles <= 256, "gcs_export_map_points: n_tiles=%d outside [1, 256]", n_tiles);
  GCS_REQUIRE(ctx, capacity >= (int64_t)n_tiles * atlas->m_tile && capacity < (1ll << 31),
              "gcs_export_map_points: capacity %lld < n_tiles * m_tile = %lld (or >= 2^31)", (long long)capacity,
              (long long)n_tiles * atlas->m_tile);
  GCS_REQUIRE(ctx, out->mu_world && out->Sigma_world && out->Lambda_world && out->eta && out->mass && out->color &&
                       out->primitive_ids && out->last_supported_scan_seq && out->cloud,
              "gcs_export_map_points: an output array is NULL");
  GCS_REQUIRE(ctx, ((uintptr_t)out->cloud & 15) == 0, "gcs_export_map_points: cloud must be 16-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  ExpTiles T;
  T.n = n_tiles;
  for (int i = 0; i < 256; ++i) T.index[i] = i < n_tiles ? tile_index[i] : -1;
  const int n = (int)capacity;
  size_t cub_bytes = 0;
  GCS_CHECK_CUDA(ctx, cub::DeviceRadixSort::SortPairs(nullptr, cub_bytes, (const unsigned long long*)nullptr,
                                                      (unsigned long long*)nullptr, (const unsigned int*)nullptr,
                                                      (unsigned int*)nullptr, n, 0, 64, st));
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t o = off; off += (bytes + 255) & ~(size_t)255; return o; };
  const size_t o_cnt = take(sizeof(int) * 260), o_src = take((size_t)n * 8), o_k1 = take((size_t)n * 8), o_k2 = take((size_t)n * 8),
               o_k3 = take((size_t)n * 8), o_i0 = take((size_t)n * 4), o_i1 = take((size_t)n * 4), o_i2 = take((size_t)n * 4),
               o_cub = take(cub_bytes);
  int rc = gcs_ws_reserve(ctx, off);
  if (rc) return rc;
  char* ws = (char*)ctx->ws;
  int* counts = (int*)(ws + o_cnt);
  long long* src = (long long*)(ws + o_src);
  unsigned long long *k_id = (unsigned long long*)(ws + o_k1), *k_rec = (unsigned long long*)(ws + o_k2),
                     *k_tmp = (unsigned long long*)(ws + o_k3);
  unsigned int *i0 = (unsigned int*)(ws + o_i0), *i1 = (unsigned int*)(ws + o_i1), *i2 = (unsigned int*)(ws + o_i2);
  export_count_kernel<<<n_tiles, kExpThreads, 0, st>>>(*atlas, T, counts);
  GCS_LAUNCH_CHECK(ctx);
  export_compact_kernel<<<n_tiles, kExpThreads, 0, st>>>(*atlas, T, counts, src, k_id, k_rec, i0, out_count);
  GCS_LAUNCH_CHECK(ctx);
  // The number of valid primitives is only known on the device: all `capacity` slots are sorted, the unused tail
  // carries the largest key in both passes (it stays behind every real entry and is never gathered).
  const unsigned blocks = (unsigned)((n + 255) / 256);
  export_tail_kernel<<<blocks, 256, 0, st>>>(k_id, k_rec, i0, out_count, n);
  GCS_LAUNCH_CHECK(ctx);
  GCS_CHECK_CUDA(ctx, cub::DeviceRadixSort::SortPairs(ws + o_cub, cub_bytes, (const unsigned long long*)k_id, k_tmp,
                                                      (const unsigned int*)i0, i1, n, 0, 64, st));
  ctx->launches++;
  export_permute_keys_kernel<<<blocks, 256, 0, st>>>(k_rec, i1, n, k_id);
  GCS_LAUNCH_CHECK(ctx);
  GCS_CHECK_CUDA(ctx, cub::DeviceRadixSort::SortPairs(ws + o_cub, cub_bytes, (const unsigned long long*)k_id, k_tmp,
                                                      (const unsigned int*)i1, i2, n, 0, 64, st));
  ctx->launches++;
  export_gather_kernel<<<blocks, 256, 0, st>>>(*atlas, src, i2, out_count, eps_lift, *out);
  GCS_LAUNCH_CHECK(ctx);
  return GCS_OK;
}
