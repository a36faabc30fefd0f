// gcs_map_ops.cu -- the per-tile map operators of the reference, one C entry each ("Python
// signatures (must be kept)").  gcs_map_update (gcs_prims_map.cu) runs the same arithmetic for all active tiles of a scan
// in one pass; these entries serve callers that drive the operators one (block, tile) at a time as the reference does.
//   primitive_map_fuse           fl/backend/structures/primitive_map.py:992-1163
//   primitive_map_insert_masked  fl/backend/structures/primitive_map.py:807-981   (_select_lowest_mass_slots_fixed :326-353)
//   primitive_map_cull           fl/backend/structures/primitive_map.py:1175-1304
//   primitive_map_forget         fl/backend/structures/primitive_map.py:1314-1384
// Scatter-adds are sorted segmented sums (no float atomics): one warp per distinct slot, lanes stride the segment in
// proposal order, shuffle tree -- bit-identical run to run.  Selections follow lax.sort (stable, first operand is the key).
#include "gcs_select.cuh"

namespace gcs {

namespace {

constexpr int kOpsBig = 1024;
constexpr int kSweepMax = 64;          // blocks of the tile sweeps (256 threads each)
constexpr int kFuseMaxN = 16384;       // proposals per fuse call (in-CTA sort budget)

struct OpsSelectSmem {
  KeyIdx out[1024];
  int hist[256];
  int scan[2 * kOpsBig];
};
constexpr size_t kOpsSelectStatic = sizeof(OpsSelectSmem) + 4096 + 2048;   // + s_do + the statics of cta_select_k (30,752 B measured)

__device__ __forceinline__ void block_partials(const double* v, int n, double (*sred)[8], double* __restrict__ part) {
  for (int k = 0; k < n; ++k) {
    const double r = warp_sum(v[k]);
    if ((threadIdx.x & 31) == 0) sred[k][threadIdx.x >> 5] = r;
  }
  __syncthreads();
  if ((int)threadIdx.x < n) {
    double s = 0.0;
    for (int g = 0; g < 8; ++g) s += sred[threadIdx.x][g];
    part[(int64_t)blockIdx.x * n + threadIdx.x] = s;
  }
}

// ---------------------------------------------------------------------------------------------------- cull
// pass 1: number of valid slots and of valid slots below the weight threshold
__global__ void __launch_bounds__(256) cull_count_kernel(gcs_atlas A, int row, double thr, double* __restrict__ part) {
  __shared__ double sred[2][8];
  const int64_t base = (int64_t)row * A.m_tile;
  double v[2] = {0.0, 0.0};
  for (int s = blockIdx.x * 256 + threadIdx.x; s < A.m_tile; s += gridDim.x * 256) {
    if (A.valid[base + s]) { v[0] += 1.0; v[1] += (A.weights[base + s] < thr) ? 1.0 : 0.0; }
  }
  block_partials(v, 2, sred, part);
}

// pass 2 (one CTA): the effective threshold.  With max_primitives set and more than max_primitives survivors it is
// the weight of rank max_primitives in the descending order of weights * valid (:1226-1232) -- an 8-bit MSD radix
// rank over the tile, no sort.
__global__ void __launch_bounds__(kOpsBig) cull_threshold_kernel(gcs_atlas A, int row, double thr, int max_primitives,
                                                                 const double* __restrict__ part, int n_parts,
                                                                 double* __restrict__ thr_out) {
  __shared__ int h[256];
  __shared__ unsigned long long s_prefix;
  __shared__ int s_need;
  const int tid = threadIdx.x, M = A.m_tile;
  const int64_t base = (int64_t)row * M;
  double n_valid = 0.0, n_below = 0.0;
  for (int c = 0; c < n_parts; ++c) { n_valid += part[2 * c]; n_below += part[2 * c + 1]; }
  const int n_keep = (int)(n_valid - n_below);
  if (max_primitives < 0 || n_keep <= max_primitives || max_primitives >= M) {
    if (tid == 0) *thr_out = thr;
    return;
  }
  auto key = [&](int s) -> unsigned long long {
    const double w = A.weights[base + s] * (A.valid[base + s] ? 1.0 : 0.0);
    return f64_orderable(w + 0.0);   // -0.0 + 0.0 = +0.0: one key for both zeros, as the comparison sort sees them
  };
  unsigned long long prefix = 0ull, mask = 0ull;
  int need = M - max_primitives;   // 1-based ascending rank of the element at descending position max_primitives
  for (int shift = 56; shift >= 0; shift -= 8) {
    if (tid < 256) h[tid] = 0;
    __syncthreads();
    for (int s = tid; s < M; s += kOpsBig) {
      const unsigned long long k = key(s);
      if ((k & mask) == prefix) atomicAdd(&h[(int)((k >> shift) & 0xffull)], 1);
    }
    __syncthreads();
    if (tid == 0) {
      int acc = 0, b = 0;
      for (; b < 255; ++b) {
        if (acc + h[b] >= need) break;
        acc += h[b];
      }
      s_prefix = prefix | ((unsigned long long)b << shift);
      s_need = need - acc;
    }
    __syncthreads();
    prefix = s_prefix;
    need = s_need;
    mask |= 0xffull << shift;
    __syncthreads();
  }
  if (tid == 0) {
    // invert f64_orderable
    const unsigned long long b = (prefix & 0x8000000000000000ull) ? (prefix & 0x7fffffffffffffffull) : ~prefix;
    *thr_out = __longlong_as_double((long long)b);
  }
}

// pass 3: clear the slots below the effective threshold; partials n_culled, mass_dropped, sum of ALL weights (the
// certificate's denominator, :1293), valid slots left
__global__ void __launch_bounds__(256) cull_apply_kernel(gcs_atlas A, int row, const double* __restrict__ thr_eff,
                                                         double* __restrict__ part) {
  __shared__ double sred[4][8];
  const int64_t base = (int64_t)row * A.m_tile;
  const double thr = *thr_eff;
  double v[4] = {0.0, 0.0, 0.0, 0.0};
  for (int s = blockIdx.x * 256 + threadIdx.x; s < A.m_tile; s += gridDim.x * 256) {
    const int64_t o = base + s;
    const double w = A.weights[o];
    v[2] += w;
    if (A.valid[o]) {
      if (w < thr) { A.valid[o] = 0; v[0] += 1.0; v[1] += w; }
      else v[3] += 1.0;
    }
  }
  block_partials(v, 4, sred, part);
}

__global__ void ops_sum_parts_kernel(const double* __restrict__ part, int n_parts, int width, double* __restrict__ out) {
  const int k = threadIdx.x;
  if (k >= width) return;
  double a = 0.0;
  for (int c = 0; c < n_parts; ++c) a += part[(int64_t)c * width + k];
  out[k] = a;
}

// ---------------------------------------------------------------------------------------------------- forget
__global__ void __launch_bounds__(256) forget_kernel(gcs_atlas A, int row, double gamma) {
  const int64_t base = (int64_t)row * A.m_tile;
  for (int s = blockIdx.x * 256 + threadIdx.x; s < A.m_tile; s += gridDim.x * 256) A.weights[base + s] = gamma * A.weights[base + s];
}

// ---------------------------------------------------------------------------------------------------- fuse
struct FuseArgs {
  const int32_t* slots; const double* Lam; const double* th; const double* eta; const double* w; const double* resp;
  const uint8_t* valid; const double* colors; const int32_t* sources;
  int n;
};

// sort (slot, proposal) pairs; proposals whose slot is outside the tile sort last and are dropped, as a JAX scatter
// drops out-of-bounds updates.  stats[0] = number of distinct in-range slots = n_fused (:1155).
__global__ void __launch_bounds__(kOpsBig) fuse_sort_kernel(FuseArgs F, int m_tile, int n_pow2, unsigned long long* __restrict__ pairs,
                                                            double* __restrict__ stats) {
  extern __shared__ unsigned long long fs[];
  __shared__ int cnt;
  // n_fused of the reference = len(jnp.unique(target_slots)) over ALL entries as given (:1155): a first sort by the raw
  // value counts those
  for (int e = threadIdx.x; e < n_pow2; e += kOpsBig)
    fs[e] = e < F.n ? (((unsigned long long)((unsigned)F.slots[e] ^ 0x80000000u) << 32) | (unsigned)e) : ~0ull;
  if (threadIdx.x == 0) cnt = 0;
  cta_bitonic_sort_u64(fs, n_pow2);
  int local = 0;
  for (int e = threadIdx.x; e < F.n; e += kOpsBig)
    if (e == 0 || (unsigned)(fs[e - 1] >> 32) != (unsigned)(fs[e] >> 32)) ++local;
  if (local) atomicAdd(&cnt, local);
  __syncthreads();
  if (threadIdx.x == 0) stats[0] = (double)cnt;
  __syncthreads();
  // segments of the scatter-add: a negative index in [-m_tile, -1] wraps as .at[].add wraps it, anything else outside
  // [0, m_tile) is dropped as a JAX scatter drops it
  for (int e = threadIdx.x; e < n_pow2; e += kOpsBig) {
    unsigned long long x = ~0ull;
    if (e < F.n) {
      int s = F.slots[e];
      if (s < 0) s += m_tile;
      if (s >= 0 && s < m_tile) x = ((unsigned long long)(unsigned)s << 32) | (unsigned)e;
    }
    fs[e] = x;
  }
  cta_bitonic_sort_u64(fs, n_pow2);
  for (int e = threadIdx.x; e < n_pow2; e += kOpsBig) pairs[e] = fs[e];
}

constexpr int kFuseVals = 29;   // dLambda 9, deta 9, dtheta 3, dw, dr, dcam, dlid, dacc 3, dden
__global__ void __launch_bounds__(256) fuse_apply_kernel(gcs_atlas A, int row, FuseArgs F, const unsigned long long* __restrict__ pairs,
                                                         int n_pow2, double timestamp, long long scan_seq) {
  const int lane = threadIdx.x & 31;
  const int q = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (q >= n_pow2) return;
  const unsigned long long e = pairs[q];
  const unsigned key = (unsigned)(e >> 32);
  const bool head = e != ~0ull && (q == 0 || (unsigned)(pairs[q - 1] >> 32) != key);   // warp-uniform
  if (!head) return;
  double v[kFuseVals];
#pragma unroll
  for (int k = 0; k < kFuseVals; ++k) v[k] = 0.0;
  for (int base = q;; base += 32) {
    const int j = base + lane;
    bool in_seg = j < n_pow2;
    unsigned long long ej = 0;
    if (in_seg) { ej = pairs[j]; in_seg = ej != ~0ull && (unsigned)(ej >> 32) == key; }
    if (in_seg) {
      const int i = (int)(unsigned)(ej & 0xffffffffull);
      const double r = F.resp[i] * ((F.valid && !F.valid[i]) ? 0.0 : 1.0);
      const double wm = F.w[i];
#pragma unroll
      for (int k = 0; k < 9; ++k) { v[k] += r * F.Lam[9 * i + k]; v[9 + k] += r * F.eta[9 * i + k]; }
#pragma unroll
      for (int k = 0; k < 3; ++k) v[18 + k] += r * F.th[3 * i + k];
      v[21] += r * wm; v[22] += r;
      if (F.sources) {
        const double wc = r * wm * (F.sources[i] == 0 ? 1.0 : 0.0), wl = r * wm * (F.sources[i] == 1 ? 1.0 : 0.0);
        v[23] += wc; v[24] += wl;
        if (F.colors) {
          v[28] += wc;
#pragma unroll
          for (int k = 0; k < 3; ++k) v[25 + k] += fmin(fmax(F.colors[3 * i + k], 0.0), 1.0) * wc;
        }
      }
    }
    if (!__all_sync(0xffffffffu, in_seg)) break;
  }
#pragma unroll
  for (int k = 0; k < kFuseVals; ++k) v[k] = warp_sum(v[k]);
  if (lane == 0) {
    const int64_t o = (int64_t)row * A.m_tile + (int)key;
    for (int k = 0; k < 9; ++k) { A.Lambdas[9 * o + k] += v[k]; A.etas[9 * o + k] += v[9 + k]; }
    for (int k = 0; k < 3; ++k) { A.thetas[3 * o + k] += v[18 + k]; A.rgb_cam_accum[3 * o + k] += v[25 + k]; }
    A.weights[o] += v[21]; A.cam_mass[o] += v[23]; A.lidar_mass[o] += v[24]; A.rgb_cam_denom[o] += v[28];
    if (v[22] > 0.0) { A.last_supported_scan_seq[o] = scan_seq; A.last_update_scan_seq[o] = scan_seq; }
    A.timestamps[o] = timestamp;   // every slot named in target_slots, masked or not (:1112)
  }
}

// rgb / colors of every slot of the tile from the camera accumulators (:1097-1104)
__global__ void __launch_bounds__(256) fuse_rgb_sweep_kernel(gcs_atlas A, int row, double eps_mass) {
  const int64_t base = (int64_t)row * A.m_tile;
  for (int s = blockIdx.x * 256 + threadIdx.x; s < A.m_tile; s += gridDim.x * 256) {
    const int64_t o = base + s;
    const bool has = A.cam_mass[o] > 0.0;
    const double den = fmax(A.rgb_cam_denom[o], eps_mass);
    for (int k = 0; k < 3; ++k) {
      const double est = fmin(fmax(A.rgb_cam_accum[3 * o + k] / den, 0.0), 1.0);
      const double c = has ? est : 0.5;
      A.rgb[3 * o + k] = c;
      A.colors[3 * o + k] = c;
    }
  }
}

// ---------------------------------------------------------------------------------------------------- insert_masked
struct InsertArgs {
  const double* Lam; const double* th; const double* eta; const double* w; const uint8_t* valid_new;
  const double* colors; const int32_t* sources;
  int k;
  double timestamp, lambda;
  long long scan_seq, next_global_id;
};

__global__ void __launch_bounds__(kOpsBig) insert_masked_kernel(gcs_atlas A, int row, InsertArgs I, long long* __restrict__ out_ids,
                                                                int* __restrict__ out_slots, double* __restrict__ stats,
                                                                int use_cache) {
  __shared__ OpsSelectSmem sm;
  __shared__ int s_do[kOpsBig];
  extern __shared__ uint32_t key_cache[];
  const int tid = threadIdx.x, k = I.k;
  const int64_t base = (int64_t)row * A.m_tile;
  // eviction targets: k lowest retention, empty slots first, ties by slot index
  auto key = [&](int s) -> unsigned long long {
    const int64_t o = base + s;
    const uint8_t v = A.valid[o];                         // three independent loads: the sweeps are latency-bound
    const long long lss = A.last_supported_scan_seq[o];
    const double w = A.weights[o];
    double keyv = -INFINITY;
    if (v) {
      long long dt = I.scan_seq - lss;
      if (dt < 0) dt = 0;
      keyv = w * exp(-I.lambda * (double)dt);
    }
    return f64_orderable(keyv);
  };
  auto empty = [&](int s) -> bool { return A.valid[base + s] == 0; };
  if (!cta_select_min_sentinel(A.m_tile, k, empty, f64_orderable(-INFINITY), sm.out, sm.scan))
    cta_select_k(A.m_tile, k, key, sm.out, sm.hist, sm.scan, use_cache ? key_cache : nullptr);
  s_do[tid] = (tid < k && I.valid_new[tid]) ? 1 : 0;
  __syncthreads();
  // inclusive scan of the proposal mask (Hillis-Steele over 1024 entries, double-buffered through sm.scan)
  int* buf = sm.scan;
  buf[tid] = s_do[tid];
  __syncthreads();
  int cur = 0;
  for (int off = 1; off < kOpsBig; off <<= 1, cur ^= 1) {
    int x = buf[cur * kOpsBig + tid];
    if (tid >= off) x += buf[cur * kOpsBig + tid - off];
    buf[(cur ^ 1) * kOpsBig + tid] = x;
    __syncthreads();
  }
  const int prefix = buf[cur * kOpsBig + tid];
  const int n_ins = buf[cur * kOpsBig + kOpsBig - 1];
  if (tid < k) {
    const bool doit = s_do[tid] != 0;
    const int slot = sm.out[tid].idx;
    const long long nid = doit ? I.next_global_id + (prefix - 1) : -1;
    out_ids[tid] = nid;
    out_slots[tid] = slot;
    if (doit) {
      const int64_t o = base + slot;
      const double w = I.w[tid];
      const bool is_cam = I.sources ? I.sources[tid] == 0 : false, is_lid = I.sources ? I.sources[tid] == 1 : true;
      const double cam = is_cam ? w : 0.0;
      for (int c = 0; c < 9; ++c) { A.Lambdas[9 * o + c] = I.Lam[9 * tid + c]; A.etas[9 * o + c] = I.eta[9 * tid + c]; }
      for (int c = 0; c < 3; ++c) {
        A.thetas[3 * o + c] = I.th[3 * tid + c];
        const double col = I.colors ? I.colors[3 * tid + c] : 0.0;
        const double rgbn = (cam > 0.0) ? fmin(fmax(col, 0.0), 1.0) : 0.5;
        A.colors[3 * o + c] = rgbn; A.rgb[3 * o + c] = rgbn;
        A.rgb_cam_accum[3 * o + c] = col * cam;
      }
      A.weights[o] = w; A.timestamps[o] = I.timestamp; A.created_timestamps[o] = I.timestamp;
      A.last_supported_scan_seq[o] = I.scan_seq; A.last_update_scan_seq[o] = I.scan_seq;
      A.primitive_ids[o] = nid; A.valid[o] = 1;
      A.cam_mass[o] = cam; A.lidar_mass[o] = is_lid ? w : 0.0; A.rgb_cam_denom[o] = cam;
    }
  }
  if (tid == 0) { stats[0] = (double)n_ins; stats[1] = (double)(k - n_ins); }
}

__global__ void __launch_bounds__(256) count_valid_kernel(gcs_atlas A, int row, double* __restrict__ part) {
  __shared__ double sred[1][8];
  const int64_t base = (int64_t)row * A.m_tile;
  double v[1] = {0.0};
  for (int s = blockIdx.x * 256 + threadIdx.x; s < A.m_tile; s += gridDim.x * 256) v[0] += A.valid[base + s] ? 1.0 : 0.0;
  block_partials(v, 1, sred, part);
}

int sweep_blocks_of(int m_tile) {
  const int b = (m_tile + 255) / 256;
  return b < kSweepMax ? b : kSweepMax;
}

int check_tile(gcs_ctx* ctx, const gcs_atlas* a, int32_t row, const char* who) {
  GCS_REQUIRE(ctx, a && a->Lambdas && a->thetas && a->etas && a->weights && a->timestamps && a->created_timestamps &&
                       a->last_supported_scan_seq && a->last_update_scan_seq && a->primitive_ids && a->valid && a->colors &&
                       a->cam_mass && a->lidar_mass && a->rgb_cam_accum && a->rgb_cam_denom && a->rgb,
              "%s: atlas pointer is NULL", who);
  GCS_REQUIRE(ctx, a->m_tile >= 1 && a->n_tiles_cap >= 1, "%s: bad atlas shape", who);
  GCS_REQUIRE(ctx, row >= 0 && row < a->n_tiles_cap, "%s: tile_index %d outside the pool of %d tiles", who, row, a->n_tiles_cap);
  return GCS_OK;
}

}  // namespace

}  // namespace gcs

using namespace gcs;

extern "C" int gcs_map_cull(gcs_ctx* ctx, void* stream, const gcs_atlas* atlas, int32_t tile_index, double weight_threshold,
                            int32_t max_primitives, double* stats) {
  if (!ctx) return GCS_EINVAL;
  GCS_CHECK_CUDA(ctx, cudaSetDevice(ctx->device));
  int rc = check_tile(ctx, atlas, tile_index, "gcs_map_cull");
  if (rc) return rc;
  GCS_REQUIRE(ctx, stats, "gcs_map_cull: stats is NULL");
  cudaStream_t st = (cudaStream_t)stream;
  const int nb = sweep_blocks_of(atlas->m_tile);
  rc = gcs_ws_reserve(ctx, (size_t)nb * 4 * 8 + 256);
  if (rc) return rc;
  double* part = (double*)ctx->ws;
  double* thr_eff = part + (size_t)nb * 4;
  cull_count_kernel<<<nb, 256, 0, st>>>(*atlas, tile_index, weight_threshold, part);
  GCS_LAUNCH_CHECK(ctx);
  cull_threshold_kernel<<<1, kOpsBig, 0, st>>>(*atlas, tile_index, weight_threshold, max_primitives, part, nb, thr_eff);
  GCS_LAUNCH_CHECK(ctx);
  cull_apply_kernel<<<nb, 256, 0, st>>>(*atlas, tile_index, thr_eff, part);
  GCS_LAUNCH_CHECK(ctx);
  ops_sum_parts_kernel<<<1, 32, 0, st>>>(part, nb, 4, stats);
  GCS_LAUNCH_CHECK(ctx);
  return GCS_OK;
}

extern "C" int gcs_map_forget(gcs_ctx* ctx, void* stream, const gcs_atlas* atlas, int32_t tile_index, double forgetting_factor) {
  if (!ctx) return GCS_EINVAL;
  GCS_CHECK_CUDA(ctx, cudaSetDevice(ctx->device));
  int rc = check_tile(ctx, atlas, tile_index, "gcs_map_forget");
  if (rc) return rc;
  forget_kernel<<<sweep_blocks_of(atlas->m_tile), 256, 0, (cudaStream_t)stream>>>(*atlas, tile_index, forgetting_factor);
  GCS_LAUNCH_CHECK(ctx);
  return GCS_OK;
}

extern "C" int gcs_map_fuse(gcs_ctx* ctx, void* stream, const gcs_atlas* atlas, int32_t tile_index, const int32_t* target_slots,
                            const double* Lambdas_meas, const double* thetas_meas, const double* etas_meas,
                            const double* weights_meas, const double* responsibilities, const uint8_t* valid_mask,
                            const double* colors_meas, const int32_t* sources_meas, int32_t n, double timestamp,
                            int64_t scan_seq, double eps_mass, double* stats) {
  if (!ctx) return GCS_EINVAL;
  GCS_CHECK_CUDA(ctx, cudaSetDevice(ctx->device));
  int rc = check_tile(ctx, atlas, tile_index, "gcs_map_fuse");
  if (rc) return rc;
  GCS_REQUIRE(ctx, target_slots && Lambdas_meas && thetas_meas && etas_meas && weights_meas && responsibilities && stats,
              "gcs_map_fuse: NULL pointer");
  GCS_REQUIRE(ctx, n >= 1 && n <= kFuseMaxN, "gcs_map_fuse: n=%d outside [1, %d] proposals per call", n, kFuseMaxN);
  cudaStream_t st = (cudaStream_t)stream;
  int n_pow2 = 1;
  while (n_pow2 < n) n_pow2 <<= 1;
  rc = gcs_ws_reserve(ctx, (size_t)n_pow2 * 8);
  if (rc) return rc;
  unsigned long long* pairs = (unsigned long long*)ctx->ws;
  FuseArgs F;
  F.slots = target_slots; F.Lam = Lambdas_meas; F.th = thetas_meas; F.eta = etas_meas; F.w = weights_meas; F.resp = responsibilities;
  F.valid = valid_mask; F.colors = colors_meas; F.sources = sources_meas; F.n = n;
  GCS_CHECK_CUDA(ctx, gcs_smem_attr_once((const void*)fuse_sort_kernel, kFuseMaxN * 8));
  fuse_sort_kernel<<<1, kOpsBig, (size_t)n_pow2 * 8, st>>>(F, atlas->m_tile, n_pow2, pairs, stats);
  GCS_LAUNCH_CHECK(ctx);
  fuse_apply_kernel<<<(n_pow2 + 7) / 8, 256, 0, st>>>(*atlas, tile_index, F, pairs, n_pow2, timestamp, (long long)scan_seq);
  GCS_LAUNCH_CHECK(ctx);
  fuse_rgb_sweep_kernel<<<sweep_blocks_of(atlas->m_tile), 256, 0, st>>>(*atlas, tile_index, eps_mass);
  GCS_LAUNCH_CHECK(ctx);
  return GCS_OK;
}

extern "C" int gcs_map_insert_masked(gcs_ctx* ctx, void* stream, const gcs_atlas* atlas, int32_t tile_index,
                                     const double* Lambdas_new, const double* thetas_new, const double* etas_new,
                                     const double* weights_new, const uint8_t* valid_new_mask, const double* colors_new,
                                     const int32_t* sources_new, int32_t k, double timestamp, int64_t scan_seq,
                                     double recency_decay_lambda, int64_t next_global_id, int64_t* out_new_ids,
                                     int32_t* out_target_slots, double* stats) {
  if (!ctx) return GCS_EINVAL;
  GCS_CHECK_CUDA(ctx, cudaSetDevice(ctx->device));
  int rc = check_tile(ctx, atlas, tile_index, "gcs_map_insert_masked");
  if (rc) return rc;
  GCS_REQUIRE(ctx, Lambdas_new && thetas_new && etas_new && weights_new && valid_new_mask && out_new_ids && out_target_slots && stats,
              "gcs_map_insert_masked: NULL pointer");
  GCS_REQUIRE(ctx, k >= 1 && k <= 1024 && k <= atlas->m_tile, "gcs_map_insert_masked: k=%d outside [1, min(1024, m_tile=%d)]", k,
              atlas->m_tile);
  cudaStream_t st = (cudaStream_t)stream;
  InsertArgs I;
  I.Lam = Lambdas_new; I.th = thetas_new; I.eta = etas_new; I.w = weights_new; I.valid_new = valid_new_mask; I.colors = colors_new;
  I.sources = sources_new; I.k = k; I.timestamp = timestamp; I.lambda = recency_decay_lambda; I.scan_seq = scan_seq;
  I.next_global_id = next_global_id;
  const size_t want = (size_t)atlas->m_tile * sizeof(uint32_t);
  const size_t kc = want + kOpsSelectStatic <= 232448 ? want : 0;
  if (kc + kOpsSelectStatic > 48 * 1024) GCS_CHECK_CUDA(ctx, gcs_smem_attr_once((const void*)insert_masked_kernel, (int)kc));
  insert_masked_kernel<<<1, kOpsBig, kc, st>>>(*atlas, tile_index, I, (long long*)out_new_ids, out_target_slots, stats, kc ? 1 : 0);
  GCS_LAUNCH_CHECK(ctx);
  const int nb = sweep_blocks_of(atlas->m_tile);
  rc = gcs_ws_reserve(ctx, (size_t)nb * 8);
  if (rc) return rc;
  count_valid_kernel<<<nb, 256, 0, st>>>(*atlas, tile_index, (double*)ctx->ws);
  GCS_LAUNCH_CHECK(ctx);
  ops_sum_parts_kernel<<<1, 32, 0, st>>>((double*)ctx->ws, nb, 1, stats + 2);
  GCS_LAUNCH_CHECK(ctx);
  return GCS_OK;
}
