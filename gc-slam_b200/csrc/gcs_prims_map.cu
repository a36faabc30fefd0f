// gcs_prims_map.cu -- primitive-map side of the LiDAR evidence path:
//   a11  recency inflation, per-tile top-k view extraction
//   a12  OT association (pool cost -> stable top-K -> unbalanced Sinkhorn, 50 fixed iterations)
//   a13  pose evidence from soft correspondences (WLS translation + scatter-SVD rotation)
//   a14  map update: PoE fuse (own stable radix sort by target + segmented sums, no float atomics), insert/evict, cull, forget
// All selections reproduce jnp.argsort / lax.sort semantics (stable, first operand is the only key) through
// cta_select_k (gcs_select.cuh).  All floating reductions are fixed-order.
#include <cooperative_groups.h>

#include "gcs_assoc.cuh"
#include "gcs_select.cuh"
#include "gcs_tma.cuh"

namespace gcs {

constexpr int kBig = 1024;  // CTA size of the single-CTA kernels

__device__ __forceinline__ double block_sum_1024(double v, double* sred) {
  v = warp_sum(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) sred[threadIdx.x >> 5] = v;
  __syncthreads();
  double s = 0.0;
  for (int i = 0; i < (int)(blockDim.x >> 5); ++i) s += sred[i];
  __syncthreads();
  return s;  // valid in every thread
}

__device__ __forceinline__ void load_mat3(const double* p, Mat3& M) {
#pragma unroll
  for (int k = 0; k < 9; ++k) M.m[k] = p[k];
}

// mean position / direction / kappa of one measurement row (measurement_batch.py:390-411)
__device__ inline void meas_row_moments(const gcs_meas_batch& B, int i, double eps_lift, double eps_mass, double* mu,
                                        double* dir, double* kappa) {
  Mat3 L;
  load_mat3(B.Lambdas + 9 * i, L);
  L(0, 0) += eps_lift; L(1, 1) += eps_lift; L(2, 2) += eps_lift;
  const double th[3] = {B.thetas[3 * i], B.thetas[3 * i + 1], B.thetas[3 * i + 2]};
  mat3_solve(L, th, mu);
  double es[3];
  for (int k = 0; k < 3; ++k) es[k] = B.etas[9 * i + k] + B.etas[9 * i + 3 + k] + B.etas[9 * i + 6 + k];
  const double n = sqrt(es[0] * es[0] + es[1] * es[1] + es[2] * es[2]);
  *kappa = n;
  for (int k = 0; k < 3; ++k) dir[k] = es[k] / (n + eps_mass);
}

__device__ __forceinline__ long long pack_tile_id(long long c1, long long c2, long long cz) {
  const long long m = (1ll << 21) - 1, b = 1ll << 20;  // fl/common/tiling.py:74-100
  return (((c1 + b) & m) << 42) | (((c2 + b) & m) << 21) | ((cz + b) & m);
}
__device__ __forceinline__ void tile_cell(const double* p, double h, long long* c) {
  const double s2 = p[0] * 0.5 + p[1] * (sqrt(3.0) * 0.5);
  c[0] = (long long)floor(p[0] / h); c[1] = (long long)floor(s2 / h); c[2] = (long long)floor(p[2] / h);
}

// =================================================================================================
// a11: recency inflation
// =================================================================================================
struct TileList {
  int32_t index[16];
  int64_t id[16];
  int n;
};

// apply == 0: statistics only, the map is left untouched (the read-only hypotheses of a batch see the inflated values
// through map_view_kernel's functional inflation)
__global__ void __launch_bounds__(256) recency_inflate_kernel(gcs_atlas A, TileList T, long long scan_seq, double lam,
                                                              double min_scale, double* __restrict__ part, int apply) {
  __shared__ double sred[3][8];
  const int a = blockIdx.y;
  const int ti = T.index[a];
  double s_down = 0.0, s_tr = 0.0, s_n = 0.0;
  if (ti >= 0) {
    const int64_t base = (int64_t)ti * A.m_tile;
    for (int s = blockIdx.x * 256 + threadIdx.x; s < A.m_tile; s += gridDim.x * 256) {
      const int64_t o = base + s;
      const bool v = A.valid[o] != 0;
      long long dt = scan_seq - A.last_supported_scan_seq[o];
      if (dt < 0) dt = 0;
      double decay = exp(-lam * (double)dt);
      decay = fmin(fmax(decay, min_scale), 1.0);
      if (!v) decay = 1.0;
      if (v) {
        if (apply) {
          for (int k = 0; k < 9; ++k) A.Lambdas[9 * o + k] *= decay;
          for (int k = 0; k < 3; ++k) A.thetas[3 * o + k] *= decay;
        }
        s_down += 1.0 - decay; s_tr += 1.0 / decay - 1.0; s_n += 1.0;
      }
    }
  }
  double v3[3] = {s_down, s_tr, s_n};
  for (int k = 0; k < 3; ++k) {
    double r = warp_sum(v3[k]);
    if ((threadIdx.x & 31) == 0) sred[k][threadIdx.x >> 5] = r;
  }
  __syncthreads();
  if (threadIdx.x < 3) {
    double s = 0.0;
    for (int g = 0; g < 8; ++g) s += sred[threadIdx.x][g];
    part[((int64_t)a * gridDim.x + blockIdx.x) * 3 + threadIdx.x] = s;
  }
}
// one warp: lane l adds parts l, l+32, ... of every column, then a fixed shuffle tree (deterministic)
__global__ void sum_parts_kernel(const double* __restrict__ part, int n_parts, int width, double* __restrict__ out, int out_len) {
  const int lane = threadIdx.x & 31;
  for (int k = 0; k < out_len; ++k) {
    double a = 0.0;
    if (k < width)
      for (int c = lane; c < n_parts; c += 32) a += part[(int64_t)c * width + k];
    a = warp_sum(a);
    if (lane == 0) out[k] = a;
  }
}

// =================================================================================================
// a11: view extraction
// =================================================================================================
struct SelectSmem {
  KeyIdx out[1024];
  int hist[256];
  int scan[2 * kBig];
};
// dynamic shared memory of the tile-wide selects: the 32-bit key cache of cta_select_k (one word per tile slot), when a
// tile fits next to the static buffers (50,000 slots: 200,000 of the 232,448 bytes an sm_100 CTA may use)
constexpr size_t kSelectStaticBytes = sizeof(SelectSmem) + 256;
inline size_t select_cache_bytes(int m_tile) {
  const size_t want = (size_t)m_tile * sizeof(uint32_t);
  return want + kSelectStaticBytes <= 232448 ? want : 0;
}
template <typename Kern>
cudaError_t select_cache_attr(Kern kern, size_t bytes) {
  return bytes > 48 * 1024 - kSelectStaticBytes ? gcs_smem_attr_once((const void*)kern, (int)bytes) : cudaSuccess;
}

// Functional recency inflation (primitive_map.py:1400-1484 applied to a copy): the gathered Lambda / theta of a valid
// slot are scaled by its decay exactly as recency_inflate_kernel scales them in place; the selection key (weight) and
// every other field are untouched by the inflation, so view(inflate(map)) == this kernel on the un-inflated map.
struct InflateArg {
  int on;
  long long scan_seq;
  double lam, min_scale;
};
// Selection (1024 threads, one CTA per stencil tile): the first m_view slots of the stable sort by weight -> V.candidate_slots.
__global__ void __launch_bounds__(kBig) map_view_kernel(gcs_atlas A, TileList T, int m_view, gcs_map_view V, int use_cache) {
  __shared__ SelectSmem sm;
  extern __shared__ uint32_t key_cache[];
  const int a = blockIdx.x;
  const int ti = T.index[a];
  const int M = A.m_tile;
  const int64_t base = (int64_t)(ti < 0 ? 0 : ti) * M;
  auto key = [&](int s) -> unsigned long long {
    // both loads are issued unconditionally (the weight of an empty slot is simply not used): a dependent chain
    // valid -> weight doubles the exposed latency of every item, and the select sweeps are latency-bound
    const uint8_t v = A.valid[base + s];
    const double w = A.weights[base + s];
    const double score = (ti >= 0 && v) ? w : -1e30;
    return f64_orderable(-score);  // ascending sort of -score (primitive_map.py:316-320)
  };
  // a tile with fewer than m_view valid slots: all of them by weight, then invalid slots in index order
  auto invalid = [&](int s) -> bool { return !(ti >= 0 && A.valid[base + s]); };
  if (!cta_select_max_sentinel(M, m_view, key, invalid, f64_orderable(1e30), sm.out, sm.scan))
    cta_select_k(M, m_view, key, sm.out, sm.hist, sm.scan, use_cache ? key_cache : nullptr);
  for (int j = threadIdx.x; j < m_view; j += kBig) V.candidate_slots[a * m_view + j] = sm.out[j].idx;
}

// Gather of the selected slots (128-thread blocks, 255 registers available: the 3x3 solve and inverse of an entry stay in
// registers; inside the 1024-thread selection kernel they spilled 256 bytes): moments of every view entry, with the
// recency inflation applied functionally when asked for.
__global__ void __launch_bounds__(128) map_view_gather_kernel(gcs_atlas A, TileList T, int m_view, double eps_lift, double eps_mass,
                                                              gcs_map_view V, int32_t* __restrict__ n_valid_out, InflateArg I) {
  const int a = blockIdx.y;
  const int ti = T.index[a];
  const int64_t base = (int64_t)(ti < 0 ? 0 : ti) * A.m_tile;
  const int j = blockIdx.x * 128 + threadIdx.x;
  bool v = false;
  if (j < m_view) {
    const int r = a * m_view + j;
    const int slot = V.candidate_slots[r];
    const int64_t o = base + slot;
    Mat3 L;
    double th[3] = {0, 0, 0}, et[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0}, w = 0.0, col[3] = {0.5, 0.5, 0.5};
    long long pid = 0, last = 0;
    for (int k = 0; k < 9; ++k) L.m[k] = 0.0;
    if (ti >= 0) {
      load_mat3(A.Lambdas + 9 * o, L);
      for (int k = 0; k < 3; ++k) { th[k] = A.thetas[3 * o + k]; col[k] = A.rgb[3 * o + k]; }
      for (int k = 0; k < 9; ++k) et[k] = A.etas[9 * o + k];
      w = A.weights[o]; pid = A.primitive_ids[o]; last = A.last_supported_scan_seq[o]; v = A.valid[o] != 0;
      if (I.on && v) {
        long long dt = I.scan_seq - last;
        if (dt < 0) dt = 0;
        const double decay = fmin(fmax(exp(-I.lam * (double)dt), I.min_scale), 1.0);
        for (int k = 0; k < 9; ++k) L.m[k] *= decay;
        for (int k = 0; k < 3; ++k) th[k] *= decay;
      }
    }
    L(0, 0) += eps_lift; L(1, 1) += eps_lift; L(2, 2) += eps_lift;
    double mu[3];
    mat3_solve(L, th, mu);
    Mat3 S = mat3_inv(L);
    const double es[3] = {et[0] + et[3] + et[6], et[1] + et[4] + et[7], et[2] + et[5] + et[8]};
    const double kap = sqrt(es[0] * es[0] + es[1] * es[1] + es[2] * es[2]);
    V.candidate_tile_ids[r] = T.id[a];
    V.valid[r] = v ? 1 : 0;
    for (int k = 0; k < 3; ++k) {
      V.positions[3 * r + k] = mu[k];
      V.directions[3 * r + k] = es[k] / (kap + eps_mass);
      V.colors[3 * r + k] = col[k];
    }
    for (int k = 0; k < 9; ++k) { V.covariances[9 * r + k] = S.m[k]; V.etas[9 * r + k] = et[k]; }
    V.kappas[r] = kap; V.weights[r] = w; V.primitive_ids[r] = pid; V.last_supported_scan_seq[r] = last;
  }
  const unsigned nv = __popc(__ballot_sync(0xffffffffu, v));
  if ((threadIdx.x & 31) == 0 && nv) atomicAdd(n_valid_out, (int)nv);
}

// =================================================================================================
// a12: association  (pair cost: gcs_assoc.cuh)
// =================================================================================================
struct AssocWs {
  double* mpos;     // (N,3)
  double* mdir;     // (N,3)
  double* mkap;     // (N)
  int8_t* stencil;  // (N, n_st)  view tile index or -1
  double* vAk;      // (P) A_vmf(kappa) of every view entry: one evaluation per scan instead of one per candidate pair
  // the view regrouped for the top-K scan (shared by all units): per tile the valid entries in Morton order, cut into
  // groups of 32 with their bounding boxes
  double* gpos;     // (P,3) positions in group order
  uint16_t* goff;   // (P)   offset within the tile of the entry at this sorted slot
  uint8_t* gval;    // (P)   validity in group order (invalid entries last)
  double* gbox;     // (n_tiles, ceil(m_view / 32), 6) min xyz, max xyz of the valid entries of a group (empty: min > max)
};

// unit u (hypothesis) of the stacked work arrays; the view and its vAk are shared by all units
__device__ __forceinline__ AssocWs assoc_ws_unit(AssocWs W, int64_t u, int N, int n_st) {
  W.mpos += u * 3 * N; W.mdir += u * 3 * N; W.mkap += u * N; W.stencil += u * N * n_st;
  return W;
}
struct StencilOffsets {   // kernel parameter (no host -> device copy: the call sequence can be captured in a CUDA graph)
  int8_t dq[64], dr[64], dz[64];
};

// blockIdx.y = unit
__global__ void __launch_bounds__(128) assoc_prepare_kernel(gcs_meas_batch B, int N, TileList T, gcs_assoc_cfg cfg,
                                                            int n_st, StencilOffsets SO,
                                                            AssocWs W, gcs_map_view V, int n_pool, int row_blocks) {
  (void)V; (void)n_pool; (void)row_blocks;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  B = meas_batch_unit(B, blockIdx.y);
  W = assoc_ws_unit(W, blockIdx.y, N, n_st);
  const int8_t *dq = SO.dq, *dr = SO.dr, *dz = SO.dz;
  double mu[3], dir[3], kap;
  meas_row_moments(B, i, cfg.eps_lift, cfg.eps_mass, mu, dir, &kap);
  for (int k = 0; k < 3; ++k) { W.mpos[3 * i + k] = mu[k]; W.mdir[3 * i + k] = dir[k]; }
  W.mkap[i] = kap;
  long long c[3];
  tile_cell(mu, fmax(cfg.h_tile, 1e-12), c);
  for (int s = 0; s < n_st; ++s) {
    const long long tid = pack_tile_id(c[0] + dq[s], c[1] + dr[s], c[2] + dz[s]);
    int idx = -1;
    for (int a = 0; a < T.n; ++a)
      if (T.id[a] == tid) { idx = a; break; }  // argmax of the equality mask = first match
    W.stencil[i * n_st + s] = (int8_t)idx;
  }
}

// The view regrouped for the top-K scan: one CTA per view tile sorts the tile's valid entries along a Morton curve through
// their bounding box (30-bit keys, bitonic sort in shared memory; tiles above 1024 entries keep their order), writes the
// positions / original offsets / validity in that order and the bounding box of every group of 32 consecutive entries.
// A measurement row then tests 32 boxes per tile with one instruction stream and scans only the groups whose box comes
// within its current K-th best cost, nearest box first -- instead of the five-flop bound of every one of the 7 x 1024
// candidates.  Any order is correct (ties are broken on the ORIGINAL offset); the order only decides how tight the boxes are.
__device__ __forceinline__ unsigned morton_spread10(unsigned x) {
  x &= 0x3ffu;
  x = (x | (x << 16)) & 0x030000ffu;
  x = (x | (x << 8)) & 0x0300f00fu;
  x = (x | (x << 4)) & 0x030c30c3u;
  x = (x | (x << 2)) & 0x09249249u;
  return x;
}

__global__ void __launch_bounds__(1024) assoc_view_groups_kernel(gcs_map_view V, int m_view, AssocWs W) {
  __shared__ KeyIdx srt[1024];
  __shared__ double red[6][32];
  __shared__ double bb[6];
  const int t = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const size_t t0 = (size_t)t * m_view;
  const bool sortable = m_view <= 1024;
  // bounding box of the tile's valid entries
  double lo[3] = {1e300, 1e300, 1e300}, hi[3] = {-1e300, -1e300, -1e300};
  for (int r = tid; r < m_view; r += 1024) {
    if (V.valid[t0 + r]) {
#pragma unroll
      for (int k = 0; k < 3; ++k) { const double x = V.positions[3 * (t0 + r) + k]; lo[k] = fmin(lo[k], x); hi[k] = fmax(hi[k], x); }
    }
  }
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    for (int o = 16; o > 0; o >>= 1) {
      lo[k] = fmin(lo[k], __shfl_xor_sync(0xffffffffu, lo[k], o));
      hi[k] = fmax(hi[k], __shfl_xor_sync(0xffffffffu, hi[k], o));
    }
    if (lane == 0) { red[k][warp] = lo[k]; red[3 + k][warp] = hi[k]; }
  }
  __syncthreads();
  if (tid < 6) {
    double a = red[tid][0];
    for (int w = 1; w < 32; ++w) a = tid < 3 ? fmin(a, red[tid][w]) : fmax(a, red[tid][w]);
    bb[tid] = a;
  }
  __syncthreads();
  if (sortable) {
    int n_pow2 = 32;
    while (n_pow2 < m_view) n_pow2 <<= 1;
    if (tid < n_pow2) {
      KeyIdx e;
      e.key = ~0ull; e.idx = tid; e.pad = 0;      // invalid entries and the padding sort last, in index order
      if (tid < m_view && V.valid[t0 + tid]) {
        unsigned q[3];
#pragma unroll
        for (int k = 0; k < 3; ++k) {
          const double ext = bb[3 + k] - bb[k];
          const double u = ext > 0.0 ? (V.positions[3 * (t0 + tid) + k] - bb[k]) / ext : 0.0;
          q[k] = (unsigned)fmin(fmax(u * 1023.0, 0.0), 1023.0);
        }
        e.key = (unsigned long long)(morton_spread10(q[0]) | (morton_spread10(q[1]) << 1) | (morton_spread10(q[2]) << 2));
      }
      srt[tid] = e;
    }
    cta_bitonic_sort(srt, n_pow2);
  }
  // A_vmf(kappa) of every view entry: one evaluation per view instead of one per candidate pair
  for (int r = tid; r < m_view; r += 1024) W.vAk[t0 + r] = A_vmf(fmax(V.kappas[t0 + r], 1e-12), 1e-12);
  const int G = (m_view + 31) / 32;
  for (int g = warp; g < G; g += 32) {
    const int r = g * 32 + lane;
    int src = r;
    bool v = false;
    double p[3] = {0.0, 0.0, 0.0};
    if (r < m_view) {
      if (sortable) src = srt[r].idx;
      v = src < m_view && V.valid[t0 + src] != 0;
      if (src >= m_view) src = 0;
#pragma unroll
      for (int k = 0; k < 3; ++k) p[k] = V.positions[3 * (t0 + src) + k];
#pragma unroll
      for (int k = 0; k < 3; ++k) W.gpos[3 * (t0 + r) + k] = p[k];
      W.goff[t0 + r] = (uint16_t)src;
      W.gval[t0 + r] = v ? 1 : 0;
    }
    double glo[3], ghi[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      glo[k] = v ? p[k] : 1e300; ghi[k] = v ? p[k] : -1e300;
      for (int o = 16; o > 0; o >>= 1) {
        glo[k] = fmin(glo[k], __shfl_xor_sync(0xffffffffu, glo[k], o));
        ghi[k] = fmax(ghi[k], __shfl_xor_sync(0xffffffffu, ghi[k], o));
      }
    }
    const double out6 = lane == 0 ? glo[0] : lane == 1 ? glo[1] : lane == 2 ? glo[2] : lane == 3 ? ghi[0] : lane == 4 ? ghi[1] : ghi[2];
    if (lane < 6) W.gbox[((size_t)t * G + g) * 6 + lane] = out6;
  }
}

// First K of the row-wise stable sort by cost over the 7 x 1024 candidate pool (primitive_association.py:367-376).
//
// PERSISTENT kernel, one CTA per SM, kTopkWarps warps.  The positions of ALL view tiles (7 x 24 KB at the reference
// budget) are staged in shared memory ONCE per CTA by the Tensor Memory Accelerator -- one cp.async.bulk.tensor box per
// tile, each completing on its own mbarrier, so a warp starts scanning a tile as soon as that tile has landed -- and
// every warp then pulls (unit, row) work items from a global counter until none is left: staging traffic no longer
// grows with the number of rows or hypotheses (it was 175 KB per 4 rows), no CTA-wide barrier sits between tiles, and a
// dense row delays only its own warp.  Views the TMA box cannot describe (m_view not a multiple of 256, unaligned
// base, more tiles than shared memory holds) are read through the same pointers straight from global memory / L1.
//
// Per row (one warp):
//   cost = |dp|^2 + beta * d_dir with d_dir in [0, 1], so |dp|^2 is a lower bound that costs five flops.  T is an upper
//   bound of the row's K-th best (cost, j) (the K-th smallest of the lanes' current best entries: K distinct candidates
//   at or before it); a candidate whose (bound, j) lies beyond (T, Tj) cannot be among the first K and is skipped --
//   its three log / sinh / exp evaluations, or, for the many rows whose stencil misses the view (every cost 1e12, the
//   first K offsets win), everything after the first few candidates.  Survivors are queued and evaluated 32 at a time
//   (all lanes busy), which leaves a few hundred of the 7,168 exact costs per row.  The row starts with the tile at the
//   centre of its own stencil: its candidates tighten T at once and the outer tiles are then pruned almost entirely.
//   Ties keep the smaller j = stencil position * m_view + offset whatever the processing order, so the selection is the
//   reference's whichever warp takes the row.
constexpr int kTopkWarps = 16;
constexpr int kTopkMaxTiles = 16;

// Lane holding the smallest (cost, j) among the `alive` lanes of the warp -- costs are non-negative float64, whose bit
// patterns order like the numbers -- with three 32-bit REDUX.MIN steps (high word, low word, j) instead of a five-level
// shuffle butterfly over three registers.  *key / *j receive the winning pair in every lane.
__device__ __forceinline__ int warp_argmin_cost_j(long long cost_bits, int j, bool alive, int lane, unsigned long long* key,
                                                  int* jmin) {
  const unsigned hi = alive ? (unsigned)((unsigned long long)cost_bits >> 32) : 0xffffffffu;
  const unsigned mhi = __reduce_min_sync(0xffffffffu, hi);
  const bool c1 = alive && hi == mhi;
  const unsigned lo = c1 ? (unsigned)(unsigned long long)cost_bits : 0xffffffffu;
  const unsigned mlo = __reduce_min_sync(0xffffffffu, lo);
  const bool c2 = c1 && lo == mlo;
  const unsigned jj = c2 ? (unsigned)j : 0xffffffffu;
  const unsigned mj = __reduce_min_sync(0xffffffffu, jj);
  const unsigned win = __ballot_sync(0xffffffffu, c2 && jj == mj);
  *key = ((unsigned long long)mhi << 32) | mlo;
  *jmin = (int)mj;
  (void)lane;
  return __ffs(win) - 1;
}
struct TopkShared {               // behind the staged tiles in dynamic shared memory
  unsigned long long bar[kTopkMaxTiles];
  int queue[kTopkWarps][64];
};
// staged per view entry: position (24 B), validity (1 B), original offset (2 B); per group of 32 entries a box (48 B)
inline size_t topk_val_bytes(int n_tiles, int m_view) { return ((size_t)n_tiles * m_view + 127) & ~(size_t)127; }
inline size_t topk_box_bytes(int n_tiles, int m_view) { return (size_t)n_tiles * ((m_view + 31) / 32) * 48; }
inline size_t topk_smem_bytes(int n_tiles, int m_view, bool staged) {
  return (staged ? (size_t)n_tiles * m_view * 24 + 3 * topk_val_bytes(n_tiles, m_view) + topk_box_bytes(n_tiles, m_view) : 0) + 128 +
         sizeof(TopkShared);
}

template <int K>
__global__ void __launch_bounds__(32 * kTopkWarps, 1) assoc_topk_kernel(gcs_meas_batch B0, int N, int n_units, gcs_map_view V,
                                                                        int m_view, int n_st, int n_view_tiles, AssocWs W0,
                                                                        gcs_assoc_cfg cfg, gcs_assoc_result R0,
                                                                        int* __restrict__ row_counter,
                                                                        const __grid_constant__ CUtensorMap tmap, int staged) {
  extern __shared__ __align__(128) unsigned char topk_smem[];
  const int n_grp = (m_view + 31) / 32;          // groups of 32 sorted entries per tile
  const size_t pos_bytes = staged ? (size_t)n_view_tiles * m_view * 24 : 0;
  const size_t val_bytes = staged ? (((size_t)n_view_tiles * m_view + 127) & ~(size_t)127) : 0;
  const size_t box_bytes = staged ? (size_t)n_view_tiles * n_grp * 48 : 0;
  double* s_pos = reinterpret_cast<double*>(topk_smem);
  double* s_box = reinterpret_cast<double*>(topk_smem + pos_bytes);
  uint16_t* s_off = reinterpret_cast<uint16_t*>(topk_smem + pos_bytes + box_bytes);
  uint8_t* s_val = topk_smem + pos_bytes + box_bytes + 2 * val_bytes;
  TopkShared& S = *reinterpret_cast<TopkShared*>(topk_smem + pos_bytes + box_bytes + 3 * val_bytes);
  const int wq = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint64_t* bars = reinterpret_cast<uint64_t*>(S.bar);
  if (staged) {
    if (threadIdx.x == 0) {
      tma::prefetch_map(&tmap);
      for (int t = 0; t < n_view_tiles; ++t) tc::mbar_init(&bars[t], 1);
      tc::mbar_init_fence();
      const int rows_per_tile = (3 * m_view) >> 8;     // 256 float64 per tensor row
      for (int t = 0; t < n_view_tiles; ++t) {
        tma::mbar_expect_tx(&bars[t], (uint32_t)(24 * m_view));
        tma::load_2d(s_pos + (size_t)t * 3 * m_view, &tmap, 0, t * rows_per_tile, &bars[t]);
      }
    }
    for (int e = threadIdx.x; e < n_view_tiles * m_view; e += 32 * kTopkWarps) { s_val[e] = W0.gval[e]; s_off[e] = W0.goff[e]; }
    for (int e = threadIdx.x; e < n_view_tiles * n_grp * 6; e += 32 * kTopkWarps) s_box[e] = W0.gbox[e];
    __syncthreads();   // barriers initialised, validity bytes / offsets / group boxes staged
  }
  int* queue = S.queue[wq];
  const long long total_rows = (long long)n_units * N;
  unsigned tiles_seen = 0;      // tiles whose mbarrier this warp has already seen complete
  const bool prune = cfg.beta >= 0.0;

  // the work item of the NEXT row is requested while the current row is processed: the counter's round trip to L2 was
  // a serial 4 % of the kernel
  // the work item of the NEXT row is requested while the current row is processed (the counter's round trip to L2 is
  // serial otherwise).  Rows are taken one at a time: chunks of four were measured 30 % slower (heavy rows -- the camera
  // features near the pose -- are neighbours, and a single hypothesis has too few chunks for 2,368 warps)
  int next_row = 0;
  if (lane == 0) next_row = atomicAdd(row_counter, 1);
  for (;;) {
    const int row = __shfl_sync(0xffffffffu, next_row, 0);
    if (row >= total_rows) break;
    if (lane == 0) next_row = atomicAdd(row_counter, 1);
    const int u = row / N, i = row - u * N;
    const gcs_meas_batch B = meas_batch_unit(B0, u);
    const AssocWs W = assoc_ws_unit(W0, u, N, n_st);
    const gcs_assoc_result R = assoc_result_unit(R0, u, N, K);
    const double mp[3] = {W.mpos[3 * i], W.mpos[3 * i + 1], W.mpos[3 * i + 2]};
    const double md[3] = {W.mdir[3 * i], W.mdir[3 * i + 1], W.mdir[3 * i + 2]};
    const double mk = W.mkap[i];
    const int8_t* my_st_mem = W.stencil + (size_t)i * n_st;
    // the row's stencil (view tile index per stencil position, -1: not in the view) in a register for n_st <= 8
    unsigned long long stw = 0ull;
    for (int q = 0; q < n_st && q < 8; ++q) stw |= (unsigned long long)(uint8_t)my_st_mem[q] << (8 * q);
    auto my_st = [&](int q) -> int { return q < 8 ? (int)(int8_t)(stw >> (8 * q)) : (int)my_st_mem[q]; };
    const bool mvalid = B.valid[i] != 0;
    // A row whose stencil has no tile in the view sees the cost 1e12 for every candidate: the first K of the stable
    // sort are offsets 0..K-1 of stencil position 0.  Most surfels of a scan lie further from the pose than the view
    // reaches, so this is the common case.
    bool any_tile = false;
    for (int q = 0; q < n_st; ++q) any_tile |= my_st(q) >= 0;
    if (!any_tile) {
      if (lane < K) {
        const int v = mvalid ? lane : 0;   // invalid measurement rows point at pool entry 0 (:379)
        R.candidate_pool_indices[i * K + lane] = v;
        R.candidate_tile_ids[i * K + lane] = V.candidate_tile_ids[v];
        R.candidate_slots[i * K + lane] = (long long)V.candidate_slots[v];
      }
      continue;
    }
    const double A_k1 = A_vmf(fmax(mk, 1e-12), 1e-12);
    double bc[K];
    int bj[K];
#pragma unroll
    for (int k = 0; k < K; ++k) { bc[k] = 1.0e300; bj[k] = 0x7fffffff; }
    int qn = 0;
    double T = 1.0e300;   // (T, Tj): K distinct candidates at or before this (cost, j) are already held
    int Tj = 0x7fffffff;
    // One loop over [view tiles in the row's order | stencil positions without a tile | drain]: a single copy of the
    // exact-cost code in the instruction stream (three inlined copies thrashed the instruction cache: 16 warps, each
    // somewhere else in 80 KB of code).
    const int c0 = my_st(n_st / 2);
    const int t_first = (c0 >= 0 && c0 < n_view_tiles) ? c0 : 0;
    const int n_seg = n_view_tiles + n_st;
    for (int g = 0; g <= n_seg; ++g) {
      const bool drain = g == n_seg;
      bool present = false;
      int s = 0;
      const double* tile_pos = nullptr;
      const uint8_t* tile_valid = nullptr;
      const uint16_t* tile_off = nullptr;
      const double* tile_box = nullptr;
      if (g < n_view_tiles) {
        const int t = g == 0 ? t_first : (g <= t_first ? g - 1 : g);
        s = -1;
        for (int q = 0; q < n_st; ++q)
          if (my_st(q) == t) { s = q; break; }
        if (s < 0) continue;
        present = true;
        if (staged) {
          if (!((tiles_seen >> t) & 1u)) { tc::mbar_wait(&bars[t], 0); tiles_seen |= 1u << t; }
          tile_pos = s_pos + (size_t)t * 3 * m_view;
          tile_valid = s_val + (size_t)t * m_view;
          tile_off = s_off + (size_t)t * m_view;
          tile_box = s_box + (size_t)t * n_grp * 6;
        } else {
          tile_pos = W.gpos + (size_t)t * 3 * m_view;
          tile_valid = W.gval + (size_t)t * m_view;
          tile_off = W.goff + (size_t)t * m_view;
          tile_box = W.gbox + (size_t)t * n_grp * 6;
        }
      } else if (!drain) {
        // stencil positions whose tile is not in the view carry cost 1e12 for every offset: they only matter while
        // fewer than K better candidates exist
        s = g - n_view_tiles;
        if (my_st(s) >= 0) continue;
        if (prune && (1e12 > T || (1e12 == T && s * m_view > Tj))) continue;
      } else if (qn == 0) {
        break;
      }
      const int jbase = s * m_view;
      // Steps of 32 candidates (one per lane) from one of three sources, then ONE copy of the queue / exact-cost code:
      //   tile in view: the groups of the tile, nearest bounding box first, until the nearest remaining box lies beyond T;
      //   stencil position without a tile: offsets in order (cost 1e12 each) while they can still matter;
      //   drain: nothing new, the queue is emptied.
      int step = 0, g0 = 0;
      unsigned remaining = 0u;      // groups of the current chunk of 32 (bit = lane holding the group's bound) still to scan
      unsigned lbt_hi = 0xffffffffu;   // this lane's group bound: high word of |dp to box|^2 rounded DOWN (a valid bound, and
                                       // 32-bit keys order the groups exactly)
      for (;;) {
        bool pass = false;
        int j = 0;
        if (present) {
          if (remaining == 0u) {
            if (g0 >= n_grp) break;
            const int gi = g0 + lane;
            lbt_hi = 0xffffffffu;
            if (gi < n_grp) {
              const double* bx = tile_box + 6 * gi;
              if (bx[0] <= bx[3]) {
                const double d0 = fmax(fmax(bx[0] - mp[0], mp[0] - bx[3]), 0.0), d1 = fmax(fmax(bx[1] - mp[1], mp[1] - bx[4]), 0.0),
                             d2 = fmax(fmax(bx[2] - mp[2], mp[2] - bx[5]), 0.0);
                // (1 - 2^-40): the box distance is exact-arithmetic <= every |dp|^2 of the group, the rounding of either is not
                const double lbg = (d0 * d0 + d1 * d1 + d2 * d2) * (1.0 - 9.094947017729282e-13);
                lbt_hi = (unsigned)((unsigned long long)__double_as_longlong(lbg) >> 32);
              }
            }
            remaining = __ballot_sync(0xffffffffu, lbt_hi != 0xffffffffu);
            g0 += 32;
            if (remaining == 0u) continue;
          }
          const unsigned mine = ((remaining >> lane) & 1u) ? lbt_hi : 0xffffffffu;
          const unsigned best = __reduce_min_sync(0xffffffffu, mine);
          const int gsel = __ffs(__ballot_sync(0xffffffffu, mine == best)) - 1;
          // every remaining group of the chunk is at least this far: beyond T none of them holds one of the first K
          if (prune && __longlong_as_double((long long)((unsigned long long)best << 32)) > T) { remaining = 0u; continue; }
          remaining &= ~(1u << gsel);
          const int r = (g0 - 32 + gsel) * 32 + lane;
          if (r < m_view && tile_valid[r]) {
            j = jbase + (int)tile_off[r];
            const double d0 = mp[0] - tile_pos[3 * r], d1 = mp[1] - tile_pos[3 * r + 1], d2 = mp[2] - tile_pos[3 * r + 2];
            const double lb = d0 * d0 + d1 * d1 + d2 * d2;
            pass = !prune || !(lb > T || (lb == T && j > Tj));   // lb <= cost: beyond (T, Tj) it cannot be among the first K
          }
        } else if (!drain) {
          const int base = 32 * step;
          if (base >= m_view) break;
          if (prune && (1e12 > T || (1e12 == T && jbase + base > Tj))) break;
          const int off = base + lane;
          j = jbase + off;
          pass = off < m_view && (!prune || !(1e12 > T || (1e12 == T && j > Tj)));
          ++step;
        } else if (step++ > 0 || qn == 0) {
          break;
        }
        const unsigned pm = __ballot_sync(0xffffffffu, pass);
        if (pm != 0u) {
          if (pass) queue[qn + __popc(pm & ((1u << lane) - 1u))] = j;
          qn += __popc(pm);
          __syncwarp();
        }
        // exact costs as soon as a full warp of candidates waits -- or, while no bound exists yet, as soon as K wait: the
        // first K exact costs are what lets the boxes prune
        if (qn >= 32 || (drain && qn > 0) || (T >= 1.0e299 && qn >= K)) {
          const int n_take = qn < 32 ? qn : 32;
          // ---- exact costs of n_take queued candidates, one per lane; each lane keeps its K best in (cost, j) order
          bool head_changed = false;
          if (lane < n_take) {
            const int jq = queue[lane];
            const int sq = jq / m_view, offq = jq - sq * m_view;
            const int tix = my_st(sq);
            const int v = (tix < 0 ? 0 : tix) * m_view + offq;
            double c = 1e12;
            if (tix >= 0 && V.valid[v])
              c = pair_cost_pre(mp, md, mk, A_k1, V.positions + 3 * v, V.directions + 3 * v, V.kappas[v], W.vAk[v], cfg.beta);
            if (c < bc[K - 1] || (c == bc[K - 1] && jq < bj[K - 1])) {
              bc[K - 1] = c; bj[K - 1] = jq;
#pragma unroll
              for (int k = K - 1; k > 0; --k) {
                if (bc[k] < bc[k - 1] || (bc[k] == bc[k - 1] && bj[k] < bj[k - 1])) {
                  const double tc_ = bc[k]; bc[k] = bc[k - 1]; bc[k - 1] = tc_;
                  const int tj = bj[k]; bj[k] = bj[k - 1]; bj[k - 1] = tj;
                }
              }
              head_changed = bj[0] == jq;
            }
          }
          // (T, Tj) = K-th smallest of the 32 lane heads in (cost, j) order; it can only move when a head moved
          if (__any_sync(0xffffffffu, head_changed)) {
            bool alive = true;
            unsigned long long kkey = 0ull;
            int kj = 0;
#pragma unroll 1
            for (int r = 0; r < K; ++r) {
              const int wl = warp_argmin_cost_j(__double_as_longlong(bc[0]), bj[0], alive, lane, &kkey, &kj);
              if (lane == wl) alive = false;
            }
            T = __longlong_as_double((long long)kkey);
            Tj = kj;
          }
          __syncwarp();
          qn -= n_take;
          int moved = 0;
          if (lane < qn) moved = queue[32 + lane];   // at most 31 left over
          __syncwarp();
          if (lane < qn) queue[lane] = moved;
          __syncwarp();
        }
      }   // steps of this segment
    }
    // K rounds of arg-min over the lanes' heads by (cost, j); the winner pops its list.  Lane r keeps the r-th winner and
    // the K lanes write the row's outputs together (the gathers of the tile ids / slots were K serial round trips)
    int win_j = 0;
#pragma unroll 1
    for (int r = 0; r < K; ++r) {
      unsigned long long hkey;
      int hj;
      const int hl = warp_argmin_cost_j(__double_as_longlong(bc[0]), bj[0], true, lane, &hkey, &hj);
      if (lane == hl) {
#pragma unroll
        for (int k = 0; k < K - 1; ++k) { bc[k] = bc[k + 1]; bj[k] = bj[k + 1]; }
        bc[K - 1] = 1.0e300; bj[K - 1] = 0x7fffffff;
      }
      if (lane == r) win_j = hj;
    }
    if (lane < K) {
      const int s = win_j / m_view, off = win_j - s * m_view;
      const int tix = my_st(s);
      int v = (tix < 0 ? 0 : tix) * m_view + off;
      if (!mvalid) v = 0;  // invalid measurement rows point at pool entry 0 (:379)
      R.candidate_pool_indices[i * K + lane] = v;
      R.candidate_tile_ids[i * K + lane] = V.candidate_tile_ids[v];
      R.candidate_slots[i * K + lane] = (long long)V.candidate_slots[v];
    }
    __syncwarp();
  }
  // a CTA must not retire while the TMA may still write into its shared memory
  if (staged && wq == 0)
    for (int t = 0; t < n_view_tiles; ++t) tc::mbar_wait(&bars[t], 0);
}

// single CTA: cost of the selected candidates, recency term, row-min shift, unbalanced Sinkhorn, certificates
// Unbalanced Sinkhorn on the (N, K) sparse costs (primitive_association.py:105-138, :379-470) as a thread-block CLUSTER of
// eight CTAs: every CTA owns 256 measurement rows (one per thread), the K column sums that couple all rows -- the
// reference's v is one (K,) vector shared by every row -- are exchanged through distributed shared memory once per
// iteration: each CTA stores its K partial sums into every CTA's shared memory, one cluster barrier, every CTA adds the
// eight partial vectors in rank order (same order everywhere: identical v in every CTA, bit-identical reruns).  The
// exchange buffers alternate with the iteration parity, so one barrier per iteration suffices.  A single CTA spent
// 7 us per iteration on the float64 pow() of 1,536 rows (one SM's FP64 pipe); eight SMs share that work.
constexpr int kSkCtas = 8;
constexpr int kSkThreads = 256;
// K_ASSOC values this build instantiates (the reference default is 8, config k_assoc: primitive_association.py:60)
#define GCS_K_ASSOC_DISPATCH(k_, ...)                         \
  switch (k_) {                                               \
    case 4: { constexpr int KK = 4; __VA_ARGS__; } break;     \
    case 8: { constexpr int KK = 8; __VA_ARGS__; } break;     \
    case 16: { constexpr int KK = 16; __VA_ARGS__; } break;   \
    default: break;                                           \
  }
#define GCS_K_ASSOC_OK(k_) ((k_) == 4 || (k_) == 8 || (k_) == 16)
constexpr int kSkMaxVals = 24;   // widest row of a cluster_sum: 6 + K certificate columns at K_ASSOC = 16

// Sum of NV (<= kSkMaxVals) values per row over all rows; every thread of every CTA gets post(k, total_k).
// The rows are cut into kSkCtas = 8 VIRTUAL CTAs of 256 rows; a real CTA of a cluster of C = 8 / RPT CTAs carries RPT of
// them (RPT rows per thread).  Fixed order whatever C is: shuffle tree per (virtual CTA, warp), warps in index order,
// virtual CTAs in index order -- the batched launch (small clusters, many hypotheses in flight) and the single-scan
// launch (eight SMs on one hypothesis) give bit-identical results.  Thread (r, q, k) adds up value k of this CTA's
// virtual CTA q over the warps and PUSHES it into real CTA r's exchange buffer; after the cluster barrier (release /
// acquire) every CTA sums its local copy.  A pull (barrier, then remote loads) puts a distributed-shared-memory round
// trip behind every barrier.  `phase` alternates the exchange buffer, so that a CTA running ahead never overwrites
// values a slower one still reads.
template <int NV, int RPT, typename Post>
__device__ __forceinline__ void cluster_sum(double (&v)[RPT][NV], double (*xch)[kSkCtas][kSkMaxVals],
                                            double (*sredw)[kSkThreads / 32][kSkMaxVals], double* tot, unsigned& phase,
                                            Post post) {
  namespace cg = cooperative_groups;
  cg::cluster_group cluster = cg::this_cluster();
  constexpr int C = kSkCtas / RPT;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
#pragma unroll
  for (int q = 0; q < RPT; ++q)
#pragma unroll
    for (int k = 0; k < NV; ++k) v[q][k] = warp_sum(v[q][k]);
  __syncthreads();   // sredw / tot are free (the previous call's readers are done)
  if (lane == 0) {
#pragma unroll
    for (int q = 0; q < RPT; ++q)
#pragma unroll
      for (int k = 0; k < NV; ++k) sredw[q][warp][k] = v[q][k];
  }
  __syncthreads();
  const unsigned buf = phase & 1u;
  const unsigned my_rank = cluster.block_rank();
  if (tid < NV * kSkCtas) {       // NV * C * RPT threads
    const int k = tid % NV, rest = tid / NV, r = rest % C, q = rest / C;
    double t = 0.0;
    for (int w = 0; w < kSkThreads / 32; ++w) t += sredw[q][w][k];
    cluster.map_shared_rank(&xch[buf][my_rank * RPT + q][0], r)[k] = t;
  }
  cluster.sync();
  if (tid < NV) {
    double t = 0.0;
    for (int r = 0; r < kSkCtas; ++r) t += xch[buf][r][tid];
    tot[tid] = post(tid, t);
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < NV; ++k) v[0][k] = tot[k];
  ++phase;
}

// RPT rows per thread, cluster of 8 / RPT CTAs (set at launch): 1 for a single hypothesis, 2 or 4 when many hypotheses
// share the device (the iterations are bound by the float64 pipe of the SMs a hypothesis runs on and by the cluster
// barrier: with 64 hypotheses in flight two SMs each are enough and leave room for all of them at once)
template <int K, int RPT>
__global__ void __launch_bounds__(kSkThreads)
    assoc_sinkhorn_kernel(gcs_meas_batch B, int N, gcs_map_view V, AssocWs W, gcs_assoc_cfg cfg, gcs_assoc_result R,
                          double* __restrict__ cert, double* __restrict__ brow_ws, double* __restrict__ a_ws) {
  namespace cg = cooperative_groups;
  cg::cluster_group cluster = cg::this_cluster();
  __shared__ double xch[2][kSkCtas][kSkMaxVals];
  __shared__ double sredw[RPT][kSkThreads / 32][kSkMaxVals];
  __shared__ double tot[kSkMaxVals];
  __shared__ SelectSmem sel;
  const int tid = threadIdx.x;
  const int rank = (int)cluster.block_rank();
  {
    const int64_t un = blockIdx.y;          // unit: one cluster per hypothesis
    B = meas_batch_unit(B, un);
    W = assoc_ws_unit(W, un, N, 0);
    R = assoc_result_unit(R, un, N, K);
    cert += un * GCS_OT_NCERT;
    brow_ws += un * N * K;
    a_ws += un * N;
  }
  int row[RPT];                             // N <= kSkCtas * kSkThreads = 2048
#pragma unroll
  for (int q = 0; q < RPT; ++q) row[q] = (rank * RPT + q) * kSkThreads + tid;
  unsigned phase = 0;
  double Km[RPT][K], u[RPT], a[RPT];
  const double eps = fmax(cfg.epsilon, 1e-12);
  auto ident = [](int, double t) { return t; };
  // measurement marginal (primitive_association.py:412-424): UNIFORM valid / sum(valid), or WEIGHT_PROPORTIONAL
  // valid * weight / sum(valid * weight)
  const bool wprop = cfg.a_policy == 1;
  double nv[RPT][2], araw[RPT];
#pragma unroll
  for (int q = 0; q < RPT; ++q) {
    const bool v = row[q] < N && B.valid[row[q]];
    nv[q][0] = v ? 1.0 : 0.0;
    araw[q] = v ? (wprop ? B.weights[row[q]] : 1.0) : 0.0;
    nv[q][1] = araw[q];
  }
  cluster_sum<2, RPT>(nv, xch, sredw, tot, phase, ident);
  const double sum_valid = nv[0][0];
  const double sum_a = fmax(wprop ? nv[0][1] : sum_valid, cfg.eps_mass);
#pragma unroll
  for (int q = 0; q < RPT; ++q) {
    const int i = row[q];
    u[q] = 1.0; a[q] = 0.0;
#pragma unroll
    for (int k = 0; k < K; ++k) Km[q][k] = 0.0;
    if (i < N) {
      a[q] = araw[q] / sum_a;
      if (wprop) a_ws[i] = a[q];
      const double mp[3] = {W.mpos[3 * i], W.mpos[3 * i + 1], W.mpos[3 * i + 2]};
      const double md[3] = {W.mdir[3 * i], W.mdir[3 * i + 1], W.mdir[3 * i + 2]};
      const double mk = W.mkap[i];
      const double A_k1 = A_vmf(fmax(mk, 1e-12), 1e-12);
      double rmin = 1.0e300, bsum = 0.0, bdec[K], Cm[K];
      for (int k = 0; k < K; ++k) {
        const int v = R.candidate_pool_indices[i * K + k];
        double c = pair_cost_pre(mp, md, mk, A_k1, V.positions + 3 * v, V.directions + 3 * v, V.kappas[v], W.vAk[v], cfg.beta);
        long long dt = cfg.scan_seq - V.last_supported_scan_seq[v];
        if (dt < 0) dt = 0;
        c += cfg.epsilon * cfg.recency_decay_lambda * (double)dt;
        Cm[k] = c;
        rmin = fmin(rmin, c);
        double d = exp(-cfg.recency_decay_lambda * (double)dt);
        if (!(d > 0.0)) d = 0.0;
        bdec[k] = d; bsum += d;
      }
      for (int k = 0; k < K; ++k) {
        Cm[k] -= rmin;
        Km[q][k] = exp(-Cm[k] / eps);
        R.cost_matrix[i * K + k] = Cm[k];    // read back for the certificate sums after the iterations
        brow_ws[i * K + k] = bdec[k] / fmax(bsum, cfg.eps_mass);  // diagnostics only (:420-425)
      }
    }
  }
  double sv[K];
#pragma unroll
  for (int k = 0; k < K; ++k) sv[k] = 1.0;
  const double ua = 1.0 / (1.0 + cfg.tau_a / eps), vb = 1.0 / (1.0 + cfg.tau_b / eps);
  const double bk = 1.0 / (double)K;
  for (int it = 0; it < cfg.k_sinkhorn; ++it) {
    double ktu[RPT][K];
#pragma unroll
    for (int q = 0; q < RPT; ++q) {
      if (row[q] < N) {
        double kv = 0.0;
#pragma unroll
        for (int k = 0; k < K; ++k) kv += Km[q][k] * sv[k];
        u[q] = pow_pos(a[q] / (kv + 1e-12), ua);
#pragma unroll
        for (int k = 0; k < K; ++k) ktu[q][k] = Km[q][k] * u[q];
      } else {
#pragma unroll
        for (int k = 0; k < K; ++k) ktu[q][k] = 0.0;
      }
    }
    // one cluster barrier per iteration; the K column scalings v_k = (b / (K^T u + eps))^(1/(1+tau_b/eps)) are
    // evaluated by the K gathering threads of each CTA
    cluster_sum<K, RPT>(ktu, xch, sredw, tot, phase, [&](int, double t) { return pow_pos(bk / (t + 1e-12), vb); });
#pragma unroll
    for (int k = 0; k < K; ++k) sv[k] = ktu[0][k];
  }
  // outputs + certificate sums
  double cs[RPT][6 + K];
#pragma unroll
  for (int q = 0; q < RPT; ++q) {
#pragma unroll
    for (int k = 0; k < 6 + K; ++k) cs[q][k] = 0.0;
    const int i = row[q];
    if (i < N) {
      const bool mv = B.valid[i] != 0;
      double rowm = 0.0;
      for (int k = 0; k < K; ++k) {
        const double pi = u[q] * Km[q][k] * sv[k];
        rowm += pi; cs[q][6 + k] += pi; cs[q][4] += pi * R.cost_matrix[i * K + k];
        R.responsibilities[i * K + k] = mv ? pi : 0.0;
      }
      R.row_masses[i] = rowm;
      cs[q][0] = rowm; cs[q][1] = rowm * rowm;
      const double d = rowm - a[q];
      cs[q][2] = d * d;
      cs[q][3] = fmax(a[q] - rowm, 0.0);
      cs[q][5] = a[q] > cfg.eps_mass ? 1.0 : 0.0;
    }
  }
  cluster_sum<6 + K, RPT>(cs, xch, sredw, tot, phase, ident);
  const double S_row = cs[0][0], S_row2 = cs[0][1], S_da = cs[0][2], S_nov = cs[0][3], S_cost = cs[0][4], n_nz_a = cs[0][5];
  double db = 0.0;
  for (int k = 0; k < K; ++k) db += (cs[0][6 + k] - bk) * (cs[0][6 + k] - bk);
  // no CTA may exit while another still reads its exchange buffer; the barrier also makes the brow_ws rows written by
  // every CTA visible to rank 0, which finishes alone
  cluster.sync();
  if (rank != 0) return;
  // p95 of the diagnostic per-row recency marginal: the (total - idx)-th largest of N*K values
  const int total = N * K;
  int idx95 = (int)(0.95 * (double)total);
  if (idx95 > total - 1) idx95 = total - 1;
  const int kk = total - idx95;
  double b95 = -1.0;
  if (kk >= 1 && kk <= 1024 && total >= kk) {
    auto key = [&](int j) -> unsigned long long { return ~f64_orderable(brow_ws[j]); };  // descending
    cta_select_k(total, kk, key, sel.out, sel.hist, sel.scan);
    b95 = brow_ws[sel.out[kk - 1].idx];
  }
  int ia = (int)(0.95 * (double)N);
  if (ia > N - 1) ia = N - 1;
  double a95 = 0.0;      // WEIGHT_PROPORTIONAL: sorted(a)[ia] = the (N - ia)-th largest of the stored marginal
  if (wprop && N - ia >= 1 && N - ia <= 1024) {
    __syncthreads();
    auto keya = [&](int j) -> unsigned long long { return ~f64_orderable(a_ws[j]); };  // descending
    cta_select_k(N, N - ia, keya, sel.out, sel.hist, sel.scan);
    a95 = a_ws[sel.out[N - ia - 1].idx];
  }
  if (tid == 0) {
    const int n0 = N - (int)sum_valid;  // zeros of a sort first
    cert[GCS_OT_MARGINAL_A] = sqrt(S_da);
    cert[GCS_OT_MARGINAL_B] = sqrt(db);
    cert[GCS_OT_MASS_TOTAL] = S_row;
    cert[GCS_OT_SUM_A] = sum_a;
    cert[GCS_OT_SUM_M] = S_row;
    cert[GCS_OT_SUM_NOVEL] = S_nov;
    cert[GCS_OT_P95_A] = wprop ? a95 : ((ia < n0) ? 0.0 : 1.0 / sum_a);
    cert[GCS_OT_NONZERO_A] = wprop ? n_nz_a : (((1.0 / sum_a) > cfg.eps_mass) ? sum_valid : 0.0);
    cert[GCS_OT_B_RECENCY_P95] = b95;
    cert[GCS_OT_ESS] = S_row * S_row / (S_row2 + cfg.eps_mass);
    cert[GCS_OT_TOTAL_COST] = S_cost;
    cert[GCS_OT_SUM_M2] = S_row2;
    for (int k = GCS_OT_SUM_M2 + 1; k < GCS_OT_NCERT; ++k) cert[k] = 0.0;
  }
}

template <int K, int RPT>
static cudaError_t sinkhorn_launch(cudaStream_t st, unsigned n_units, gcs_meas_batch B, int N, gcs_map_view V, AssocWs W,
                                   gcs_assoc_cfg cfg, gcs_assoc_result R, double* cert, double* brow, double* a_ws) {
  cudaLaunchConfig_t lc;
  memset(&lc, 0, sizeof(lc));
  lc.gridDim = dim3(kSkCtas / RPT, n_units, 1);
  lc.blockDim = dim3(kSkThreads, 1, 1);
  lc.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = kSkCtas / RPT; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  lc.attrs = at;
  lc.numAttrs = 1;
  return cudaLaunchKernelEx(&lc, assoc_sinkhorn_kernel<K, RPT>, B, N, V, W, cfg, R, cert, brow, a_ws);
}

// every unit of a stacked batch := the base batch (camera slice, zero LiDAR rows); blockIdx.y = unit
__global__ void __launch_bounds__(128) batch_replicate_kernel(gcs_meas_batch S, gcs_meas_batch D, int N) {
  const int i = blockIdx.x * 128 + threadIdx.x;
  if (i >= N) return;
  D = meas_batch_unit(D, blockIdx.y);
#pragma unroll
  for (int k = 0; k < 9; ++k) { D.Lambdas[9 * i + k] = S.Lambdas[9 * i + k]; D.etas[9 * i + k] = S.etas[9 * i + k]; }
#pragma unroll
  for (int k = 0; k < 3; ++k) { D.thetas[3 * i + k] = S.thetas[3 * i + k]; D.colors[3 * i + k] = S.colors[3 * i + k]; }
  D.weights[i] = S.weights[i]; D.sources[i] = S.sources[i]; D.source_indices[i] = S.source_indices[i];
  D.valid[i] = S.valid[i]; D.timestamps[i] = S.timestamps[i];
}

// =================================================================================================
// a13: pose evidence
// =================================================================================================
// 512 threads: three rows per thread at the reference budget (1,536 rows) and 128 registers each -- the 1024-thread
// version was capped at 64 registers and spilled the 28 accumulators
constexpr int kPeThreads = 384;   // 170 registers: no spills
// kPeParts CTAs per unit (blockIdx.x = part, blockIdx.y = unit): the rows of a unit are dealt round-robin to the parts, every
// part leaves its 25 sums in `partial`, and the part that arrives last (a counter per unit, reset for the next launch) adds
// them in part order and finishes.  One CTA per unit spent 44 us in dependent gathers of the K candidates of four rows per
// thread -- a tenth of a single hypothesis' path.  The part count is a constant, so a batch and a single scan add in the
// same order.
constexpr int kPeParts = 4;
template <int K>
__global__ void __launch_bounds__(kPeThreads) pose_evidence_kernel(gcs_meas_batch B, int N, gcs_map_view V, gcs_assoc_result R,
                                                             double p0, double p1, double p2, double r0, double r1,
                                                             double r2, double eps_lift, double eps_mass,
                                                             double* __restrict__ L22, double* __restrict__ h22,
                                                             double* __restrict__ rec, const double* __restrict__ poses_dev,
                                                             const int32_t* __restrict__ n_lidar_valid, int n_camera_valid,
                                                             const int32_t* __restrict__ view_n_valid,
                                                             double* __restrict__ partial, int* __restrict__ arrived) {
  __shared__ double tot[28];
  __shared__ int s_last;
  const int tid = threadIdx.x;
  const int part = blockIdx.x;
  partial += (int64_t)blockIdx.y * kPeParts * 25;
  arrived += blockIdx.y;
  if (poses_dev) {                           // blockIdx.y = unit: its own pose, batch, association and outputs
    const int64_t u = blockIdx.y;
    p0 = poses_dev[6 * u]; p1 = poses_dev[6 * u + 1]; p2 = poses_dev[6 * u + 2];
    r0 = poses_dev[6 * u + 3]; r1 = poses_dev[6 * u + 4]; r2 = poses_dev[6 * u + 5];
    B = meas_batch_unit(B, u);
    R = assoc_result_unit(R, u, N, K);
    L22 += u * 22 * 22; h22 += u * 22; rec += u * GCS_VP_NREC;
    // The reference's early exits, taken on the device (the host cannot look before a batch of hypotheses has been
    // enqueued): an empty measurement batch or an empty view give the all-zero association of
    // primitive_association.py:275-290 and the evidence eps_lift * I, h = 0 of visual_pose_evidence.py:300-330 -- whatever
    // the association kernels computed from 1e12 costs is overwritten, so the map update of this unit fuses nothing.
    if (n_lidar_valid && view_n_valid && (view_n_valid[0] == 0 || n_camera_valid + n_lidar_valid[u] == 0)) {
      if (part != 0) return;
      for (int e = tid; e < N * K; e += kPeThreads) {
        R.responsibilities[e] = 0.0; R.cost_matrix[e] = 0.0; R.candidate_pool_indices[e] = 0;
        R.candidate_tile_ids[e] = 0; R.candidate_slots[e] = 0;
      }
      for (int e = tid; e < N; e += kPeThreads) R.row_masses[e] = 0.0;
      for (int e = tid; e < 22 * 22; e += kPeThreads) L22[e] = (e / 22 == e % 22) ? eps_lift : 0.0;
      if (tid < 22) h22[tid] = 0.0;
      for (int e = tid; e < GCS_VP_NREC; e += kPeThreads) rec[e] = 0.0;
      return;
    }
  }
  const double rv[3] = {r0, r1, r2}, tp[3] = {p0, p1, p2};
  const Mat3 Rp = so3_exp(rv);
  double acc[28];
#pragma unroll
  for (int k = 0; k < 28; ++k) acc[k] = 0.0;
  // 0..8 L_t, 9..11 h_t, 12 trans cost, 13..21 S, 22 rot cost, 23 sum row mass, 24 n valid rows
  for (int i = part * kPeThreads + tid; i < N; i += kPeParts * kPeThreads) {
    if (!B.valid[i]) continue;
    double mu[3], dir[3], kap;
    meas_row_moments(B, i, eps_lift, eps_mass, mu, dir, &kap);
    Mat3 L;
    load_mat3(B.Lambdas + 9 * i, L);
    L(0, 0) += eps_lift; L(1, 1) += eps_lift; L(2, 2) += eps_lift;
    double mw[3], dw[3];
    mat3_vec(Rp, mu, mw);
    mat3_vec(Rp, dir, dw);
    double pis = 0.0, wt[3] = {0, 0, 0};
    for (int k = 0; k < K; ++k) {
      const double pi = R.responsibilities[i * K + k];
      const int v = R.candidate_pool_indices[i * K + k];
      const double* vp = V.positions + 3 * v;
      const double* vd = V.directions + 3 * v;
      pis += pi;
      const double tg[3] = {vp[0] - mw[0], vp[1] - mw[1], vp[2] - mw[2]};
      const double rs[3] = {tg[0] - tp[0], tg[1] - tp[1], tg[2] - tp[2]};
      for (int c = 0; c < 3; ++c) wt[c] += pi * tg[c];
      double Lr[3];
      mat3_vec(L, rs, Lr);
      acc[12] += pi * (rs[0] * Lr[0] + rs[1] * Lr[1] + rs[2] * Lr[2]);
      const double w = pi * sqrt(kap * V.kappas[v] + 1e-12);
      for (int a = 0; a < 3; ++a)
        for (int c = 0; c < 3; ++c) acc[13 + 3 * a + c] += w * vd[a] * dir[c];
      acc[22] += w * (1.0 - (dw[0] * vd[0] + dw[1] * vd[1] + dw[2] * vd[2]));
    }
    for (int k = 0; k < 9; ++k) acc[k] += pis * L.m[k];
    double Lw[3];
    mat3_vec(L, wt, Lw);
    for (int c = 0; c < 3; ++c) acc[9 + c] += Lw[c];
    acc[23] += R.row_masses[i];
    acc[24] += 1.0;
  }
  // all 25 block sums behind ONE barrier pair (same order as block_sum_1024: shuffle tree per warp, warps in index
  // order); fully unrolled so that acc[] stays in registers
  __shared__ double swarp[25][kPeThreads / 32];
#pragma unroll
  for (int k = 0; k < 25; ++k) {
    const double w = warp_sum(acc[k]);
    if ((tid & 31) == 0) swarp[k][tid >> 5] = w;
  }
  __syncthreads();
  if (tid < 25) {
    double s = 0.0;
    for (int w = 0; w < kPeThreads / 32; ++w) s += swarp[tid][w];
    partial[part * 25 + tid] = s;
  }
  // last part to arrive finishes (threadfence / atomic hand-over: its reads see every part's sums)
  __threadfence();
  __syncthreads();
  if (tid == 0) {
    const int prev = atomicAdd(arrived, 1);
    s_last = prev == kPeParts - 1;
    if (s_last) *arrived = 0;          // ready for the next launch
  }
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  if (tid < 25) {
    double s = 0.0;
    for (int q = 0; q < kPeParts; ++q) s += __ldcg(partial + q * 25 + tid);
    tot[tid] = s;
  }
  __syncthreads();
  if (tid == 0) {
    Mat3 Lt, S, U, Vm;
    for (int k = 0; k < 9; ++k) { Lt.m[k] = tot[k]; S.m[k] = tot[13 + k]; }
    Lt(0, 0) += eps_lift; Lt(1, 1) += eps_lift; Lt(2, 2) += eps_lift;
    double sv[3];
    svd3(S, U, sv, Vm);
    Mat3 Vt = mat3_T(Vm);
    Mat3 Rs = mat3_mul(U, Vt);
    if (mat3_det(Rs) < 0.0) {
      for (int r = 0; r < 3; ++r) U(r, 2) = -U(r, 2);  // U diag(1,1,-1) V^T
      Rs = mat3_mul(U, Vt);
    }
    Mat3 Rd = mat3_mul(Rs, mat3_T(Rp));
    double dr[3];
    so3_log(Rd, dr);
    const double lr[3] = {sv[0] + eps_lift, sv[1] + eps_lift, sv[2] + eps_lift};  // L_rot = diag(s + eps) (quirk Q5)
    for (int k = 0; k < 9; ++k) { rec[GCS_VP_L_TRANS + k] = Lt.m[k]; rec[GCS_VP_L_ROT + k] = 0.0; rec[GCS_VP_R_SCATTER + k] = Rs.m[k]; }
    for (int k = 0; k < 3; ++k) {
      rec[GCS_VP_H_TRANS + k] = tot[9 + k];
      rec[GCS_VP_L_ROT + 4 * k] = lr[k];
      rec[GCS_VP_H_ROT + k] = lr[k] * dr[k];
      rec[GCS_VP_SVD_S + k] = sv[k];
      rec[GCS_VP_DELTA_ROT + k] = dr[k];
    }
    rec[GCS_VP_TRANS_COST] = tot[12]; rec[GCS_VP_ROT_COST] = tot[22];
    rec[GCS_VP_SUM_ROW_MASS] = tot[23]; rec[GCS_VP_N_VALID_ROWS] = tot[24];
    for (int k = GCS_VP_R_SCATTER + 9; k < GCS_VP_NREC; ++k) rec[k] = 0.0;
  }
  __syncthreads();
  for (int idx = tid; idx < 22 * 22; idx += kPeThreads) {
    const int r = idx / 22, c = idx - r * 22;
    double v = (r == c) ? eps_lift : 0.0;
    if (r < 3 && c < 3) v = rec[GCS_VP_L_TRANS + 3 * r + c];
    else if (r >= 3 && r < 6 && c >= 3 && c < 6) v = rec[GCS_VP_L_ROT + 3 * (r - 3) + (c - 3)];
    L22[idx] = v;
  }
  if (tid < 22) h22[tid] = tid < 3 ? rec[GCS_VP_H_TRANS + tid] : (tid < 6 ? rec[GCS_VP_H_ROT + tid - 3] : 0.0);
}

// =================================================================================================
// a14: map update
// =================================================================================================
struct UpdWs {
  double* Lw;       // (N,9) world-frame precision of each measurement
  double* thw;      // (N,3)
  double* etw;      // (N,9)
  long long* mtile; // (N) packed tile id of the world-frame mean
  double* novelty;  // (N)
  double* score;    // (N)
  unsigned* pkeys;  // (n_pairs) target key = active tile * m_tile + slot (n_tiles * m_tile = no target), sorted
  unsigned* pvals;  // (n_pairs) pair index, ascending within equal keys (stable sort)
  int* ins_idx;     // (T,k)
  uint8_t* ins_new; // (T,k)
  double* ins_w;    // (T,k)
  int* ins_slot;    // (T,k)
  int* n_ins;       // (T)
  double* part;     // partial stats
};

// per-measurement world transform, tile id, novelty, insertion score (pipeline.py:1248-1256, :1331-1346)
constexpr int kPrepThreads = 512;   // 128 registers per thread: the 3x3 products of a row stay in registers (1024 x 64 spilled 272 B)
__global__ void __launch_bounds__(kPrepThreads) upd_prepare_kernel(gcs_meas_batch B, int N, gcs_assoc_result R, double p0, double p1,
                                                           double p2, double r0, double r1, double r2,
                                                           gcs_map_update_cfg cfg, UpdWs W) {
  __shared__ double sred[32];
  const int tid = threadIdx.x;
  const double rv[3] = {r0, r1, r2}, tt[3] = {p0, p1, p2};
  const Mat3 Rm = so3_exp(rv);
  const Mat3 Rt = mat3_T(Rm);
  double nv = 0.0;
  for (int i = tid; i < N; i += kPrepThreads) nv += B.valid[i] ? 1.0 : 0.0;
  const double denom = fmax(block_sum_1024(nv, sred), cfg.eps_mass);
  for (int i = tid; i < N; i += kPrepThreads) {
    Mat3 L;
    load_mat3(B.Lambdas + 9 * i, L);
    Mat3 Lw = mat3_mul(mat3_mul(Rm, L), Rt);
    Mat3 Lr = L;
    Lr(0, 0) += cfg.eps_lift; Lr(1, 1) += cfg.eps_lift; Lr(2, 2) += cfg.eps_lift;
    const double th[3] = {B.thetas[3 * i], B.thetas[3 * i + 1], B.thetas[3 * i + 2]};
    double mub[3], muw[3], thw[3];
    mat3_solve(Lr, th, mub);
    mat3_vec(Rm, mub, muw);
    for (int k = 0; k < 3; ++k) muw[k] += tt[k];
    mat3_vec(Lw, muw, thw);
    for (int k = 0; k < 9; ++k) W.Lw[9 * i + k] = Lw.m[k];
    for (int k = 0; k < 3; ++k) W.thw[3 * i + k] = thw[k];
    for (int b = 0; b < 3; ++b) {
      const double e[3] = {B.etas[9 * i + 3 * b], B.etas[9 * i + 3 * b + 1], B.etas[9 * i + 3 * b + 2]};
      double ew[3];
      mat3_vec(Rm, e, ew);
      for (int k = 0; k < 3; ++k) W.etw[9 * i + 3 * b + k] = ew[k];
    }
    long long c[3];
    tile_cell(muw, fmax(cfg.h_tile, 1e-12), c);
    W.mtile[i] = pack_tile_id(c[0], c[1], c[2]);
    const double v = B.valid[i] ? 1.0 : 0.0;
    const double nov = fmax(v / denom - R.row_masses[i], 0.0);
    W.novelty[i] = nov;
    W.score[i] = nov * B.weights[i] - (1.0 - v) * 1e6;
  }
}

// Target key of every (measurement, candidate) pair: (active tile, slot), or `none` for pairs without a target.  A
// stable radix sort by that key (upd_sort_pairs_kernel, pair index as the value) then lines up every target's
// contributions in pair order: deterministic, no floating atomics.
__global__ void __launch_bounds__(256) upd_pair_keys_kernel(gcs_meas_batch B, int N, int K, gcs_assoc_result R, TileList T,
                                                            int m_tile, unsigned none, unsigned* __restrict__ keys,
                                                            unsigned* __restrict__ vals) {
  const int p = blockIdx.x * 256 + threadIdx.x;
  if (p >= N * K) return;
  unsigned key = none;
  const int i = p / K;
  if (B.valid[i]) {
    const long long tid = R.candidate_tile_ids[p];
    int a = -1;
    for (int q = 0; q < T.n; ++q)
      if (T.id[q] == tid) { a = q; break; }
    const long long slot = R.candidate_slots[p];
    if (a >= 0 && slot >= 0 && slot < m_tile) key = (unsigned)a * (unsigned)m_tile + (unsigned)slot;
  }
  keys[p] = key;
  vals[p] = (unsigned)p;
}

// Stable LSD radix sort of the (target key, pair index) list in ONE CTA (12,288 pairs at the reference budget): 8 bits per
// pass, keys and values ping-pong between two global buffers (L2-resident).  Every warp owns a contiguous run of items.
// Per pass: (1) per-warp digit counts in shared memory, (2) exclusive scan over (digit-major, warp-minor) -- the start of
// every (digit, warp) bucket, (3) each warp walks its run in order, 32 items at a time: an item's place is the bucket start
// plus the number of earlier items of its run with the same digit (running counter + `match_any` rank inside the step).
// Order inside a bucket = input order: stable, deterministic.  Replaces a library sort of five launches.
constexpr int kSortWarps = 32;
__global__ void __launch_bounds__(32 * kSortWarps) upd_sort_pairs_kernel(unsigned* __restrict__ k0, unsigned* __restrict__ v0,
                                                                         unsigned* __restrict__ k1, unsigned* __restrict__ v1, int n,
                                                                         int n_passes) {
  __shared__ int cnt[kSortWarps][256];     // counts, then running bucket positions
  __shared__ int dig_tot[256];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int per_warp = ((n + kSortWarps - 1) / kSortWarps + 31) & ~31;
  const int i0 = warp * per_warp, i1 = min(n, i0 + per_warp);
  unsigned *ks = k0, *vs = v0, *kd = k1, *vd = v1;
  for (int pass = 0; pass < n_passes; ++pass) {
    const int shift = 8 * pass;
    for (int e = tid; e < kSortWarps * 256; e += 32 * kSortWarps) (&cnt[0][0])[e] = 0;
    __syncthreads();
    for (int i = i0 + lane; i < i1; i += 32) atomicAdd(&cnt[warp][(ks[i] >> shift) & 255u], 1);
    __syncthreads();
    // exclusive scan: thread d < 256 walks the warps of digit d; then a scan over the 256 digit totals
    if (tid < 256) {
      int a = 0;
      for (int w = 0; w < kSortWarps; ++w) { const int c = cnt[w][tid]; cnt[w][tid] = a; a += c; }
      dig_tot[tid] = a;
    }
    __syncthreads();
    if (warp == 0) {   // 256 totals: 8 per lane, shuffle scan across lanes
      int loc[8], s_ = 0;
#pragma unroll
      for (int j = 0; j < 8; ++j) { loc[j] = s_; s_ += dig_tot[8 * lane + j]; }
      int inc = s_;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
      const int base = inc - s_;
#pragma unroll
      for (int j = 0; j < 8; ++j) dig_tot[8 * lane + j] = base + loc[j];
    }
    __syncthreads();
    for (int e = tid; e < kSortWarps * 256; e += 32 * kSortWarps) (&cnt[0][0])[e] += dig_tot[e & 255];
    __syncthreads();
    // stable scatter: the warp's run in order
    for (int base = i0; base < i1; base += 32) {
      const int i = base + lane;
      const bool on = i < i1;
      const unsigned key = on ? ks[i] : 0u, val = on ? vs[i] : 0u;
      const int d = on ? (int)((key >> shift) & 255u) : 256 + lane;     // inactive lanes: unique pseudo-digits
      const unsigned peers = __match_any_sync(0xffffffffu, d);
      const int before = __popc(peers & ((1u << lane) - 1u));
      if (on) {
        const int pos = cnt[warp][d] + before;
        kd[pos] = key; vd[pos] = val;
      }
      __syncwarp();
      if (on && lane == 31 - __clz(peers)) cnt[warp][d] += __popc(peers);
      __syncwarp();
    }
    __syncthreads();
    unsigned* t = ks; ks = kd; kd = t;
    t = vs; vs = vd; vd = t;
  }
}

// One WARP per sorted position; the warp of a segment head folds its segment into the tile slot
// (primitive_map.py:1037-1123): lanes stride over the segment's pairs (ascending pair index), then a fixed shuffle tree
// combines the 32 lane sums -- deterministic, and popular slots (hundreds of pairs) cost L/32 iterations instead of L.
constexpr int kFuseVals = 29;   // dLambda 9, deta 9, dtheta 3, dw, dr, dcam, dlid, dacc 3, dden
constexpr int kFuseU = 4;       // sub-steps of 32 sorted positions whose loads are in flight together
__global__ void __launch_bounds__(256) upd_fuse_kernel(gcs_atlas A, TileList T, gcs_meas_batch B, int K, gcs_assoc_result R,
                                                       UpdWs W, int n_pairs, unsigned none, gcs_map_update_cfg cfg,
                                                       double* __restrict__ part) {
  __shared__ double sred[8];
  const int lane = threadIdx.x & 31;
  const int q = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  double fused_mass = 0.0;
  if (q < n_pairs) {
    const unsigned key = W.pkeys[q];
    const bool head = key != none && (q == 0 || W.pkeys[q - 1] != key);   // warp-uniform
    if (head) {
      double v[kFuseVals];
#pragma unroll
      for (int k = 0; k < kFuseVals; ++k) v[k] = 0.0;
      // Rounds of kFuseU x 32 sorted positions.  The loads of a round are issued for all kFuseU sub-steps before any of
      // them is accumulated (clamped indices, responsibility 0 outside the segment: no branch between the loads), so a
      // popular slot with a thousand pairs pays ~L / 128 load latencies instead of L / 32.  Every lane still adds its
      // positions in ascending order.
      for (int base = q;; base += 32 * kFuseU) {
        int pi[kFuseU];
        double rr[kFuseU];
        bool ok[kFuseU];
#pragma unroll
        for (int u = 0; u < kFuseU; ++u) {
          const int j = base + 32 * u + lane;
          const int jc = j < n_pairs ? j : n_pairs - 1;
          ok[u] = j < n_pairs && W.pkeys[jc] == key;
          pi[u] = (int)W.pvals[jc];
        }
#pragma unroll
        for (int u = 0; u < kFuseU; ++u) rr[u] = ok[u] ? R.responsibilities[pi[u]] : 0.0;
#pragma unroll
        for (int u = 0; u < kFuseU; ++u) {
          const int i = pi[u] / K;
          const double r = rr[u];
          const double wm = B.weights[i];
#pragma unroll
          for (int k = 0; k < 9; ++k) { v[k] += r * W.Lw[9 * i + k]; v[9 + k] += r * W.etw[9 * i + k]; }
#pragma unroll
          for (int k = 0; k < 3; ++k) v[18 + k] += r * W.thw[3 * i + k];
          v[21] += r * wm; v[22] += r;
          const double wc = r * wm * (B.sources[i] == 0 ? 1.0 : 0.0), wl = r * wm * (B.sources[i] == 1 ? 1.0 : 0.0);
          v[23] += wc; v[24] += wl; v[28] += wc;
#pragma unroll
          for (int k = 0; k < 3; ++k) v[25 + k] += fmin(fmax(B.colors[3 * i + k], 0.0), 1.0) * wc;
          fused_mass += wm * r;
        }
        if (!__all_sync(0xffffffffu, ok[kFuseU - 1])) break;   // keys are sorted: the segment ended inside this round
      }
#pragma unroll
      for (int k = 0; k < kFuseVals; ++k) v[k] = warp_sum(v[k]);
      if (lane == 0) {
        const int a = (int)(key / (unsigned)A.m_tile), slot = (int)(key % (unsigned)A.m_tile);
        const int64_t o = (int64_t)T.index[a] * A.m_tile + slot;
        for (int k = 0; k < 9; ++k) { A.Lambdas[9 * o + k] += v[k]; A.etas[9 * o + k] += v[9 + k]; }
        for (int k = 0; k < 3; ++k) { A.thetas[3 * o + k] += v[18 + k]; A.rgb_cam_accum[3 * o + k] += v[25 + k]; }
        A.weights[o] += v[21]; A.cam_mass[o] += v[23]; A.lidar_mass[o] += v[24]; A.rgb_cam_denom[o] += v[28];
        if (v[22] > 0.0) { A.last_supported_scan_seq[o] = cfg.scan_seq; A.last_update_scan_seq[o] = cfg.scan_seq; }
        if (!cfg.strict_tile_state) A.timestamps[o] = cfg.timestamp;
      }
    }
  }
  const double s = warp_sum(fused_mass);
  if (lane == 0) sred[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int g = 0; g < 8; ++g) t += sred[g];
    part[blockIdx.x] = t;
  }
}

// quirk Q7: every fuse call stamps `timestamps` at ALL slot numbers of its block in the tile it was called for
__global__ void upd_stamp_strict_kernel(gcs_atlas A, TileList T, gcs_assoc_result R, int n_pairs, double ts) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n_pairs) return;
  const long long slot = R.candidate_slots[p];
  if (slot < 0 || slot >= A.m_tile) return;
  for (int a = 0; a < T.n; ++a) A.timestamps[(int64_t)T.index[a] * A.m_tile + slot] = ts;
}
// n_fused of one fuse call = number of distinct slot numbers in its block (:1155), summed over blocks x tiles
__global__ void __launch_bounds__(kBig) upd_block_unique_kernel(gcs_assoc_result R, int N, int K, int block_rows,
                                                                int* __restrict__ out_unique) {
  __shared__ unsigned long long s[4096];
  __shared__ int cnt;
  const int b = blockIdx.x;
  const int n = block_rows * K;
  int np = 1;
  while (np < n) np <<= 1;
  for (int e = threadIdx.x; e < np; e += kBig) {
    unsigned long long x = ~0ull;
    if (e < n) {
      int row = b * block_rows + e / K;
      if (row > N - 1) row = N - 1;  // rows past N are clipped to the last row (:578-579)
      x = (unsigned long long)R.candidate_slots[row * K + (e % K)];
    }
    s[e] = x;
  }
  if (threadIdx.x == 0) cnt = 0;
  cta_bitonic_sort_u64(s, np);
  int local = 0;
  for (int e = threadIdx.x; e < n; e += kBig)
    if (e == 0 || s[e] != s[e - 1]) ++local;
  if (local) atomicAdd(&cnt, local);
  __syncthreads();
  if (threadIdx.x == 0) out_unique[b] = cnt;
}

// rgb of every slot of the active tiles after fusion (:1100-1107)
__global__ void __launch_bounds__(256) upd_rgb_sweep_kernel(gcs_atlas A, TileList T, double eps_mass) {
  const int a = blockIdx.y;
  const int64_t base = (int64_t)T.index[a] * A.m_tile;
  for (int s = blockIdx.x * 256 + threadIdx.x; s < A.m_tile; s += gridDim.x * 256) {
    const int64_t o = base + s;
    const bool has = A.cam_mass[o] > 0.0;
    const double den = fmax(A.rgb_cam_denom[o], eps_mass);
    for (int k = 0; k < 3; ++k) {
      const double est = fmin(fmax(A.rgb_cam_accum[3 * o + k] / den, 0.0), 1.0);
      const double v = has ? est : 0.5;
      A.rgb[3 * o + k] = v;
      A.colors[3 * o + k] = v;
    }
  }
}

// per active tile: which measurements to insert and which slots to evict (pipeline.py:1348-1366, primitive_map.py:837-855)
__global__ void __launch_bounds__(kBig) upd_insert_select_kernel(gcs_atlas A, TileList T, gcs_meas_batch B, int N, UpdWs W,
                                                                 gcs_map_update_cfg cfg, int use_cache) {
  __shared__ SelectSmem sm;
  extern __shared__ uint32_t key_cache[];
  __shared__ int s_any;
  const int a = blockIdx.x, tid = threadIdx.x, k = cfg.k_insert_tile;
  const long long my_id = T.id[a];
  auto score_t = [&](int i) -> double { return (W.mtile[i] == my_id) ? W.score[i] : -1e30; };
  auto key1 = [&](int i) -> unsigned long long { return f64_orderable(-score_t(i)); };
  cta_select_k(N, k, key1, sm.out, sm.hist, sm.scan);
  if (tid == 0) s_any = 0;
  __syncthreads();
  int idx = 0;
  bool in_t = false, vnew = false;
  double w_ins = 0.0;
  if (tid < k) {
    idx = sm.out[tid].idx;
    in_t = W.mtile[idx] == my_id;
    vnew = in_t && (score_t(idx) > -1e20);
    w_ins = in_t ? W.novelty[idx] * B.weights[idx] : 0.0;
    if (vnew) atomicOr(&s_any, 1);
  }
  __syncthreads();
  if (tid < k) {
    if (!s_any) vnew = true;  // zero-mass placeholders keep the insert count fixed (:1355, quirk Q6)
    W.ins_idx[a * k + tid] = idx;
    W.ins_new[a * k + tid] = vnew ? 1 : 0;
    W.ins_w[a * k + tid] = w_ins;
  }
  __syncthreads();
  // eviction targets: k lowest retention, empty slots first
  const int64_t base = (int64_t)T.index[a] * A.m_tile;
  auto key2 = [&](int s) -> unsigned long long {
    const int64_t o = base + s;
    const uint8_t v = A.valid[o];                         // three independent loads (see map_view_kernel)
    const long long lss = A.last_supported_scan_seq[o];
    const double w = A.weights[o];
    double keyv = -INFINITY;
    if (v) {
      long long dt = cfg.scan_seq - lss;
      if (dt < 0) dt = 0;
      keyv = w * exp(-cfg.recency_decay_lambda * (double)dt);
    }
    return f64_orderable(keyv);
  };
  // a tile with at least k empty slots evicts nothing: the first k empty slots in index order
  auto empty = [&](int s) -> bool { return A.valid[base + s] == 0; };
  if (!cta_select_min_sentinel(A.m_tile, k, empty, f64_orderable(-INFINITY), sm.out, sm.scan))
    cta_select_k(A.m_tile, k, key2, sm.out, sm.hist, sm.scan, use_cache ? key_cache : nullptr);
  if (tid < k) W.ins_slot[a * k + tid] = sm.out[tid].idx;
  if (tid == 0) {
    int n = 0;
    double mass = 0.0;
    for (int j = 0; j < k; ++j) { n += W.ins_new[a * k + j]; mass += W.ins_w[a * k + j]; }
    W.n_ins[a] = n;
    // p95 of the insert masses of this tile: ascending rank int(0.95 k)
    int r95 = (int)(0.95 * (double)k);
    if (r95 > k - 1) r95 = k - 1;
    double p95 = 0.0;
    for (int j = 0; j < k; ++j) {
      const double x = W.ins_w[a * k + j];
      int less = 0, eq = 0;
      for (int m = 0; m < k; ++m) { const double y = W.ins_w[a * k + m]; less += (y < x); eq += (y == x); }
      if (less <= r95 && r95 < less + eq) { p95 = x; break; }
    }
    W.part[64 + 2 * a] = mass;
    W.part[64 + 2 * a + 1] = p95;
  }
}

__global__ void upd_insert_apply_kernel(gcs_atlas A, TileList T, gcs_meas_batch B, UpdWs W, gcs_map_update_cfg cfg,
                                        long long* __restrict__ out_ids, int* __restrict__ out_slots) {
  const int a = blockIdx.x, j = threadIdx.x, k = cfg.k_insert_tile;
  if (j >= k) return;
  long long base_id = cfg.next_global_id_dev ? (long long)*cfg.next_global_id_dev : (long long)cfg.next_global_id;
  for (int q = 0; q < a; ++q) base_id += W.n_ins[q];
  int prefix = 0;
  for (int m = 0; m <= j; ++m) prefix += W.ins_new[a * k + m];
  const bool doit = W.ins_new[a * k + j] != 0;
  const int slot = W.ins_slot[a * k + j];
  const long long nid = doit ? base_id + (prefix - 1) : -1;
  out_ids[a * k + j] = nid;
  out_slots[a * k + j] = slot;
  if (!doit) return;
  const int i = W.ins_idx[a * k + j];
  const int64_t o = (int64_t)T.index[a] * A.m_tile + slot;
  const double w = W.ins_w[a * k + j];
  for (int c = 0; c < 9; ++c) { A.Lambdas[9 * o + c] = W.Lw[9 * i + c]; A.etas[9 * o + c] = W.etw[9 * i + c]; }
  const bool is_cam = B.sources[i] == 0, is_lid = B.sources[i] == 1;
  const double cam = is_cam ? w : 0.0;
  for (int c = 0; c < 3; ++c) {
    A.thetas[3 * o + c] = W.thw[3 * i + c];
    const double col = B.colors[3 * i + c];
    const double rgbn = (cam > 0.0) ? fmin(fmax(col, 0.0), 1.0) : 0.5;
    A.colors[3 * o + c] = rgbn; A.rgb[3 * o + c] = rgbn;
    A.rgb_cam_accum[3 * o + c] = col * cam;
  }
  A.weights[o] = w; A.timestamps[o] = cfg.timestamp; A.created_timestamps[o] = cfg.timestamp;
  A.last_supported_scan_seq[o] = cfg.scan_seq; A.last_update_scan_seq[o] = cfg.scan_seq;
  A.primitive_ids[o] = nid; A.valid[o] = 1;
  A.cam_mass[o] = cam; A.lidar_mass[o] = is_lid ? w : 0.0; A.rgb_cam_denom[o] = cam;
}

// cull (w < threshold) then forget (w *= gamma) over the active tiles  (primitive_map.py:1175-1384)
__global__ void __launch_bounds__(256) upd_maintain_kernel(gcs_atlas A, TileList T, double thr, double gamma,
                                                           double* __restrict__ part) {
  __shared__ double sred[3][8];
  const int a = blockIdx.y;
  const int64_t base = (int64_t)T.index[a] * A.m_tile;
  double n_cull = 0, m_cull = 0, n_valid = 0;
  for (int s = blockIdx.x * 256 + threadIdx.x; s < A.m_tile; s += gridDim.x * 256) {
    const int64_t o = base + s;
    const double w = A.weights[o];
    if (A.valid[o]) {
      if (w < thr) { A.valid[o] = 0; n_cull += 1.0; m_cull += w; }
      else n_valid += 1.0;
    }
    A.weights[o] = gamma * w;
  }
  double v3[3] = {n_cull, m_cull, n_valid};
  for (int k = 0; k < 3; ++k) {
    double r = warp_sum(v3[k]);
    if ((threadIdx.x & 31) == 0) sred[k][threadIdx.x >> 5] = r;
  }
  __syncthreads();
  if (threadIdx.x < 3) {
    double s = 0.0;
    for (int g = 0; g < 8; ++g) s += sred[threadIdx.x][g];
    part[((int64_t)a * gridDim.x + blockIdx.x) * 3 + threadIdx.x] = s;
  }
}

// fixed-order block sum of a strided array by 256 threads (thread t adds elements t, t+256, ..., then a shuffle /
// shared-memory tree): deterministic, and ~n/256 dependent loads instead of n
__device__ __forceinline__ double sum256(const double* __restrict__ p, int n, int stride, double* sbuf) {
  double a = 0.0;
  for (int i = threadIdx.x; i < n; i += 256) a += p[(int64_t)i * stride];
  a = warp_sum(a);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) sbuf[threadIdx.x >> 5] = a;
  __syncthreads();
  double r = 0.0;
  for (int w = 0; w < 8; ++w) r += sbuf[w];
  __syncthreads();
  return r;
}

__global__ void __launch_bounds__(256) upd_stats_kernel(UpdWs W, const double* __restrict__ fuse_part, int n_fuse_parts,
                                                        const int* __restrict__ blk_unique, int n_blocks,
                                                        const double* __restrict__ maint_part, int maint_blocks, int n_tiles,
                                                        int k_ins, gcs_map_update_cfg cfg, double* __restrict__ stats) {
  __shared__ double sbuf[8];
  const double fm = sum256(fuse_part, n_fuse_parts, 1, sbuf);
  double nc = 0.0, mc = 0.0;
  for (int a = 0; a < n_tiles; ++a) {
    const double* mp = maint_part + (int64_t)a * maint_blocks * 3;
    nc += sum256(mp, maint_blocks, 3, sbuf);
    mc += sum256(mp + 1, maint_blocks, 3, sbuf);
    const double tv = sum256(mp + 2, maint_blocks, 3, sbuf);
    if (threadIdx.x == 0 && a < 16) stats[GCS_MU_TILE_COUNT0 + a] = tv;
  }
  if (threadIdx.x != 0) return;
  long long uq = 0;
  for (int b = 0; b < n_blocks; ++b) uq += blk_unique[b];
  double im = 0.0, p95 = 0.0;
  long long ni = 0;
  for (int a = 0; a < n_tiles; ++a) {
    im += W.part[64 + 2 * a];
    p95 = fmax(p95, W.part[64 + 2 * a + 1]);
    ni += W.n_ins[a];
  }
  stats[GCS_MU_FUSED_COUNT] = (double)(uq * n_tiles);
  stats[GCS_MU_FUSED_MASS] = fm;
  stats[GCS_MU_INSERT_COUNT] = (double)ni;
  stats[GCS_MU_INSERT_MASS] = im;
  stats[GCS_MU_INSERT_MASS_P95] = p95;
  stats[GCS_MU_EVICTED_COUNT] = nc;
  stats[GCS_MU_EVICTED_MASS] = mc;
  // the id counter: host scalar of this call, or the device-resident one (read above by upd_insert_apply_kernel, advanced here,
  // behind it in stream order, by the single thread that reaches this point)
  const long long id0 = cfg.next_global_id_dev ? (long long)*cfg.next_global_id_dev : (long long)cfg.next_global_id;
  stats[GCS_MU_NEXT_GLOBAL_ID] = (double)(id0 + ni);
  if (cfg.next_global_id_dev) *cfg.next_global_id_dev = (int64_t)(id0 + ni);
  (void)k_ins;
}

}  // namespace gcs

using namespace gcs;

static int64_t cdivm(int64_t a, int64_t b) { return (a + b - 1) / b; }

static int check_atlas(gcs_ctx* ctx, const gcs_atlas* a, const char* who) {
  GCS_REQUIRE(ctx, a && a->Lambdas && a->thetas && a->etas && a->weights && a->timestamps && a->created_timestamps &&
                       a->last_supported_scan_seq && a->last_update_scan_seq && a->primitive_ids && a->valid && a->colors &&
                       a->cam_mass && a->lidar_mass && a->rgb_cam_accum && a->rgb_cam_denom && a->rgb,
              "%s: atlas pointer is NULL", who);
  GCS_REQUIRE(ctx, a->m_tile >= 1 && a->n_tiles_cap >= 1, "%s: bad atlas shape", who);
  return GCS_OK;
}
static int make_tile_list(gcs_ctx* ctx, const gcs_atlas* a, const int32_t* idx, const int64_t* ids, int n, bool allow_missing,
                          TileList* T, const char* who) {
  GCS_REQUIRE(ctx, idx && n >= 1 && n <= 16, "%s: n_tiles=%d not in [1,16]", who, n);
  T->n = n;
  for (int i = 0; i < 16; ++i) { T->index[i] = -1; T->id[i] = 0; }
  for (int i = 0; i < n; ++i) {
    GCS_REQUIRE(ctx, idx[i] < a->n_tiles_cap, "%s: tile index %d out of range", who, idx[i]);
    GCS_REQUIRE(ctx, allow_missing || idx[i] >= 0, "%s: tile %d missing from the pool", who, i);
    T->index[i] = idx[i];
    if (ids) T->id[i] = ids[i];
  }
  return GCS_OK;
}
static int check_view(gcs_ctx* ctx, const gcs_map_view* v, const char* who) {
  GCS_REQUIRE(ctx, v && v->candidate_tile_ids && v->candidate_slots && v->valid && v->positions && v->covariances &&
                       v->directions && v->kappas && v->weights && v->primitive_ids && v->last_supported_scan_seq && v->etas &&
                       v->colors, "%s: view pointer is NULL", who);
  return GCS_OK;
}
static int check_assoc(gcs_ctx* ctx, const gcs_assoc_result* r, const char* who) {
  GCS_REQUIRE(ctx, r && r->responsibilities && r->candidate_pool_indices && r->candidate_tile_ids && r->candidate_slots &&
                       r->row_masses && r->cost_matrix, "%s: association pointer is NULL", who);
  return GCS_OK;
}
static int check_mbatch(gcs_ctx* ctx, const gcs_meas_batch* b, const char* who) {
  GCS_REQUIRE(ctx, b && b->Lambdas && b->thetas && b->etas && b->weights && b->sources && b->source_indices && b->valid &&
                       b->timestamps && b->colors, "%s: measurement batch pointer is NULL", who);
  return GCS_OK;
}

extern "C" {

int gcs_map_recency_inflate(gcs_ctx* ctx, void* stream, const gcs_atlas* atlas, const int32_t* tile_index, int32_t n_tiles,
                            int64_t scan_seq, double lam, double min_scale, double* stats) {
  if (!ctx) return GCS_EINVAL;
  GCS_CHECK_CUDA(ctx, cudaSetDevice(ctx->device));
  int rc = check_atlas(ctx, atlas, "map_recency_inflate");
  if (rc) return rc;
  TileList T;
  rc = make_tile_list(ctx, atlas, tile_index, nullptr, n_tiles, true, &T, "map_recency_inflate");
  if (rc) return rc;
  GCS_REQUIRE(ctx, stats != nullptr, "map_recency_inflate: stats is NULL");
  const int blocks = (int)(cdivm(atlas->m_tile, 256) < 64 ? cdivm(atlas->m_tile, 256) : 64);
  cudaStream_t st;
  char* wsb;
  rc = gcs_maint_stream(ctx, (cudaStream_t)stream, (uint64_t)n_tiles * blocks * 3 * 8, &st, &wsb);
  if (rc) return rc;
  double* part = (double*)wsb;
  gcs_timing_begin(ctx, st, GCS_TIME_INFLATE);
  recency_inflate_kernel<<<dim3(blocks, n_tiles), 256, 0, st>>>(*atlas, T, scan_seq, lam, min_scale, part, 1);
  gcs_timing_end(ctx, st, GCS_TIME_INFLATE);
  GCS_LAUNCH_CHECK(ctx);
  sum_parts_kernel<<<1, 32, 0, st>>>(part, n_tiles * blocks, 3, stats, 4);
  GCS_LAUNCH_CHECK(ctx);
  return GCS_OK;
}

static int map_view_launch(gcs_ctx* ctx, cudaStream_t st, const gcs_atlas* atlas, const int32_t* tile_index,
                           const int64_t* tile_ids, int32_t n_tiles, int32_t m_tile_view, double eps_lift, double eps_mass,
                           const gcs_map_view* view, int32_t* out_n_valid, InflateArg I, double* inflate_stats, const char* who,
                           double* part_ws = nullptr /* scratch for the inflation partials; default: the context workspace */) {
  int rc = check_atlas(ctx, atlas, who);
  if (rc) return rc;
  GCS_REQUIRE(ctx, m_tile_view > 0, "%s: m_tile_view must be > 0, got %d", who, m_tile_view);
  GCS_REQUIRE(ctx, m_tile_view <= 1024 && m_tile_view <= atlas->m_tile, "%s: m_tile_view=%d exceeds 1024 or m_tile", who, m_tile_view);
  rc = check_view(ctx, view, who);
  if (rc) return rc;
  GCS_REQUIRE(ctx, tile_ids && out_n_valid, "%s: NULL pointer", who);
  TileList T;
  rc = make_tile_list(ctx, atlas, tile_index, tile_ids, n_tiles, true, &T, who);
  if (rc) return rc;
  if (inflate_stats) {   // statistics of the inflation the view applies functionally; the map itself is not modified
    const int blocks = (int)(cdivm(atlas->m_tile, 256) < 64 ? cdivm(atlas->m_tile, 256) : 64);
    double* part = part_ws;
    if (!part) {
      rc = gcs_ws_reserve(ctx, (uint64_t)n_tiles * blocks * 3 * 8);
      if (rc) return rc;
      part = (double*)ctx->ws;
    }
    recency_inflate_kernel<<<dim3(blocks, n_tiles), 256, 0, st>>>(*atlas, T, I.scan_seq, I.lam, I.min_scale, part, 0);
    GCS_LAUNCH_CHECK(ctx);
    sum_parts_kernel<<<1, 32, 0, st>>>(part, n_tiles * blocks, 3, inflate_stats, 4);
    GCS_LAUNCH_CHECK(ctx);
  }
  GCS_CHECK_CUDA(ctx, cudaMemsetAsync(out_n_valid, 0, sizeof(int32_t), st));
  gcs_timing_begin(ctx, st, GCS_TIME_MAP_VIEW);
  // no key cache here: the view's key is two cached loads, and the 200 KB carve-out it would take from L1 costs more
  // than the re-evaluations (measured 208 us with, 158 us without; the eviction select of the map update gains 35 %)
  map_view_kernel<<<n_tiles, kBig, 0, st>>>(*atlas, T, m_tile_view, *view, 0);
  gcs_timing_end(ctx, st, GCS_TIME_MAP_VIEW);
  GCS_LAUNCH_CHECK(ctx);
  map_view_gather_kernel<<<dim3((m_tile_view + 127) / 128, n_tiles), 128, 0, st>>>(*atlas, T, m_tile_view, eps_lift, eps_mass, *view,
                                                                                  out_n_valid, I);
  GCS_LAUNCH_CHECK(ctx);
  return GCS_OK;
}

int gcs_extract_atlas_map_view(gcs_ctx* ctx, void* stream, const gcs_atlas* atlas, const int32_t* tile_index,
                               const int64_t* tile_ids, int32_t n_tiles, int32_t m_tile_view, double eps_lift,
                               double eps_mass, const gcs_map_view* view, int32_t* out_n_valid) {
  if (!ctx) return GCS_EINVAL;
  GCS_CHECK_CUDA(ctx, cudaSetDevice(ctx->device));
  InflateArg I = {0, 0, 0.0, 1.0};
  return map_view_launch(ctx, (cudaStream_t)stream, atlas, tile_index, tile_ids, n_tiles, m_tile_view, eps_lift, eps_mass, view,
                         out_n_valid, I, nullptr, "extract_atlas_map_view");
}

int gcs_extract_atlas_map_view_inflated(gcs_ctx* ctx, void* stream, const gcs_atlas* atlas, const int32_t* tile_index,
                                        const int64_t* tile_ids, int32_t n_tiles, int32_t m_tile_view, double eps_lift,
                                        double eps_mass, int64_t scan_seq, double recency_decay_lambda, double min_scale,
                                        const gcs_map_view* view, int32_t* out_n_valid, double* out_inflate_stats) {
  if (!ctx) return GCS_EINVAL;
  GCS_CHECK_CUDA(ctx, cudaSetDevice(ctx->device));
  InflateArg I = {1, (long long)scan_seq, recency_decay_lambda, min_scale};
  return map_view_launch(ctx, (cudaStream_t)stream, atlas, tile_index, tile_ids, n_tiles, m_tile_view, eps_lift, eps_mass, view,
                         out_n_valid, I, out_inflate_stats, "extract_atlas_map_view_inflated");
}

// The per-view part of the association (A_vmf of the entries, the grouped view): buffers + launch.  The fused entry runs
// it on the side stream, in the side workspace, while the scan's surfels are extracted.
struct ViewPrep {
  double* vAk; double* gpos; uint16_t* goff; uint8_t* gval; double* gbox;
};
static size_t view_prep_bytes(int n_tiles, int m_view) {
  const size_t P = (size_t)n_tiles * m_view;
  auto up = [](size_t b) { return (b + 255) & ~(size_t)255; };
  return up(P * 8) + up(P * 24) + up(P * 2) + up(P) + up(topk_box_bytes(n_tiles, m_view));
}
static ViewPrep view_prep_carve(char* base, int n_tiles, int m_view) {
  const size_t P = (size_t)n_tiles * m_view;
  auto up = [](size_t b) { return (b + 255) & ~(size_t)255; };
  ViewPrep p;
  p.vAk = (double*)base; base += up(P * 8);
  p.gpos = (double*)base; base += up(P * 24);
  p.goff = (uint16_t*)base; base += up(P * 2);
  p.gval = (uint8_t*)base; base += up(P);
  p.gbox = (double*)base;
  return p;
}
static int view_prep_launch(gcs_ctx* ctx, cudaStream_t st, const gcs_map_view* view, int n_tiles, int m_view, const ViewPrep& p) {
  AssocWs W;
  memset(&W, 0, sizeof(W));
  W.vAk = p.vAk; W.gpos = p.gpos; W.goff = p.goff; W.gval = p.gval; W.gbox = p.gbox;
  assoc_view_groups_kernel<<<n_tiles, 1024, 0, st>>>(*view, m_view, W);
  GCS_LAUNCH_CHECK(ctx);
  return GCS_OK;
}

static int assoc_launch(gcs_ctx* ctx, cudaStream_t st, const gcs_meas_batch* batch, int n_units, const gcs_map_view* view,
                        const int64_t* view_tile_ids, int32_t n_tiles, int32_t m_tile_view, const gcs_assoc_cfg* cfg,
                        const gcs_assoc_result* out, double* cert, const char* who, const ViewPrep* pre = nullptr) {
  int rc = check_mbatch(ctx, batch, who);
  if (rc) return rc;
  rc = check_view(ctx, view, who);
  if (rc) return rc;
  rc = check_assoc(ctx, out, who);
  if (rc) return rc;
  GCS_REQUIRE(ctx, cfg && cert && view_tile_ids, "%s: NULL pointer", who);
  GCS_REQUIRE(ctx, GCS_K_ASSOC_OK(cfg->k_assoc), "%s: k_assoc=%d (this build instantiates K_ASSOC = 4, 8, 16)", who, cfg->k_assoc);
  GCS_REQUIRE(ctx, cfg->a_policy == 0 || cfg->a_policy == 1, "%s: a_policy=%d (0 UNIFORM, 1 WEIGHT_PROPORTIONAL)", who, cfg->a_policy);
  GCS_REQUIRE(ctx, n_tiles >= 1 && n_tiles <= 16 && m_tile_view >= cfg->k_assoc, "%s: bad view shape", who);
  const int N = batch->n_feat + batch->n_surfel;
  GCS_REQUIRE(ctx, N >= 1 && N <= 2048, "%s: N_total=%d exceeds the single-cluster Sinkhorn budget 2048", who, N);
  GCS_REQUIRE(ctx, cfg->r_stencil_xy >= 0 && cfg->r_stencil_xy <= 2 && cfg->r_stencil_z >= 0 && cfg->r_stencil_z <= 1,
              "%s: stencil radius out of range", who);
  // stencil offsets in the reference's order: z slab outer, sorted axial disk inner (tiling.py:171-186)
  StencilOffsets SO;
  memset(&SO, 0, sizeof(SO));
  int n_st = 0;
  const int rr = cfg->r_stencil_xy;
  for (int z = -cfg->r_stencil_z; z <= cfg->r_stencil_z; ++z)
    for (int q = -rr; q <= rr; ++q) {
      const int r_min = (-rr > -q - rr) ? -rr : -q - rr, r_max = (rr < -q + rr) ? rr : -q + rr;
      for (int r = r_min; r <= r_max; ++r) { SO.dq[n_st] = (int8_t)q; SO.dr[n_st] = (int8_t)r; SO.dz[n_st] = (int8_t)z; ++n_st; }
    }
  TileList T;
  T.n = n_tiles;
  for (int i = 0; i < 16; ++i) { T.index[i] = i; T.id[i] = i < n_tiles ? view_tile_ids[i] : 0; }
  const size_t H = (size_t)n_units;
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t o = off; off += (bytes + 255) & ~(size_t)255; return o; };
  const size_t o_pos = take(H * N * 3 * 8), o_dir = take(H * N * 3 * 8), o_kap = take(H * N * 8),
               o_st = take(H * N * n_st), o_brow = take(H * N * (size_t)cfg->k_assoc * 8),
               o_ctr = take(256), o_aws = take(H * N * 8), o_prep = take(pre ? 0 : view_prep_bytes(n_tiles, m_tile_view));
  rc = gcs_ws_reserve(ctx, off);
  if (rc) return rc;
  char* ws = (char*)ctx->ws;
  AssocWs W;
  W.mpos = (double*)(ws + o_pos); W.mdir = (double*)(ws + o_dir); W.mkap = (double*)(ws + o_kap); W.stencil = (int8_t*)(ws + o_st);
  GCS_REQUIRE(ctx, m_tile_view <= 65535, "%s: m_tile_view=%d exceeds the 16-bit offsets of the grouped view", who, m_tile_view);
  const ViewPrep vp = pre ? *pre : view_prep_carve(ws + o_prep, n_tiles, m_tile_view);
  W.vAk = vp.vAk; W.gpos = vp.gpos; W.goff = vp.goff; W.gval = vp.gval; W.gbox = vp.gbox;
  const int n_pool = n_tiles * m_tile_view, row_blocks = (N + 127) / 128;
  const unsigned Hu = (unsigned)n_units;
  assoc_prepare_kernel<<<dim3(row_blocks, Hu), 128, 0, st>>>(*batch, N, T, *cfg, n_st, SO, W, *view, n_pool, row_blocks);
  GCS_LAUNCH_CHECK(ctx);
  // the view in group order + A_vmf of its entries (once per view: shared by every unit and every row), unless the
  // caller has prepared them already
  if (!pre) {
    rc = view_prep_launch(ctx, st, view, n_tiles, m_tile_view, vp);
    if (rc) return rc;
  }
  // top-K: persistent CTAs pulling rows from a counter; view tiles staged once per CTA by TMA when the box fits
  CUtensorMap tmap;
  memset(&tmap, 0, sizeof(tmap));
  bool staged = (m_tile_view % 256 == 0) && topk_smem_bytes(n_tiles, m_tile_view, true) <= 220 * 1024 &&
                tma::encode_2d_f64(&tmap, W.gpos, 256, (uint64_t)n_pool * 3 / 256, 256, (uint32_t)(3 * m_tile_view / 256));
  if (getenv("GCS_TOPK_NO_TMA")) staged = false;
  const size_t topk_smem = topk_smem_bytes(n_tiles, m_tile_view, staged);
  if (topk_smem > 40 * 1024)
    GCS_K_ASSOC_DISPATCH(cfg->k_assoc, GCS_CHECK_CUDA(ctx, gcs_smem_attr_once((const void*)assoc_topk_kernel<KK>, (int)topk_smem)));
  int* row_counter = (int*)(ws + o_ctr);
  GCS_CHECK_CUDA(ctx, cudaMemsetAsync(row_counter, 0, sizeof(int), st));
  const long long total_rows = (long long)n_units * N;
  int topk_ctas = (int)((total_rows + kTopkWarps - 1) / kTopkWarps);
  if (topk_ctas > ctx->sm_count) topk_ctas = ctx->sm_count;
  gcs_timing_begin(ctx, st, GCS_TIME_TOPK);
  GCS_K_ASSOC_DISPATCH(cfg->k_assoc, (assoc_topk_kernel<KK><<<topk_ctas, 32 * kTopkWarps, topk_smem, st>>>(
                                         *batch, N, n_units, *view, m_tile_view, n_st, n_tiles, W, *cfg, *out, row_counter, tmap, staged ? 1 : 0)));
  gcs_timing_end(ctx, st, GCS_TIME_TOPK);
  GCS_LAUNCH_CHECK(ctx);
  // cluster size per hypothesis: eight SMs for one, fewer when the batch fills the device anyway
  double* brow = (double*)(ws + o_brow);
  double* a_ws = (double*)(ws + o_aws);
  // the iterations are a latency chain (log / exp / barrier), so as many CTAs as stay resident (three per SM) overlap
  const int sk_rpt = (n_units * 8 <= 3 * ctx->sm_count) ? 1 : (n_units * 4 <= 3 * ctx->sm_count ? 2 : 4);
  gcs_timing_begin(ctx, st, GCS_TIME_SINKHORN);
  GCS_K_ASSOC_DISPATCH(cfg->k_assoc,
    if (sk_rpt == 1) GCS_CHECK_CUDA(ctx, (sinkhorn_launch<KK, 1>(st, Hu, *batch, N, *view, W, *cfg, *out, cert, brow, a_ws)));
    else if (sk_rpt == 2) GCS_CHECK_CUDA(ctx, (sinkhorn_launch<KK, 2>(st, Hu, *batch, N, *view, W, *cfg, *out, cert, brow, a_ws)));
    else GCS_CHECK_CUDA(ctx, (sinkhorn_launch<KK, 4>(st, Hu, *batch, N, *view, W, *cfg, *out, cert, brow, a_ws))));
  gcs_timing_end(ctx, st, GCS_TIME_SINKHORN);
  GCS_LAUNCH_CHECK(ctx);
  return GCS_OK;
}

int gcs_associate_primitives_ot(gcs_ctx* ctx, void* stream, const gcs_meas_batch* batch, const gcs_map_view* view,
                                const int64_t* view_tile_ids, int32_t n_tiles, int32_t m_tile_view,
                                const gcs_assoc_cfg* cfg, const gcs_assoc_result* out, double* cert) {
  if (!ctx) return GCS_EINVAL;
  GCS_CHECK_CUDA(ctx, cudaSetDevice(ctx->device));
  return assoc_launch(ctx, (cudaStream_t)stream, batch, 1, view, view_tile_ids, n_tiles, m_tile_view, cfg, out, cert,
                      "associate_primitives_ot");
}

int gcs_associate_primitives_ot_batched(gcs_ctx* ctx, void* stream, const gcs_meas_batch* batch, int32_t n_units,
                                        const gcs_map_view* view, const int64_t* view_tile_ids, int32_t n_tiles,
                                        int32_t m_tile_view, const gcs_assoc_cfg* cfg, const gcs_assoc_result* out,
                                        double* cert) {
  if (!ctx) return GCS_EINVAL;
  GCS_CHECK_CUDA(ctx, cudaSetDevice(ctx->device));
  GCS_REQUIRE(ctx, n_units >= 1 && n_units <= 65535, "associate_primitives_ot_batched: n_units=%d", n_units);
  return assoc_launch(ctx, (cudaStream_t)stream, batch, n_units, view, view_tile_ids, n_tiles, m_tile_view, cfg, out, cert,
                      "associate_primitives_ot_batched");
}

}  // extern "C"

// per-unit partial sums + arrival counters of pose_evidence_kernel, at the start of the context workspace (whatever an
// earlier call of this stream left there is dead by the time this one runs)
static int pose_evidence_scratch(gcs_ctx* ctx, cudaStream_t st, int n_units, double** partial, int** arrived) {
  const size_t pbytes = (((size_t)n_units * kPeParts * 25 * 8) + 255) & ~(size_t)255;
  const int rc = gcs_ws_reserve(ctx, pbytes + (size_t)n_units * 4);
  if (rc) return rc;
  *partial = (double*)ctx->ws;
  *arrived = (int*)((char*)ctx->ws + pbytes);
  GCS_CHECK_CUDA(ctx, cudaMemsetAsync(*arrived, 0, (size_t)n_units * 4, st));
  return GCS_OK;
}

extern "C" {

int gcs_visual_pose_evidence(gcs_ctx* ctx, void* stream, const gcs_meas_batch* batch, const gcs_map_view* view,
                             const gcs_assoc_result* assoc, int32_t k_assoc, const double* pose6, double eps_lift,
                             double eps_mass, double* out_L22, double* out_h22, double* out_rec) {
  if (!ctx) return GCS_EINVAL;
  GCS_CHECK_CUDA(ctx, cudaSetDevice(ctx->device));
  int rc = check_mbatch(ctx, batch, "visual_pose_evidence");
  if (rc) return rc;
  rc = check_view(ctx, view, "visual_pose_evidence");
  if (rc) return rc;
  rc = check_assoc(ctx, assoc, "visual_pose_evidence");
  if (rc) return rc;
  GCS_REQUIRE(ctx, pose6 && out_L22 && out_h22 && out_rec, "visual_pose_evidence: NULL pointer");
  GCS_REQUIRE(ctx, GCS_K_ASSOC_OK(k_assoc), "visual_pose_evidence: k_assoc=%d (this build instantiates K_ASSOC = 4, 8, 16)", k_assoc);
  const int N = batch->n_feat + batch->n_surfel;
  double* partial;
  int* arrived;
  rc = pose_evidence_scratch(ctx, (cudaStream_t)stream, 1, &partial, &arrived);
  if (rc) return rc;
  GCS_K_ASSOC_DISPATCH(k_assoc, (pose_evidence_kernel<KK><<<dim3(kPeParts, 1), kPeThreads, 0, (cudaStream_t)stream>>>(
                                    *batch, N, *view, *assoc, pose6[0], pose6[1], pose6[2], pose6[3], pose6[4], pose6[5], eps_lift,
                                    eps_mass, out_L22, out_h22, out_rec, nullptr, nullptr, 0, nullptr, partial, arrived)));
  GCS_LAUNCH_CHECK(ctx);
  return GCS_OK;
}

int gcs_visual_pose_evidence_batched(gcs_ctx* ctx, void* stream, const gcs_meas_batch* batch, int32_t n_units,
                                     const gcs_map_view* view, const gcs_assoc_result* assoc, int32_t k_assoc,
                                     const double* poses_dev, double eps_lift, double eps_mass, double* out_L22,
                                     double* out_h22, double* out_rec, const int32_t* n_lidar_valid,
                                     int32_t n_camera_valid, const int32_t* view_n_valid) {
  if (!ctx) return GCS_EINVAL;
  GCS_CHECK_CUDA(ctx, cudaSetDevice(ctx->device));
  int rc = check_mbatch(ctx, batch, "visual_pose_evidence_batched");
  if (rc) return rc;
  rc = check_view(ctx, view, "visual_pose_evidence_batched");
  if (rc) return rc;
  rc = check_assoc(ctx, assoc, "visual_pose_evidence_batched");
  if (rc) return rc;
  GCS_REQUIRE(ctx, poses_dev && out_L22 && out_h22 && out_rec && n_units >= 1, "visual_pose_evidence_batched: bad args");
  GCS_REQUIRE(ctx, GCS_K_ASSOC_OK(k_assoc), "visual_pose_evidence_batched: k_assoc=%d (this build instantiates K_ASSOC = 4, 8, 16)", k_assoc);
  const int N = batch->n_feat + batch->n_surfel;
  GCS_REQUIRE(ctx, n_units <= 65535, "visual_pose_evidence_batched: n_units=%d", n_units);
  double* partial;
  int* arrived;
  rc = pose_evidence_scratch(ctx, (cudaStream_t)stream, n_units, &partial, &arrived);
  if (rc) return rc;
  GCS_K_ASSOC_DISPATCH(k_assoc, (pose_evidence_kernel<KK><<<dim3(kPeParts, (unsigned)n_units), kPeThreads, 0, (cudaStream_t)stream>>>(
                                    *batch, N, *view, *assoc, 0, 0, 0, 0, 0, 0, eps_lift, eps_mass, out_L22, out_h22, out_rec,
                                    poses_dev, n_lidar_valid, n_camera_valid, view_n_valid, partial, arrived)));
  GCS_LAUNCH_CHECK(ctx);
  return GCS_OK;
}

int gcs_lidar_evidence_primitives_batched(gcs_ctx* ctx, void* stream, const gcs_prim_batch_args* a) {
  if (!ctx) return GCS_EINVAL;
  GCS_CHECK_CUDA(ctx, cudaSetDevice(ctx->device));
  GCS_REQUIRE(ctx, a && a->atlas && a->n_units >= 1 && a->n >= 1 && a->n_tiles >= 1 && a->n_tiles <= 16,
              "lidar_evidence_primitives_batched: bad args");
  cudaStream_t st = (cudaStream_t)stream;
  // Fork: the map side of the path (inflation statistics, view selection + gather, grouped view) does not depend on the
  // scan; it runs on the context's side stream, with scratch of its own, while the scan is deskewed and its surfels are
  // extracted, and joins the caller's stream in front of the association.  The view kernels occupy one CTA per tile.
  const int infl_blocks = (int)(cdivm(a->atlas->m_tile, 256) < 64 ? cdivm(a->atlas->m_tile, 256) : 64);
  const size_t part_bytes = (((size_t)a->n_tiles * infl_blocks * 3 * 8) + 255) & ~(size_t)255;
  int rc = gcs_side_reserve(ctx, part_bytes + view_prep_bytes(a->n_tiles, a->m_tile_view));
  if (rc) return rc;
  cudaStream_t side = ctx->side_stream;
  GCS_CHECK_CUDA(ctx, cudaEventRecord(ctx->ev_fork, st));
  GCS_CHECK_CUDA(ctx, cudaStreamWaitEvent(side, ctx->ev_fork, 0));
  {
    InflateArg I = {a->inflate ? 1 : 0, (long long)a->assoc_cfg.scan_seq, a->inflate ? a->assoc_cfg.recency_decay_lambda : 0.0,
                    a->inflate ? a->recency_min_scale : 1.0};
    rc = map_view_launch(ctx, side, a->atlas, a->tile_index, a->tile_ids, a->n_tiles, a->m_tile_view, a->eps_lift, a->eps_mass,
                         &a->view, a->view_n_valid, I, a->inflate ? a->inflate_stats : nullptr, "lidar_evidence_primitives_batched",
                         (double*)ctx->ws_side);
    if (rc) return rc;
  }
  const ViewPrep vp = view_prep_carve((char*)ctx->ws_side + part_bytes, a->n_tiles, a->m_tile_view);
  rc = view_prep_launch(ctx, side, &a->view, a->n_tiles, a->m_tile_view, vp);
  if (rc) return rc;
  GCS_CHECK_CUDA(ctx, cudaEventRecord(ctx->ev_join, side));
  rc = gcs_deskew_constant_twist_batched(ctx, stream, a->pts, a->t, a->w, a->n, a->xi, a->n_units, a->scan_start_time,
                                         a->scan_end_time, a->dk_pts, a->dk_w, a->dk_cert);
  if (rc) return rc;
  if (a->base.Lambdas) {
    rc = check_mbatch(ctx, &a->base, "lidar_evidence_primitives_batched");
    if (rc) return rc;
    rc = check_mbatch(ctx, &a->batch, "lidar_evidence_primitives_batched");
    if (rc) return rc;
    GCS_REQUIRE(ctx, a->base.n_feat == a->batch.n_feat && a->base.n_surfel == a->batch.n_surfel,
                "lidar_evidence_primitives_batched: base batch budget differs from the stacked batch");
    const int N = a->batch.n_feat + a->batch.n_surfel;
    batch_replicate_kernel<<<dim3((N + 127) / 128, a->n_units), 128, 0, st>>>(a->base, a->batch, N);
    GCS_LAUNCH_CHECK(ctx);
  }
  rc = gcs_extract_lidar_surfels_batched(ctx, stream, a->dk_pts, a->t, a->dk_w, a->n, a->n_units, 1, &a->surfel_cfg, &a->batch,
                                         a->n_lidar_valid);
  if (rc) return rc;
  // join: the view and its grouped form are complete
  GCS_CHECK_CUDA(ctx, cudaStreamWaitEvent(st, ctx->ev_join, 0));
  rc = assoc_launch(ctx, st, &a->batch, a->n_units, &a->view, a->tile_ids, a->n_tiles, a->m_tile_view, &a->assoc_cfg, &a->assoc,
                    a->ot_cert, "lidar_evidence_primitives_batched", &vp);
  if (rc) return rc;
  return gcs_visual_pose_evidence_batched(ctx, stream, &a->batch, a->n_units, &a->view, &a->assoc, a->assoc_cfg.k_assoc, a->poses,
                                          a->eps_lift, a->eps_mass, a->L22, a->h22, a->rec, a->n_lidar_valid, a->n_camera_valid,
                                          a->view_n_valid);
}

int gcs_map_update(gcs_ctx* ctx, void* stream, const gcs_atlas* atlas, const int32_t* tile_index, const int64_t* tile_ids,
                   int32_t n_tiles, const gcs_meas_batch* batch, const gcs_assoc_result* assoc, const double* pose6,
                   const gcs_map_update_cfg* cfg, int64_t* out_new_ids, int32_t* out_insert_slots, double* stats) {
  if (!ctx) return GCS_EINVAL;
  GCS_CHECK_CUDA(ctx, cudaSetDevice(ctx->device));
  int rc = check_atlas(ctx, atlas, "map_update");
  if (rc) return rc;
  rc = check_mbatch(ctx, batch, "map_update");
  if (rc) return rc;
  rc = check_assoc(ctx, assoc, "map_update");
  if (rc) return rc;
  GCS_REQUIRE(ctx, tile_ids && pose6 && cfg && out_new_ids && out_insert_slots && stats, "map_update: NULL pointer");
  TileList T;
  rc = make_tile_list(ctx, atlas, tile_index, tile_ids, n_tiles, false, &T, "map_update");
  if (rc) return rc;
  const int N = batch->n_feat + batch->n_surfel, K = cfg->k_assoc, k_ins = cfg->k_insert_tile;
  GCS_REQUIRE(ctx, K >= 1 && k_ins >= 1 && k_ins <= 1024 && k_ins <= N && k_ins <= atlas->m_tile, "map_update: bad k_assoc/k_insert_tile");
  GCS_REQUIRE(ctx, (int64_t)n_tiles * atlas->m_tile < 0xffffffffll, "map_update: active tiles x m_tile overflow the 32-bit target key");
  const int n_pairs = N * K;
  GCS_REQUIRE(ctx, n_pairs <= (1 << 20), "map_update: N_total*K_ASSOC=%d exceeds the pair budget 2^20", n_pairs);
  const unsigned none = (unsigned)n_tiles * (unsigned)atlas->m_tile;   // key of a pair without a target: sorts last
  int key_bits = 1;
  while ((none >> key_bits) != 0u) ++key_bits;
  const int sort_passes = (key_bits + 7) / 8;
  const int block_rows = cfg->assoc_block_size > 0 ? cfg->assoc_block_size : 256;
  GCS_REQUIRE(ctx, block_rows * K <= 4096, "map_update: assoc_block_size*K_ASSOC exceeds 4096");
  const int n_blocks = (N + block_rows - 1) / block_rows;
  cudaStream_t st = (cudaStream_t)stream;
  const int sweep_blocks = (int)(cdivm(atlas->m_tile, 256) < 64 ? cdivm(atlas->m_tile, 256) : 64);
  const int fuse_blocks = (n_pairs + 7) / 8;   // upd_fuse_kernel: one warp per sorted position, 8 warps per block
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t o = off; off += (bytes + 255) & ~(size_t)255; return o; };
  const size_t o_Lw = take((size_t)N * 72), o_thw = take((size_t)N * 24), o_etw = take((size_t)N * 72), o_mt = take((size_t)N * 8),
               o_nov = take((size_t)N * 8), o_sc = take((size_t)N * 8), o_pk0 = take((size_t)n_pairs * 4), o_pv0 = take((size_t)n_pairs * 4),
               o_pk1 = take((size_t)n_pairs * 4), o_pv1 = take((size_t)n_pairs * 4),
               o_ii = take((size_t)n_tiles * k_ins * 4), o_in = take((size_t)n_tiles * k_ins), o_iw = take((size_t)n_tiles * k_ins * 8),
               o_is = take((size_t)n_tiles * k_ins * 4), o_ni = take(16 * 4), o_part = take(128 * 8),
               o_fpart = take((size_t)fuse_blocks * 8), o_uq = take((size_t)n_blocks * 4),
               o_mpart = take((size_t)n_tiles * sweep_blocks * 3 * 8);
  char* ws;
  rc = gcs_maint_stream(ctx, (cudaStream_t)stream, off, &st, &ws);
  if (rc) return rc;
  UpdWs W;
  W.Lw = (double*)(ws + o_Lw); W.thw = (double*)(ws + o_thw); W.etw = (double*)(ws + o_etw); W.mtile = (long long*)(ws + o_mt);
  W.novelty = (double*)(ws + o_nov); W.score = (double*)(ws + o_sc); W.pkeys = (unsigned*)(ws + o_pk1); W.pvals = (unsigned*)(ws + o_pv1);
  W.ins_idx = (int*)(ws + o_ii); W.ins_new = (uint8_t*)(ws + o_in); W.ins_w = (double*)(ws + o_iw); W.ins_slot = (int*)(ws + o_is);
  W.n_ins = (int*)(ws + o_ni); W.part = (double*)(ws + o_part);
  double* fpart = (double*)(ws + o_fpart);
  int* uq = (int*)(ws + o_uq);
  double* mpart = (double*)(ws + o_mpart);

  upd_prepare_kernel<<<1, kPrepThreads, 0, st>>>(*batch, N, *assoc, pose6[0], pose6[1], pose6[2], pose6[3], pose6[4], pose6[5], *cfg, W);
  GCS_LAUNCH_CHECK(ctx);
  upd_pair_keys_kernel<<<(n_pairs + 255) / 256, 256, 0, st>>>(*batch, N, K, *assoc, T, atlas->m_tile, none, (unsigned*)(ws + o_pk0),
                                                              (unsigned*)(ws + o_pv0));
  GCS_LAUNCH_CHECK(ctx);
  // sorted list ends in buffer 0 after an even number of passes, in buffer 1 after an odd number
  upd_sort_pairs_kernel<<<1, 32 * kSortWarps, 0, st>>>((unsigned*)(ws + o_pk0), (unsigned*)(ws + o_pv0), (unsigned*)(ws + o_pk1),
                                                      (unsigned*)(ws + o_pv1), n_pairs, sort_passes);
  GCS_LAUNCH_CHECK(ctx);
  if ((sort_passes & 1) == 0) { W.pkeys = (unsigned*)(ws + o_pk0); W.pvals = (unsigned*)(ws + o_pv0); }
  gcs_timing_begin(ctx, st, GCS_TIME_FUSE);
  upd_fuse_kernel<<<fuse_blocks, 256, 0, st>>>(*atlas, T, *batch, K, *assoc, W, n_pairs, none, *cfg, fpart);
  gcs_timing_end(ctx, st, GCS_TIME_FUSE);
  GCS_LAUNCH_CHECK(ctx);
  if (cfg->strict_tile_state) {
    upd_stamp_strict_kernel<<<(n_pairs + 255) / 256, 256, 0, st>>>(*atlas, T, *assoc, n_pairs, cfg->timestamp);
    GCS_LAUNCH_CHECK(ctx);
  }
  upd_block_unique_kernel<<<n_blocks, kBig, 0, st>>>(*assoc, N, K, block_rows, uq);
  GCS_LAUNCH_CHECK(ctx);
  upd_rgb_sweep_kernel<<<dim3(sweep_blocks, n_tiles), 256, 0, st>>>(*atlas, T, cfg->eps_mass);
  GCS_LAUNCH_CHECK(ctx);
  const size_t kc_ins = select_cache_bytes(atlas->m_tile);
  GCS_CHECK_CUDA(ctx, select_cache_attr(upd_insert_select_kernel, kc_ins));
  upd_insert_select_kernel<<<n_tiles, kBig, kc_ins, st>>>(*atlas, T, *batch, N, W, *cfg, kc_ins ? 1 : 0);
  GCS_LAUNCH_CHECK(ctx);
  upd_insert_apply_kernel<<<n_tiles, ((k_ins + 31) / 32) * 32, 0, st>>>(*atlas, T, *batch, W, *cfg, (long long*)out_new_ids, out_insert_slots);
  GCS_LAUNCH_CHECK(ctx);
  upd_maintain_kernel<<<dim3(sweep_blocks, n_tiles), 256, 0, st>>>(*atlas, T, cfg->cull_weight_threshold, cfg->forgetting_factor, mpart);
  GCS_LAUNCH_CHECK(ctx);
  upd_stats_kernel<<<1, 256, 0, st>>>(W, fpart, fuse_blocks, uq, n_blocks, mpart, sweep_blocks, n_tiles, k_ins, *cfg, stats);
  GCS_LAUNCH_CHECK(ctx);
  return GCS_OK;
}

}  // extern "C"
