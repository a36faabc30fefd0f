// gcs_select.cuh -- deterministic CTA-wide "first k of a stable sort" and small bitonic sorts.
//
// The reference selects with jnp.argsort / jax.lax.sort, which are STABLE and (for lax.sort with several operands)
// keyed on the first operand only.  "First k entries of a stable ascending sort by key" is the
// same set and order as sorting by the pair (key, original index).  cta_select_k() produces exactly that without
// sorting all n items: an 8-pass MSD radix select for the k-th key, an index-ordered compaction, then a bitonic
// sort of k (key, index) pairs in shared memory.
#pragma once
#include "gcs_common.cuh"

namespace gcs {

// order-preserving map double -> uint64 (ascending)
__device__ __forceinline__ unsigned long long f64_orderable(double x) {
  unsigned long long b = (unsigned long long)__double_as_longlong(x);
  return (b & 0x8000000000000000ull) ? ~b : (b | 0x8000000000000000ull);
}

constexpr int kSelBatch = 4;   // keys evaluated back to back per thread before their shared-memory atomics

struct KeyIdx {
  unsigned long long key;
  int idx;
  int pad;
};

__device__ __forceinline__ bool keyidx_less(const KeyIdx& a, const KeyIdx& b) {
  return a.key < b.key || (a.key == b.key && a.idx < b.idx);
}

// In-place ascending bitonic sort of n_pow2 (key, idx) pairs in shared memory by the whole CTA.
__device__ inline void cta_bitonic_sort(KeyIdx* s, int n_pow2) {
  for (int size = 2; size <= n_pow2; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      __syncthreads();
      for (int t = threadIdx.x; t < (n_pow2 >> 1); t += blockDim.x) {
        const int lo = 2 * t - (t & (stride - 1));
        const int hi = lo + stride;
        const bool up = ((lo & size) == 0);
        KeyIdx a = s[lo], b = s[hi];
        if (keyidx_less(b, a) == up) { s[lo] = b; s[hi] = a; }
      }
    }
  }
  __syncthreads();
}

// Same for bare 64-bit keys.
__device__ inline void cta_bitonic_sort_u64(unsigned long long* s, int n_pow2) {
  for (int size = 2; size <= n_pow2; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      __syncthreads();
      for (int t = threadIdx.x; t < (n_pow2 >> 1); t += blockDim.x) {
        const int lo = 2 * t - (t & (stride - 1));
        const int hi = lo + stride;
        const bool up = ((lo & size) == 0);
        const unsigned long long a = s[lo], b = s[hi];
        if ((b < a) == up) { s[lo] = b; s[hi] = a; }
      }
    }
  }
  __syncthreads();
}

// Select the k smallest (key, index) pairs among n items; result sorted ascending in out[0..k).
//   KeyFn: unsigned long long operator()(int i) const   (evaluated ~4-5 times per item)
//   out:   shared memory, capacity >= max(1024, next_pow2(k)) entries (also used as the candidate buffer)
//   hist:  unused (kept for the call sites);  scan: shared memory, >= 1024 + 64 ints
//   kc:    optional shared-memory cache of n 32-bit words (NULL: none).  The first histogram pass stores the high word
//          of every key there; the later passes and sweeps then decide from the cache and evaluate key(i) only for
//          items whose high word equals the prefix / threshold -- one evaluation per item instead of four or five
//          (the eviction key costs an exp() and three global loads per evaluation).
// Requires k <= n and k <= 1024.  All threads of the CTA must call; blockDim.x >= 32.
//
// MSD radix select with 10-bit digits.  As soon as the items that still match the prefix fit the candidate buffer
// (usually after two passes) they are gathered and sorted by (key, index), which yields the k-th key T and the index
// T_idx of the last tie that is taken; a single sweep then collects {key < T} and {key == T, index <= T_idx} in any
// order and a bitonic sort by (key, index) puts them in the order of a stable sort.  When more than 1024 items tie on
// the k-th key (e.g. a tile with fewer than k valid slots) T_idx comes from an index-ordered count of the ties.
template <bool KC, typename KeyFn>
__device__ inline void cta_select_k_impl(int n, int k, KeyFn key, KeyIdx* out, int* hist, int* scan, uint32_t* kc) {
  (void)hist;
  __shared__ unsigned long long s_prefix;
  __shared__ int s_need, s_count, s_slot, s_tidx;
  const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31;
  int* h = scan;            // 1024 bins
  int* wsum = scan + 1024;  // per-warp partials of the tie count (<= 32)
  unsigned long long prefix = 0ull, mask = 0ull;
  int need = k;             // rank (1-based) of the threshold among the items matching the prefix
  bool have_cands = false;
  for (int shift = 54; shift >= -6; shift -= 10) {
    const int sh = shift < 0 ? 0 : shift;
    const int bits = shift < 0 ? 4 : 10;
    const unsigned long long dm = (1ull << bits) - 1ull;
    for (int b = tid; b < 1024; b += nt) h[b] = 0;
    __syncthreads();
    if (KC && shift == 54) {
      // kSelBatch keys are evaluated before any of them votes: the shared-memory atomics order the memory operations
      // around them, so without the explicit batch every item pays its global-load latency on its own
      for (int i0 = tid; i0 < n; i0 += kSelBatch * nt) {
        unsigned long long kk[kSelBatch];
#pragma unroll
        for (int u = 0; u < kSelBatch; ++u) { const int i = i0 + u * nt; kk[u] = key(i < n ? i : n - 1); }   // clamped, branch-free
#pragma unroll
        for (int u = 0; u < kSelBatch; ++u) {
          const int i = i0 + u * nt;
          if (i < n) { kc[i] = (uint32_t)(kk[u] >> 32); atomicAdd(&h[(int)((kk[u] >> sh) & dm)], 1); }
        }
      }
    } else if (KC && sh >= 32) {
      // digit and prefix live in the cached high word
      const uint32_t m32 = (uint32_t)(mask >> 32), p32 = (uint32_t)(prefix >> 32);
      for (int i = tid; i < n; i += nt) {
        const uint32_t hw = kc[i];
        if ((hw & m32) == p32) atomicAdd(&h[(int)((hw >> (sh - 32)) & (uint32_t)dm)], 1);
      }
    } else {
      const uint32_t m32 = (uint32_t)(mask >> 32), p32 = (uint32_t)(prefix >> 32);
      for (int i0 = tid; i0 < n; i0 += kSelBatch * nt) {
        unsigned long long kk[kSelBatch];
        bool act[kSelBatch];
#pragma unroll
        for (int u = 0; u < kSelBatch; ++u) {
          const int i = i0 + u * nt;
          act[u] = i < n && !(KC && (kc[i] & m32) != p32);
          if (KC) kk[u] = act[u] ? key(i) : 0ull;     // few items survive the cached high word
          else kk[u] = key(i < n ? i : n - 1);        // clamped index: no branch between the loads of the batch
        }
#pragma unroll
        for (int u = 0; u < kSelBatch; ++u)
          if (act[u] && (kk[u] & mask) == prefix) atomicAdd(&h[(int)((kk[u] >> sh) & dm)], 1);
      }
    }
    __syncthreads();
    if (tid < 32) {
      // lane l owns bins [32 l, 32 l + 32): local sum, exclusive scan over lanes, then the owning lane walks its bins
      int loc = 0;
      for (int b = 0; b < 32; ++b) loc += h[32 * lane + b];
      int inc = loc;
      for (int o = 1; o < 32; o <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += v;
      }
      const int exc = inc - loc;
      if (exc < need && need <= inc) {
        int acc = exc, b = 32 * lane;
        for (;; ++b) {
          if (acc + h[b] >= need) break;
          acc += h[b];
        }
        s_prefix = prefix | ((unsigned long long)b << sh);
        s_need = need - acc;
        s_count = h[b];
      }
    }
    __syncthreads();
    prefix = s_prefix;
    need = s_need;
    mask |= (dm << sh);
    const int m = s_count;
    __syncthreads();
    if (m <= 1024) { have_cands = true; break; }
  }
  unsigned long long T = prefix;
  int T_idx = 0x7fffffff;
  if (have_cands) {
    if (tid == 0) s_slot = 0;
    __syncthreads();
    {
      const uint32_t m32 = (uint32_t)(mask >> 32), p32 = (uint32_t)(prefix >> 32);
      for (int i0 = tid; i0 < n; i0 += kSelBatch * nt) {
        unsigned long long kk[kSelBatch];
        bool act[kSelBatch];
#pragma unroll
        for (int u = 0; u < kSelBatch; ++u) {
          const int i = i0 + u * nt;
          act[u] = i < n && !(KC && (kc[i] & m32) != p32);
          if (KC) kk[u] = act[u] ? key(i) : 0ull;     // few items survive the cached high word
          else kk[u] = key(i < n ? i : n - 1);        // clamped index: no branch between the loads of the batch
        }
#pragma unroll
        for (int u = 0; u < kSelBatch; ++u) {
          if (act[u] && (kk[u] & mask) == prefix) {
            const int sl = atomicAdd(&s_slot, 1);
            out[sl].key = kk[u]; out[sl].idx = i0 + u * nt;
          }
        }
      }
    }
    __syncthreads();
    const int m = s_slot;
    int mp = 1;
    while (mp < m) mp <<= 1;
    for (int i = m + tid; i < mp; i += nt) { out[i].key = ~0ull; out[i].idx = 0x7fffffff; }
    __syncthreads();
    cta_bitonic_sort(out, mp);
    T = out[need - 1].key;
    T_idx = out[need - 1].idx;
    __syncthreads();
  } else {
    // more than 1024 items equal the k-th key: take the first `need` of them in index order
    const int chunk = (n + nt - 1) / nt;
    const int c0 = tid * chunk, c1 = (c0 + chunk < n) ? c0 + chunk : n;
    int n_eq = 0;
    const uint32_t t32 = (uint32_t)(T >> 32);
    for (int i = c0; i < c1; ++i) n_eq += ((!KC || kc[i] == t32) && key(i) == T);
    int inc = n_eq;
    for (int o = 1; o < 32; o <<= 1) {
      const int v = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane >= o) inc += v;
    }
    if (lane == 31) wsum[tid >> 5] = inc;
    __syncthreads();
    int before = inc - n_eq;
    for (int w = 0; w < (tid >> 5); ++w) before += wsum[w];
    if (before < need && need <= before + n_eq) {
      int r = before;
      for (int i = c0; i < c1; ++i)
        if ((!KC || kc[i] == t32) && key(i) == T && ++r == need) { s_tidx = i; break; }
    }
    __syncthreads();
    T_idx = s_tidx;
  }
  if (tid == 0) s_slot = 0;
  __syncthreads();
  {
    const uint32_t t32f = (uint32_t)(T >> 32);
    for (int i0 = tid; i0 < n; i0 += kSelBatch * nt) {
      unsigned long long kk[kSelBatch];
      bool act[kSelBatch];
#pragma unroll
      for (int u = 0; u < kSelBatch; ++u) {
        const int i = i0 + u * nt;
        act[u] = i < n && !(KC && kc[i] > t32f);    // high word above the threshold's: key > T
        if (KC) kk[u] = act[u] ? key(i) : 0ull;
        else kk[u] = key(i < n ? i : n - 1);
      }
#pragma unroll
      for (int u = 0; u < kSelBatch; ++u) {
        const int i = i0 + u * nt;
        if (act[u] && (kk[u] < T || (kk[u] == T && i <= T_idx))) {
          const int sl = atomicAdd(&s_slot, 1);
          out[sl].key = kk[u]; out[sl].idx = i;
        }
      }
    }
  }
  int kp = 1;
  while (kp < k) kp <<= 1;
  __syncthreads();
  for (int i = k + tid; i < kp; i += nt) { out[i].key = ~0ull; out[i].idx = 0x7fffffff; }
  __syncthreads();
  cta_bitonic_sort(out, kp);
}

template <typename KeyFn>
__device__ inline void cta_select_k(int n, int k, KeyFn key, KeyIdx* out, int* hist, int* scan, uint32_t* kc = nullptr) {
  if (kc) cta_select_k_impl<true>(n, k, key, out, hist, scan, kc);
  else cta_select_k_impl<false>(n, k, key, out, hist, scan, kc);
}

// Fast path of a select whose smallest possible key `smin` is shared by many items (the empty slots of a tile in the
// eviction select: key -inf): if at least k items are flagged by sent(i), the result of the stable sort is the first k
// of them in index order -- one counting sweep and a partial second one instead of the radix passes, every one of which
// would have to walk all the tied items again.  Returns false (out untouched) if fewer than k items are flagged; all
// threads of the CTA must call and get the same answer.  scan: shared memory, >= 64 ints.
template <typename SentFn>
__device__ inline bool cta_select_min_sentinel(int n, int k, SentFn sent, unsigned long long smin, KeyIdx* out, int* scan) {
  const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, nw = nt >> 5;
  int* wsum = scan;   // per-warp totals (<= 32), then [32] = grand total
  const int chunk = (n + nt - 1) / nt;
  const int c0 = tid * chunk < n ? tid * chunk : n, c1 = (c0 + chunk < n) ? c0 + chunk : n;
  int cnt = 0;
#pragma unroll 8
  for (int i = c0; i < c1; ++i) cnt += sent(i) ? 1 : 0;
  int inc = cnt;
  for (int o = 1; o < 32; o <<= 1) {
    const int v = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += v;
  }
  __syncthreads();   // scan may still be in use by the caller's previous select
  if (lane == 31) wsum[tid >> 5] = inc;
  __syncthreads();
  int before = inc - cnt, total = 0;
  for (int w = 0; w < nw; ++w) {
    const int v = wsum[w];
    if (w < (tid >> 5)) before += v;
    total += v;
  }
  __syncthreads();   // wsum is free again
  if (total < k) return false;
  if (cnt > 0 && before < k) {
    int r = before;
    for (int i = c0; i < c1 && r < k; ++i)
      if (sent(i)) { out[r].key = smin; out[r].idx = i; ++r; }
  }
  __syncthreads();
  return true;
}

// Fast path from the other end: the largest possible key `smax` is shared by many items (the invalid slots of a tile in
// the view select) and fewer than k items carry a real key.  The result of the stable sort is then all real items in
// (key, index) order followed by the first k - n_real sentinel items in index order; the radix passes could never
// separate the tied sentinels and would run all seven digits plus the tie count.  Returns false (out untouched) if at
// least k items are real; all threads of the CTA must call and get the same answer.  Requires k <= 1024, out capacity
// >= 1024 entries; scan: shared memory, >= 64 ints.
template <typename KeyFn, typename SentFn>
__device__ inline bool cta_select_max_sentinel(int n, int k, KeyFn key, SentFn sent, unsigned long long smax, KeyIdx* out, int* scan) {
  __shared__ int s_real_slot;
  const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, nw = nt >> 5;
  int* wsum = scan;
  const int chunk = (n + nt - 1) / nt;
  const int c0 = tid * chunk < n ? tid * chunk : n, c1 = (c0 + chunk < n) ? c0 + chunk : n;
  int cnt = 0;   // sentinel items of this thread's chunk
#pragma unroll 8
  for (int i = c0; i < c1; ++i) cnt += sent(i) ? 1 : 0;
  int inc = cnt;
  for (int o = 1; o < 32; o <<= 1) {
    const int v = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += v;
  }
  __syncthreads();
  if (lane == 31) wsum[tid >> 5] = inc;
  if (tid == 0) s_real_slot = 0;
  __syncthreads();
  int before = inc - cnt, total = 0;
  for (int w = 0; w < nw; ++w) {
    const int v = wsum[w];
    if (w < (tid >> 5)) before += v;
    total += v;
  }
  __syncthreads();
  const int n_real = n - total;
  if (n_real >= k) return false;
  // real items, any order, then sorted by (key, index)
  for (int i = c0; i < c1; ++i)
    if (!sent(i)) {
      const int sl = atomicAdd(&s_real_slot, 1);
      out[sl].key = key(i); out[sl].idx = i;
    }
  int np = 1;
  while (np < n_real) np <<= 1;
  __syncthreads();
  for (int i = n_real + tid; i < np; i += nt) { out[i].key = ~0ull; out[i].idx = 0x7fffffff; }
  __syncthreads();
  if (n_real > 1) cta_bitonic_sort(out, np);
  // the first k - n_real sentinel items in index order go behind them
  const int want = k - n_real;
  if (cnt > 0 && before < want) {
    int r = before;
    for (int i = c0; i < c1 && r < want; ++i)
      if (sent(i)) { out[n_real + r].key = smax; out[n_real + r].idx = i; ++r; }
  }
  __syncthreads();
  return true;
}

}  // namespace gcs
