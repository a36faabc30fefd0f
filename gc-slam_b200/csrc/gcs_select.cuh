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
//   KeyFn: unsigned long long operator()(int i) const   (must be cheap: it is evaluated ~10 times per item)
//   out:   shared memory, capacity >= next_pow2(k)
//   hist:  shared memory, 256 ints;  scan: shared memory, 2*blockDim.x ints
// Requires k <= n.  All threads of the CTA must call.
template <typename KeyFn>
__device__ inline void cta_select_k(int n, int k, KeyFn key, KeyIdx* out, int* hist, int* scan) {
  __shared__ unsigned long long s_prefix;
  __shared__ int s_need;
  const int tid = threadIdx.x, nt = blockDim.x;
  unsigned long long prefix = 0ull, mask = 0ull;
  int need = k;  // rank (1-based) of the threshold among items matching the prefix
  for (int pass = 7; pass >= 0; --pass) {
    for (int b = tid; b < 256; b += nt) hist[b] = 0;
    __syncthreads();
    const int shift = pass * 8;
    for (int i = tid; i < n; i += nt) {
      const unsigned long long kk = key(i);
      if ((kk & mask) == prefix) atomicAdd(&hist[(int)((kk >> shift) & 0xffull)], 1);
    }
    __syncthreads();
    if (tid == 0) {
      int acc = 0, b = 0;
      for (; b < 256; ++b) {
        if (acc + hist[b] >= need) break;
        acc += hist[b];
      }
      s_prefix = prefix | ((unsigned long long)b << shift);
      s_need = need - acc;
    }
    __syncthreads();
    prefix = s_prefix;
    need = s_need;
    mask |= (0xffull << shift);
    __syncthreads();
  }
  const unsigned long long T = prefix;  // k-th smallest key; `need` = how many items equal to T are taken
  // index-ordered compaction: thread t owns the contiguous chunk [c0, c1)
  const int chunk = (n + nt - 1) / nt;
  const int c0 = tid * chunk, c1 = (c0 + chunk < n) ? c0 + chunk : n;
  int n_less = 0, n_eq = 0;
  for (int i = c0; i < c1; ++i) {
    const unsigned long long kk = key(i);
    n_less += (kk < T);
    n_eq += (kk == T);
  }
  scan[tid] = n_less;
  scan[nt + tid] = n_eq;
  __syncthreads();
  if (tid == 0) {
    int a = 0, b = 0;
    for (int t = 0; t < nt; ++t) {
      const int x = scan[t], y = scan[nt + t];
      scan[t] = a; scan[nt + t] = b;
      a += x; b += y;
    }
    s_need = a;  // total number of keys strictly below T  (== k - need)
  }
  __syncthreads();
  const int total_less = s_need;
  int o_less = scan[tid], o_eq = scan[nt + tid];
  for (int i = c0; i < c1; ++i) {
    const unsigned long long kk = key(i);
    if (kk < T) {
      out[o_less].key = kk; out[o_less].idx = i; ++o_less;
    } else if (kk == T) {
      if (o_eq < need) { out[total_less + o_eq].key = kk; out[total_less + o_eq].idx = i; }
      ++o_eq;
    }
  }
  int kp = 1;
  while (kp < k) kp <<= 1;
  for (int i = k + tid; i < kp; i += nt) { out[i].key = ~0ull; out[i].idx = 0x7fffffff; }
  __syncthreads();
  cta_bitonic_sort(out, kp);
}

}  // namespace gcs
