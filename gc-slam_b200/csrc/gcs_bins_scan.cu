// gcs_bins_scan.cu -- per-point kernels of the bin family:
//   resample masses, the fused resample+deskew+direction+soft-assign+moments kernel, and the stand-alone
//   per-operator kernels (gather, deskew, directions, soft-assign, moments-from-responsibilities).
//
// Mapping of the hot kernel (bin_scan_kernel): a CTA owns a tile of 128 output rows.
//   phase 1  one thread per point: strided gather of the raw point (PointBudgetResample), constant-twist SE(3)
//            deskew, time-window weight, ray direction; the 19 moment features go to shared memory.
//   phase 2  one half-warp per point, lane l owns bins {l, l+16, l+32, ..}: logits, exp, a 4-step shuffle sum for
//            the softmax normaliser, then 19 FMAs per owned bin into register accumulators that live for the
//            whole kernel.  No (N,B) responsibility matrix ever reaches memory unless the caller asks for it.
// All cross-thread reductions are fixed-shape trees / fixed-order loops => bit-identical results run to run.
#include <stdlib.h>

#include "gcs_bins.cuh"

namespace gcs {


__device__ __forceinline__ double hw_sum(double v) {
  v += __shfl_xor_sync(0xffffffffu, v, 8);
  v += __shfl_xor_sync(0xffffffffu, v, 4);
  v += __shfl_xor_sync(0xffffffffu, v, 2);
  v += __shfl_xor_sync(0xffffffffu, v, 1);
  return v;
}
__device__ __forceinline__ double hw_max(double v) {
  v = fmax(v, __shfl_xor_sync(0xffffffffu, v, 8));
  v = fmax(v, __shfl_xor_sync(0xffffffffu, v, 4));
  v = fmax(v, __shfl_xor_sync(0xffffffffu, v, 2));
  v = fmax(v, __shfl_xor_sync(0xffffffffu, v, 1));
  return v;
}

constexpr double kLog2e = 1.4426950408889634;
constexpr double kLn2 = 0.6931471805599453;

// Block-wide fixed-order reduction of per-half-warp register accumulators into partial[].
template <int Q, int NF>
__device__ __forceinline__ void reduce_acc_to_partial(double (&acc)[Q][NF], double* sbuf, double* partial, int n_bins) {
  const int tid = threadIdx.x, l16 = tid & 15, hw = tid >> 4;
  const int n_hw = blockDim.x >> 4;
#pragma unroll
  for (int q = 0; q < Q; ++q) {
    __syncthreads();
#pragma unroll
    for (int f = 0; f < NF; ++f) sbuf[(hw * 16 + l16) * NF + f] = acc[q][f];
    __syncthreads();
    for (int idx = tid; idx < 16 * NF; idx += blockDim.x) {
      int l = idx / NF, f = idx - l * NF;
      double s = 0.0;
      for (int g = 0; g < n_hw; ++g) s += sbuf[(g * 16 + l) * NF + f];
      int b = q * 16 + l;
      if (b < n_bins) partial[b * kRowLen + f] = s;
    }
  }
  __syncthreads();
}

// Same for the feature-split accumulators of bin_scan_kernel: thread (warp w, half h, lane l16) holds features
// [10h, 10h+10) of bins {l16 + 16q}.  Warps are combined in index order.
template <int Q>
__device__ __forceinline__ void reduce_split_acc_to_partial(double (&acc)[Q][kHalfF], double* sbuf, double* partial,
                                                            int n_bins) {
  const int tid = threadIdx.x, l16 = tid & 15, half = (tid >> 4) & 1, wid = tid >> 5;
  const int n_w = blockDim.x >> 5;
#pragma unroll
  for (int q = 0; q < Q; ++q) {
    __syncthreads();
#pragma unroll
    for (int f = 0; f < kHalfF; ++f) sbuf[(wid * 16 + l16) * kFeatStride + half * kHalfF + f] = acc[q][f];
    __syncthreads();
    for (int idx = tid; idx < 16 * kNF; idx += blockDim.x) {
      const int l = idx / kNF, f = idx - l * kNF;
      double s = 0.0;
      for (int g = 0; g < n_w; ++g) s += sbuf[(g * 16 + l) * kFeatStride + f];
      const int b = q * 16 + l;
      if (b < n_bins) partial[b * kRowLen + f] = s;
    }
  }
  __syncthreads();
}

// Block-wide fixed-order sum / max of one scalar per thread.  Result valid in thread 0.
__device__ __forceinline__ double block_sum_fixed(double v, double* sred) {
  v = warp_sum(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) sred[threadIdx.x >> 5] = v;
  __syncthreads();
  double s = 0.0;
  if (threadIdx.x == 0)
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) s += sred[i];
  return s;
}
__device__ __forceinline__ double block_max_fixed(double v, double* sred) {
  v = warp_max(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) sred[threadIdx.x >> 5] = v;
  __syncthreads();
  double s = 0.0;
  if (threadIdx.x == 0)
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) s = fmax(s, sred[i]);
  return s;
}

// ------------------------------------------------------------------------------------------------
// Resample masses (a1).  grid (chunks, S); block 256.  partial (S, chunks, 4) -> mass_final.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) mass_partial_kernel(const double* __restrict__ w, int64_t n_raw, int64_t stride,
                                                           int64_t rows_per_chunk, double* __restrict__ partial) {
  __shared__ double sred[8];
  const int s = blockIdx.y, c = blockIdx.x;
  const double* ws = w + (int64_t)s * n_raw;
  const int64_t r0 = (int64_t)c * rows_per_chunk;
  int64_t r1 = r0 + rows_per_chunk;
  if (r1 > n_raw) r1 = n_raw;
  const int n = (int)(r1 > r0 ? r1 - r0 : 0);          // rows of this chunk (chunks are < 2^31 rows)
  const double* wc = ws + r0;
  double a = 0.0, b = 0.0, q = 0.0, cnt = 0.0;
  if (stride == 1) {
    // every row is selected: one pass, 16-byte loads when the chunk is aligned.  Each thread keeps two partial sums
    // (even / odd element of its pairs) combined in a fixed order below.
    double a1 = 0.0, q1 = 0.0;
    if ((reinterpret_cast<uintptr_t>(wc) & 15) == 0) {
      const double2* w2 = reinterpret_cast<const double2*>(wc);
      const int n2 = n >> 1;
#pragma unroll 8
      for (int j = threadIdx.x; j < n2; j += 256) {
        const double2 v = __ldg(w2 + j);
        a += v.x; q = fma(v.x, v.x, q);
        a1 += v.y; q1 = fma(v.y, v.y, q1);
      }
      if ((n & 1) && threadIdx.x == 0) { const double v = wc[n - 1]; a += v; q = fma(v, v, q); }
    } else {
#pragma unroll 4
      for (int j = threadIdx.x; j < n; j += 256) { const double v = wc[j]; a += v; q = fma(v, v, q); }
    }
    a += a1; q += q1;
    b = a;
    cnt = 0.0;   // filled in by thread 0 below (exact integer)
  } else {
    // rows r0 + j with (r0 + j) % stride == 0 are selected: 32-bit phase arithmetic inside the chunk
    const uint32_t st = (uint32_t)(stride > 0x7fffffff ? 0x7fffffff : stride);
    const uint32_t ph0 = (uint32_t)(r0 % stride);        // phase of the first row of the chunk
    for (int j = threadIdx.x; j < n; j += 256) {
      const double v = wc[j];
      a += v;
      const bool sel = stride > 0x7fffffff ? (r0 + j) % stride == 0 : ((ph0 + (uint32_t)j) % st) == 0;
      if (sel) { b += v; q = fma(v, v, q); cnt += 1.0; }
    }
  }
  // one fixed-order block reduction for the four sums
  __shared__ double s4[4][8];
  a = warp_sum(a); b = warp_sum(b); q = warp_sum(q); cnt = warp_sum(cnt);
  if ((threadIdx.x & 31) == 0) {
    const int wi = threadIdx.x >> 5;
    s4[0][wi] = a; s4[1][wi] = b; s4[2][wi] = q; s4[3][wi] = cnt;
  }
  __syncthreads();
  if (threadIdx.x < 4) {
    double r = 0.0;
    for (int i = 0; i < 8; ++i) r += s4[threadIdx.x][i];
    if (threadIdx.x == kMassNSel && stride == 1) r = (double)n;
    partial[((int64_t)s * gridDim.x + c) * kNMass + threadIdx.x] = r;
  }
  (void)sred;
}
__global__ void mass_final_kernel(const double* __restrict__ partial, int chunks, int n_scans, double* __restrict__ mass) {
  int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n_scans * kNMass) return;
  int s = idx / kNMass, k = idx - s * kNMass;
  double acc = 0.0;
  for (int c = 0; c < chunks; ++c) acc += partial[((int64_t)s * chunks + c) * kNMass + k];
  mass[idx] = acc;
}
// cert of the stand-alone operator: [mass_in, mass_sel, sumsq_sel, ess, mass_scale]
__global__ void resample_cert_kernel(const double* __restrict__ mass, int64_t cap_total, double eps_mass,
                                     double* __restrict__ cert) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  double M = mass[kMassAll], Ms = mass[kMassSel], Q = mass[kMassSelSq];
  double scale = M / (Ms + eps_mass);
  double g = scale / (M + eps_mass);
  double ess = 1.0 / (g * g * Q + (double)cap_total * eps_mass);  // eps inside the sum: point_budget.py:96
  cert[GCS_RS_MASS_IN] = M; cert[GCS_RS_MASS_SEL] = Ms; cert[GCS_RS_SUMSQ_SEL] = Q;
  cert[GCS_RS_ESS] = ess; cert[GCS_RS_MASS_SCALE] = scale;
}

__global__ void __launch_bounds__(256) resample_gather_kernel(
    const double* __restrict__ pts, const double* __restrict__ t, const double* __restrict__ w,
    const uint8_t* __restrict__ ring, const uint8_t* __restrict__ tag, int64_t n_sel, int64_t cap, int64_t stride,
    const double* __restrict__ mass, double eps_mass, double* __restrict__ o_pts, double* __restrict__ o_t,
    double* __restrict__ o_w, uint8_t* __restrict__ o_ring, uint8_t* __restrict__ o_tag) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= cap) return;
  double scale = mass[kMassAll] / (mass[kMassSel] + eps_mass);
  double p0 = 0, p1 = 0, p2 = 0, tt = 0, ww = 0;
  uint8_t rg = 0, tg = 0;
  if (i < n_sel) {
    int64_t j = i * stride;
    p0 = pts[3 * j]; p1 = pts[3 * j + 1]; p2 = pts[3 * j + 2];
    tt = t[j]; ww = w[j] * scale;
    if (ring) rg = ring[j];
    if (tag) tg = tag[j];
  }
  o_pts[3 * i] = p0; o_pts[3 * i + 1] = p1; o_pts[3 * i + 2] = p2;
  o_t[i] = tt; o_w[i] = ww; o_ring[i] = rg; o_tag[i] = tg;
}

// ------------------------------------------------------------------------------------------------
// Fused hot kernel.
//   phase 1 (thread per point): gather + deskew + window weight + direction + 19 features -> sF;
//            then the whole softmax row of that point: 48 independent exp() (ILP across bins), unnormalised
//            e_b -> sW[point][b]; scale = w/sum and 1/sum go to sF.
//   phase 2 (half-warp per point, lane l owns bins l, l+16, l+32): wr = e * scale; 19 FMAs per owned bin into
//            register accumulators.  No transcendental, no shuffle: pure FMA stream with loads from smem.
// ------------------------------------------------------------------------------------------------
template <int Q>
struct ScanSmem {
  static constexpr int kWStride = 16 * Q + 1;  // odd stride (in doubles): conflict-free row writes and column reads
  double F[kScanThreads * kFeatStride];
  double W[kScanThreads * kWStride];
  double B[kMaxBins * 3];   // bin directions * log2(e)/tau  (logits in log2 units: softmax is base invariant)
  float Bf[kMaxBins * 4];   // float32 copy for the mixed-precision soft-assign
  double red[kScanThreads / 32];
};

template <int Q, int PREC, int OCC>
__global__ void __launch_bounds__(kScanThreads, OCC) bin_scan_kernel(const BinScanParams P) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  ScanSmem<Q>& sm = *reinterpret_cast<ScanSmem<Q>*>(smem_raw);
  constexpr int WS = ScanSmem<Q>::kWStride;
  const int tid = threadIdx.x, l16 = tid & 15, hw = tid >> 4;
  const int u = blockIdx.y;
  const int s = u / P.n_hyp, h = u - s * P.n_hyp;
  const int nb = P.n_bins;

  for (int k = tid; k < nb * 3; k += kScanThreads) {
    const double v = P.bin_dirs[k] * (P.inv_tau * kLog2e);
    sm.B[k] = v;
    sm.Bf[(k / 3) * 4 + (k % 3)] = (float)v;
  }

  bool bok[Q];
#pragma unroll
  for (int q = 0; q < Q; ++q) bok[q] = (q * 16 + l16) < nb;
  // phase-2 accumulators: a WARP owns a point; lanes 0-15 carry features 0..9 (N, s_dir, S_scatter), lanes 16-31
  // features 10..18 (sum_p, sum_ppT; slot 19 unused) of bins {l16, l16+16, ..}.  30 doubles per thread instead of 57.
  const int half = (tid >> 4) & 1, wid = tid >> 5;
  double acc[Q][kHalfF];
#pragma unroll
  for (int q = 0; q < Q; ++q)
#pragma unroll
    for (int f = 0; f < kHalfF; ++f) acc[q][f] = 0.0;
  double ent_dot = 0.0, ent_log = 0.0, mx_resp = 0.0, sum_wdk = 0.0, sum_wrs = 0.0, n_rows = 0.0;

  const double t0 = P.t0s[s], t1 = P.t1s[s];
  const double inv_denom = 1.0 / fmax(t1 - t0, 1e-12);
  const double inv_sig = window_inv_sigma(t0, t1);
  double xi[6];
#pragma unroll
  for (int k = 0; k < 6; ++k) xi[k] = P.xi[(int64_t)u * 6 + k];
  const double mass_scale = P.mass[s * kNMass + kMassAll] / (P.mass[s * kNMass + kMassSel] + P.eps_mass);
  const double shift2 = P.shift * kLog2e;
  const double* pts = P.pts + (int64_t)s * P.n_raw * 3;
  const double* tp = P.t + (int64_t)s * P.n_raw;
  const double* wp = P.w + (int64_t)s * P.n_raw;
  const uint8_t* rp = P.ring ? P.ring + (int64_t)s * P.n_raw : nullptr;
  const uint8_t* gp = P.tag ? P.tag + (int64_t)s * P.n_raw : nullptr;
  __syncthreads();

  const int64_t n_tiles = (P.cap + kScanThreads - 1) / kScanThreads;
  for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    // ---------------- phase 1: one thread per output row
    const int64_t i = tile * kScanThreads + tid;
    const bool row = i < P.cap;
    double p[3] = {0.0, 0.0, 0.0}, tt = 0.0, ww = 0.0;
    uint8_t rg = 0, tg = 0;
    if (i < P.n_sel) {
      const int64_t j = i * P.stride;
      p[0] = pts[3 * j]; p[1] = pts[3 * j + 1]; p[2] = pts[3 * j + 2];
      tt = tp[j]; ww = wp[j];
      if (rp) rg = rp[j];
      if (gp) tg = gp[j];
    }
    const double w_rs = ww * mass_scale;
    if (row && h == 0 && P.rs_pts) {
      const int64_t o = (int64_t)s * P.cap + i;
      P.rs_pts[3 * o] = p[0]; P.rs_pts[3 * o + 1] = p[1]; P.rs_pts[3 * o + 2] = p[2];
      P.rs_t[o] = tt; P.rs_w[o] = w_rs; P.rs_ring[o] = rg; P.rs_tag[o] = tg;
    }
    const double alpha = (tt - t0) * inv_denom;
    double p0[3];
    deskew_point(p, alpha, xi, p0);
    const double w_dk = w_rs * window_weight(tt, t0, t1, inv_sig);
    if (row) {
      const int64_t o = (int64_t)u * P.cap + i;
      if (P.dk_pts) { P.dk_pts[3 * o] = p0[0]; P.dk_pts[3 * o + 1] = p0[1]; P.dk_pts[3 * o + 2] = p0[2]; }
      if (P.dk_w) P.dk_w[o] = w_dk;
      sum_wdk += w_dk; sum_wrs += w_rs; n_rows += 1.0;
    }
    {
      const double r0 = p0[0] - P.origin[0], r1 = p0[1] - P.origin[1], r2 = p0[2] - P.origin[2];
      const double nrm = sqrt(r0 * r0 + r1 * r1 + r2 * r2);
      const double invn = 1.0 / (nrm + P.eps_mass);
      const double d0 = r0 * invn, d1 = r1 * invn, d2 = r2 * invn;
      // ---- softmax row (BinSoftAssign) in log2 units: e_b = 2^(t_b - c2), unnormalised, to sW[tid][b]
      double* wrow = &sm.W[tid * WS];
      double ssum = 0.0, dot = 0.0, emax = 0.0;
      if (PREC == 0) {
        double c2 = shift2;
        if (P.use_true_max) {
          double m = -1.0e300;
          for (int b = 0; b < nb; ++b) m = fmax(m, fma(d0, sm.B[3 * b], fma(d1, sm.B[3 * b + 1], d2 * sm.B[3 * b + 2])));
          c2 = m;
        }
        for (int b0 = 0; b0 < nb; b0 += 8) {
          double a[8], e[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const int b = (b0 + j < nb) ? b0 + j : nb - 1;  // clamp: tail lanes recompute the last bin, masked below
            a[j] = fma(d0, sm.B[3 * b], fma(d1, sm.B[3 * b + 1], fma(d2, sm.B[3 * b + 2], -c2)));
          }
          exp2_nonpos_x8(a, e);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            if (b0 + j < nb) {
              wrow[b0 + j] = e[j];
              ssum += e[j];
              dot = fma(e[j], a[j], dot);
              emax = fmax(emax, e[j]);
            }
          }
        }
      } else {
        // mixed precision: logits, MUFU ex2 and the entropy dot in float32; sum and everything downstream in float64
        const float f0 = (float)d0, f1 = (float)d1, f2 = (float)d2;
        float c2 = (float)shift2;
        if (P.use_true_max) {
          float m = -3.0e38f;
          for (int b = 0; b < nb; ++b) m = fmaxf(m, fmaf(f0, sm.Bf[4 * b], fmaf(f1, sm.Bf[4 * b + 1], f2 * sm.Bf[4 * b + 2])));
          c2 = m;
        }
        float dotf = 0.f, emaxf = 0.f;
        for (int b0 = 0; b0 < nb; b0 += 8) {
          float a[8], e[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const int b = (b0 + j < nb) ? b0 + j : nb - 1;
            const float4 bb = *reinterpret_cast<const float4*>(&sm.Bf[4 * b]);
            a[j] = fmaf(f0, bb.x, fmaf(f1, bb.y, fmaf(f2, bb.z, -c2)));
          }
#pragma unroll
          for (int j = 0; j < 8; ++j) asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e[j]) : "f"(a[j]));
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            if (b0 + j < nb) {
              const double ed = f32_bits_to_f64(e[j]);
              wrow[b0 + j] = ed;
              ssum += ed;
              dotf = fmaf(e[j], a[j], dotf);
              emaxf = fmaxf(emaxf, e[j]);
            }
          }
        }
        dot = (double)dotf;
        emax = (double)emaxf;
      }
      const double inv = 1.0 / ssum;
      if (row) {
        // natural-log entropy from log2-domain quantities: H = ln2 * (log2(sum) - sum_b r_b (t_b - c2))
        ent_dot = fma(inv * kLn2, dot, ent_dot);
        ent_log += log(ssum);
        mx_resp = fmax(mx_resp, emax * inv);
      }
      double2* F = reinterpret_cast<double2*>(&sm.F[tid * kFeatStride]);
      F[0] = make_double2(row ? w_dk * inv : 0.0, d0);
      F[1] = make_double2(d1, d2);
      F[2] = make_double2(d0 * d0, d0 * d1);
      F[3] = make_double2(d0 * d2, d1 * d1);
      F[4] = make_double2(d1 * d2, d2 * d2);
      F[5] = make_double2(p0[0], p0[1]);
      F[6] = make_double2(p0[2], p0[0] * p0[0]);
      F[7] = make_double2(p0[0] * p0[1], p0[0] * p0[2]);
      F[8] = make_double2(p0[1] * p0[1], p0[1] * p0[2]);
      F[9] = make_double2(p0[2] * p0[2], row ? inv : 0.0);
    }
    __syncthreads();

    // ---------------- phase 2: one warp per point, pure FMA stream
#pragma unroll 4
    for (int it = 0; it < kScanThreads / (kScanThreads / 32); ++it) {
      const int k = wid + (kScanThreads / 32) * it;
      const double scale = sm.F[k * kFeatStride];
      const double2* F = reinterpret_cast<const double2*>(&sm.F[k * kFeatStride + kHalfF * half]);
      double g[kHalfF];
#pragma unroll
      for (int v = 0; v < kHalfF / 2; ++v) { double2 x2 = F[v]; g[2 * v] = x2.x; g[2 * v + 1] = x2.y; }
      if (half == 0) g[0] = 1.0;       // feature 0 is the mass itself (slot 0 of F carries w/sum)
      else g[kHalfF - 1] = 0.0;        // slot 19 carries 1/sum, not a feature
      const double* wrow = &sm.W[k * WS];
#pragma unroll
      for (int q = 0; q < Q; ++q) {
        const double wr = bok[q] ? wrow[q * 16 + l16] * scale : 0.0;
#pragma unroll
        for (int ff = 0; ff < kHalfF; ++ff) acc[q][ff] = fma(wr, g[ff], acc[q][ff]);
      }
    }
    if (P.resp) {
      // materialise responsibilities of this tile: rows are contiguous in memory -> coalesced copy-out
      double* ro = P.resp + ((int64_t)u * P.cap + tile * kScanThreads) * nb;
      const int64_t rows_here = (P.cap - tile * kScanThreads) < kScanThreads ? (P.cap - tile * kScanThreads) : kScanThreads;
      for (int64_t idx = tid; idx < rows_here * nb; idx += kScanThreads) {
        const int k = (int)(idx / nb), b = (int)(idx - (int64_t)k * nb);
        ro[idx] = sm.W[k * WS + b] * sm.F[k * kFeatStride + 19];
      }
    }
    __syncthreads();
  }

  double* part = P.partial + ((int64_t)u * gridDim.x + blockIdx.x) * P.part_len;
  reduce_split_acc_to_partial<Q>(acc, sm.W, part, nb);
  double* ex = part + nb * kRowLen;
  double v;
  v = block_sum_fixed(ent_dot, sm.red); if (tid == 0) ex[kExEntDot] = v;
  v = block_sum_fixed(ent_log, sm.red); if (tid == 0) ex[kExEntLog] = v;
  v = block_sum_fixed(sum_wdk, sm.red); if (tid == 0) ex[kExSumWdk] = v;
  v = block_sum_fixed(sum_wrs, sm.red); if (tid == 0) ex[kExSumWrs] = v;
  v = block_sum_fixed(n_rows, sm.red);  if (tid == 0) { ex[kExCount] = v; ex[5] = 0.0; ex[6] = 0.0; ex[7] = 0.0; }
  v = block_max_fixed(mx_resp, sm.red); if (tid == 0) { ex[kNExtras + kMxResp] = v; ex[kNExtras + 1] = 0.0; }
}

// partial (U, n_parts, part_len) -> raw_sums (U, raw_len), raw_max (U, kNMax).  Columns >= nf of each bin row are zeroed.
// grid (ceil((raw_len+kNMax)/64), U); block 256 = 64 columns x 4 part-groups.  Each thread adds its parts in index
// order, the 4 groups are combined in fixed order: deterministic, and n_parts/4 sequential loads instead of n_parts.
__global__ void __launch_bounds__(256) reduce_partials_kernel(const double* __restrict__ partial, int n_parts, int part_len,
                                                              int n_bins, int nf, double* __restrict__ raw_sums,
                                                              double* __restrict__ raw_max) {
  __shared__ double sg[4][64];
  const int u = blockIdx.y;
  const int raw_len = raw_sums_len(n_bins);
  const int col = threadIdx.x & 63, grp = threadIdx.x >> 6;
  const int idx = blockIdx.x * 64 + col;
  const bool in = idx < raw_len + kNMax;
  const bool is_max = idx >= raw_len;
  bool skip = !in;
  if (in && idx < n_bins * kRowLen) skip = (idx % kRowLen) >= nf;
  const double* pu = partial + (int64_t)u * n_parts * part_len + idx;
  double a = 0.0;
  if (!skip) {
    const int per = (n_parts + 3) / 4;
    const int c0 = grp * per, c1 = (c0 + per < n_parts) ? c0 + per : n_parts;
#pragma unroll 8
    for (int c = c0; c < c1; ++c) {
      const double v = pu[(int64_t)c * part_len];
      a = is_max ? fmax(a, v) : a + v;
    }
  }
  sg[grp][col] = a;
  __syncthreads();
  if (grp == 0 && in) {
    double r = sg[0][col];
    for (int g = 1; g < 4; ++g) r = is_max ? fmax(r, sg[g][col]) : r + sg[g][col];
    if (is_max) raw_max[(int64_t)u * kNMax + (idx - raw_len)] = r;
    else raw_sums[(int64_t)u * raw_len + idx] = r;
  }
}

// ------------------------------------------------------------------------------------------------
// Stand-alone operator kernels
// ------------------------------------------------------------------------------------------------
// blockIdx.y = unit (hypothesis): every unit deskews the SAME raw points with its own twist.  xi_dev != NULL: twists
// read from device memory, (n_units, 6); outputs and partial sums are stacked per unit.
__global__ void __launch_bounds__(256) deskew_kernel(const double* __restrict__ pts, const double* __restrict__ t,
                                                     const double* __restrict__ w, int64_t n, double xi0, double xi1,
                                                     double xi2, double xi3, double xi4, double xi5,
                                                     const double* __restrict__ xi_dev, double t0, double t1,
                                                     double* __restrict__ o_pts, double* __restrict__ o_w,
                                                     double* __restrict__ partial, int n_vblocks, int vb_per_block) {
  __shared__ double sred[8];
  const int u = blockIdx.y;
  double xi[6] = {xi0, xi1, xi2, xi3, xi4, xi5};
  if (xi_dev) {
#pragma unroll
    for (int k = 0; k < 6; ++k) xi[k] = xi_dev[6 * u + k];
  }
  o_pts += (int64_t)u * 3 * n;
  o_w += (int64_t)u * n;
  partial += (int64_t)u * n_vblocks * 2;
  const double inv_denom = 1.0 / fmax(t1 - t0, 1e-12);
  // per-twist invariants hoisted (TwistCtx: ~60 instead of ~105 float64 operations per point, no sqrt / sincos for
  // |theta| < 0.5) and the window weight with one exponential and no division: the forms the fused bin kernels use
  const TwistCtx tw = make_twist_ctx(xi);
  const WindowCtx win = make_window_ctx(t0, t1);
  // The certificate sums are taken per VIRTUAL block of 256 threads striding over the points (n_vblocks of them, whatever
  // the grid): a CTA works through vb_per_block consecutive virtual blocks, so that a batch of many units can run few CTAs
  // per unit -- the per-thread invariants above cost five points' worth of arithmetic -- and still produce the partial
  // sums, and hence the certificates, of the single-unit launch bit for bit.
  for (int vb = blockIdx.x * vb_per_block; vb < n_vblocks && vb < (blockIdx.x + 1) * vb_per_block; ++vb) {
    double s_out = 0.0, s_in = 0.0;
    for (int64_t i = (int64_t)vb * blockDim.x + threadIdx.x; i < n; i += (int64_t)n_vblocks * blockDim.x) {
      double p[3] = {pts[3 * i], pts[3 * i + 1], pts[3 * i + 2]};
      double tt = t[i], ww = w[i];
      double p0[3];
      deskew_point_ctx(p, (tt - t0) * inv_denom, tw, p0);
      double wo = ww * window_weight_ctx(tt, win);
      o_pts[3 * i] = p0[0]; o_pts[3 * i + 1] = p0[1]; o_pts[3 * i + 2] = p0[2];
      o_w[i] = wo;
      s_out += wo; s_in += ww;
    }
    double a = block_sum_fixed(s_out, sred);
    double b = block_sum_fixed(s_in, sred);
    if (threadIdx.x == 0) { partial[2 * vb] = a; partial[2 * vb + 1] = b; }
  }
}
// blockIdx.x = unit: out[u * out_stride + k] = sum over the unit's parts.  One warp per column k (blockDim = 32 * width):
// lane l adds parts l, l + 32, ... in order, then the fixed shuffle tree -- 8 + 5 dependent additions for 256 parts instead of
// 256 (this kernel sat between the deskew and the surfel kernels of every scan with two busy threads)
__global__ void sum_pairs_kernel(const double* __restrict__ partial, int n_parts, int width, double* __restrict__ out,
                                 int out_stride) {
  const int k = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (k >= width) return;
  partial += (int64_t)blockIdx.x * n_parts * width;
  double a = 0.0;
  for (int c = lane; c < n_parts; c += 32) a += partial[(int64_t)c * width + k];
  a = warp_sum(a);
  if (lane == 0) out[(int64_t)blockIdx.x * out_stride + k] = a;
}
__global__ void __launch_bounds__(256) ray_dirs_kernel(const double* __restrict__ pts, int64_t n, double o0, double o1,
                                                       double o2, double eps, double* __restrict__ out) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double r0 = pts[3 * i] - o0, r1 = pts[3 * i + 1] - o1, r2 = pts[3 * i + 2] - o2;
  double inv = 1.0 / (sqrt(r0 * r0 + r1 * r1 + r2 * r2) + eps);
  out[3 * i] = r0 * inv; out[3 * i + 1] = r1 * inv; out[3 * i + 2] = r2 * inv;
}

// BinSoftAssign alone: half-warp per point, materialises responsibilities.  partial (blocks, 4): [ent_dot, ent_log, max, -]
template <int Q, int PREC>
__global__ void __launch_bounds__(256) soft_assign_kernel(const double* __restrict__ dirs, int64_t n,
                                                          const double* __restrict__ bin_dirs, int n_bins, double inv_tau,
                                                          double shift, int use_true_max, double* __restrict__ resp,
                                                          double* __restrict__ partial) {
  __shared__ double sred[8];
  const int tid = threadIdx.x, l16 = tid & 15, hw = tid >> 4;
  double bx[Q][3];
  bool bok[Q];
#pragma unroll
  for (int q = 0; q < Q; ++q) {
    int b = q * 16 + l16;
    bok[q] = b < n_bins;
    for (int k = 0; k < 3; ++k) bx[q][k] = bok[q] ? bin_dirs[3 * b + k] * (inv_tau * kLog2e) : 0.0;  // log2 units
  }
  double ent_dot = 0.0, ent_log = 0.0, mx = 0.0;
  const int64_t hw_global = (int64_t)blockIdx.x * (blockDim.x >> 4) + hw;
  const int64_t hw_total = (int64_t)gridDim.x * (blockDim.x >> 4);
  // all 16 lanes of a half-warp walk the same points; the tail is padded so shuffles stay convergent
  const int64_t n_iter = (n + hw_total - 1) / hw_total;
  for (int64_t itn = 0; itn < n_iter; ++itn) {
    const int64_t i = hw_global + itn * hw_total;
    const bool ok = i < n;
    double d0 = 0, d1 = 0, d2 = 0;
    if (ok) { d0 = dirs[3 * i]; d1 = dirs[3 * i + 1]; d2 = dirs[3 * i + 2]; }
    double x[Q];
#pragma unroll
    for (int q = 0; q < Q; ++q) x[q] = d0 * bx[q][0] + d1 * bx[q][1] + d2 * bx[q][2];
    double c = shift;
    if (use_true_max) {
      double m = -1.0e300;
#pragma unroll
      for (int q = 0; q < Q; ++q) m = bok[q] ? fmax(m, x[q]) : m;
      c = hw_max(m);
    }
    double e[Q], ssum = 0.0, dot = 0.0;
#pragma unroll
    for (int q = 0; q < Q; ++q) {
      double a = x[q] - c;
      if (PREC == 0) {
        e[q] = bok[q] ? exp2_nonpos(a) : 0.0;
      } else {
        float ef;
        asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(ef) : "f"((float)a));
        e[q] = bok[q] ? f32_bits_to_f64(ef) : 0.0;
      }
      ssum += e[q];
      dot = fma(e[q], a, dot);
    }
    ssum = hw_sum(ssum);
    dot = hw_sum(dot);
    const double inv = 1.0 / ssum;
    if (ok && l16 == 0) ent_dot = fma(inv * kLn2, dot, ent_dot);
    if (ok) {
      if (l16 == 0) ent_log += log(ssum);
#pragma unroll
      for (int q = 0; q < Q; ++q) {
        double r = e[q] * inv;
        mx = fmax(mx, r);
        if (bok[q]) resp[i * n_bins + q * 16 + l16] = r;
      }
    }
  }
  double a = block_sum_fixed(ent_dot, sred);
  double b = block_sum_fixed(ent_log, sred);
  double m = block_max_fixed(mx, sred);
  if (tid == 0) {
    double* o = partial + (int64_t)blockIdx.x * 4;
    o[0] = a; o[1] = b; o[2] = m; o[3] = 0.0;
  }
}
// The reference evaluates -sum_b r log(r + eps) (binning.py:72-74); we use the algebraic form log(sum e) - sum r (x - c),
// which equals -sum_b r log r.  The two differ by sum_b r log(1 + eps/r) = B*eps + O(eps^2 / r) per point: applied here.
__global__ void soft_assign_cert_kernel(const double* __restrict__ partial, int n_parts, double n_rows, int n_bins,
                                        double eps_mass, double* __restrict__ cert) {
  if (threadIdx.x != 0) return;
  double a = 0.0, b = 0.0, m = 0.0;
  for (int c = 0; c < n_parts; ++c) { a += partial[4 * c]; b += partial[4 * c + 1]; m = fmax(m, partial[4 * c + 2]); }
  cert[GCS_SA_ENTROPY_SUM] = (b - a) - n_rows * (double)n_bins * eps_mass;
  cert[GCS_SA_MAX_RESP] = m;
}

// ScanBinMomentMatch alone, responsibilities read from memory.  HAS_COV adds the 6 covariance features.
template <int Q, bool HAS_COV>
__global__ void __launch_bounds__(128, 2) moments_from_resp_kernel(
    const double* __restrict__ pts, const double* __restrict__ cov, const double* __restrict__ w,
    const double* __restrict__ resp, const double* __restrict__ lam, int64_t n, int n_bins, double o0, double o1,
    double o2, double eps_mass, double* __restrict__ partial, int part_len) {
  constexpr int NF = HAS_COV ? kNFCov : kNF;
  __shared__ __align__(16) double sbuf[128 * NF];
  const int tid = threadIdx.x, l16 = tid & 15, hw = tid >> 4;
  double acc[Q][NF];
#pragma unroll
  for (int q = 0; q < Q; ++q)
#pragma unroll
    for (int f = 0; f < NF; ++f) acc[q][f] = 0.0;
  const int64_t hw_global = (int64_t)blockIdx.x * (blockDim.x >> 4) + hw;
  const int64_t hw_total = (int64_t)gridDim.x * (blockDim.x >> 4);
  for (int64_t i = hw_global; i < n; i += hw_total) {
    double f[NF];
    const double p0 = pts[3 * i], p1 = pts[3 * i + 1], p2 = pts[3 * i + 2];
    const double r0 = p0 - o0, r1 = p1 - o1, r2 = p2 - o2;
    const double inv = 1.0 / (sqrt(r0 * r0 + r1 * r1 + r2 * r2) + eps_mass);
    const double d0 = r0 * inv, d1 = r1 * inv, d2 = r2 * inv;
    f[0] = 1.0; f[1] = d0; f[2] = d1; f[3] = d2;
    f[4] = d0 * d0; f[5] = d0 * d1; f[6] = d0 * d2; f[7] = d1 * d1; f[8] = d1 * d2; f[9] = d2 * d2;
    f[10] = p0; f[11] = p1; f[12] = p2;
    f[13] = p0 * p0; f[14] = p0 * p1; f[15] = p0 * p2; f[16] = p1 * p1; f[17] = p1 * p2; f[18] = p2 * p2;
    if (HAS_COV) {
      const double* c9 = cov + 9 * i;
      f[NF - 6] = c9[0]; f[NF - 5] = c9[1]; f[NF - 4] = c9[2]; f[NF - 3] = c9[4]; f[NF - 2] = c9[5]; f[NF - 1] = c9[8];
    }
    const double w_eff = w[i] * (lam ? lam[i] : 1.0);
#pragma unroll
    for (int q = 0; q < Q; ++q) {
      const int b = q * 16 + l16;
      const double wr = (b < n_bins) ? w_eff * resp[i * n_bins + b] : 0.0;
#pragma unroll
      for (int ff = 0; ff < NF; ++ff) acc[q][ff] = fma(wr, f[ff], acc[q][ff]);
    }
  }
  double* part = partial + (int64_t)blockIdx.x * part_len;
  reduce_acc_to_partial<Q, NF>(acc, sbuf, part, n_bins);
  if (tid < kNExtras + kNMax) part[n_bins * kRowLen + tid] = 0.0;
}

__global__ void kappa_batch_kernel(const double* __restrict__ Rbar, int64_t n, double eps_r, double d, double r0,
                                   double tau, double* __restrict__ out) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = kappa_from_resultant(Rbar[i], eps_r, d, r0, tau);
}


// Rank-ordered reduction of all-gathered partial statistics (point-sharded clouds): every rank holds the packed partial
// blocks of all ranks, (world, n_sum + n_max); sums add in rank order, maxima take the maximum -- the same arithmetic on
// the same data on every rank, so the reduced statistics are bit-identical everywhere whatever the collective library
// does inside.
__global__ void __launch_bounds__(256) reduce_gathered_kernel(const double* __restrict__ g, int world, int64_t n_sum, int64_t n_max,
                                                              double* __restrict__ out_sum, double* __restrict__ out_max) {
  const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
  const int64_t stride = n_sum + n_max;
  if (i >= stride) return;
  double a = g[i];
  if (i < n_sum) {
    for (int r = 1; r < world; ++r) a += g[(int64_t)r * stride + i];
    out_sum[i] = a;
  } else {
    for (int r = 1; r < world; ++r) a = fmax(a, g[(int64_t)r * stride + i]);
    out_max[i - n_sum] = a;
  }
}
}  // namespace gcs

// ================================================================================================
// Host side
// ================================================================================================
using namespace gcs;

static int pick_q(int n_bins) { return (n_bins + 15) / 16; }

struct SoftmaxShift { double shift; int use_true_max; };
// exp(x - shift) with a constant shift is exact algebra (softmax is shift invariant); it only needs the shifted
// logits to stay inside the float64 exponent range.  |x| <= bnorm/tau, so shift = bnorm/tau bounds the argument by
// -2*bnorm/tau.  Past ~-600 we switch to the per-point maximum (what jax.nn.softmax does).
static SoftmaxShift softmax_shift(double inv_tau, double bnorm_max) {
  SoftmaxShift s;
  s.shift = inv_tau * bnorm_max;
  s.use_true_max = (2.0 * s.shift > 600.0) ? 1 : 0;
  return s;
}

template <int Q, int PREC, int OCC>
static cudaError_t launch_scan_q(dim3 grid, cudaStream_t st, const BinScanParams& P) {
  const int smem = (int)sizeof(ScanSmem<Q>);
  {  // opt in to > 48 KB dynamic shared memory once per instantiation and device
    cudaError_t e = gcs_smem_attr_once((const void*)bin_scan_kernel<Q, PREC, OCC>, smem);
    if (e != cudaSuccess) return e;
  }
  bin_scan_kernel<Q, PREC, OCC><<<grid, kScanThreads, smem, st>>>(P);
  return cudaSuccess;
}
// resident CTAs per SM the kernel is compiled for (register budget): tunable for experiments via GCS_SCAN_OCC
static int scan_ctas_per_sm(int Q) {
  static int occ_env = -1;
  if (occ_env < 0) { const char* e = getenv("GCS_SCAN_OCC"); occ_env = e ? atoi(e) : 0; }
  if (Q > 3) return 2;
  return (occ_env == 2 || occ_env == 3) ? occ_env : 3;
}
template <int PREC>
static cudaError_t launch_scan(int Q, dim3 grid, cudaStream_t st, const BinScanParams& P) {
  const int occ = scan_ctas_per_sm(Q);
  switch (Q) {
    case 1: return occ == 3 ? launch_scan_q<1, PREC, 3>(grid, st, P) : launch_scan_q<1, PREC, 2>(grid, st, P);
    case 2: return occ == 3 ? launch_scan_q<2, PREC, 3>(grid, st, P) : launch_scan_q<2, PREC, 2>(grid, st, P);
    case 3: return occ == 3 ? launch_scan_q<3, PREC, 3>(grid, st, P) : launch_scan_q<3, PREC, 2>(grid, st, P);
    default: return launch_scan_q<4, PREC, 2>(grid, st, P);
  }
}

static int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }

struct BinsGeom {
  int64_t stride, n_sel, cap_total, n_raw_total;
  int U, ctas_per_unit, part_len, raw_len, mass_chunks;
  int tc_parts;   // partial slots per unit of the tensor-core kernel (0 when another precision was asked for)
  int64_t mass_rows_per_chunk;
};

static int bins_geometry(gcs_ctx* ctx, const gcs_bins_args* a, BinsGeom* g) {
  GCS_REQUIRE(ctx, a != nullptr, "gcs_bins: args is NULL");
  GCS_REQUIRE(ctx, a->n_scans >= 1 && a->n_hyp >= 1, "gcs_bins: n_scans=%d n_hyp=%d must be >= 1", a->n_scans, a->n_hyp);
  GCS_REQUIRE(ctx, a->n_raw >= 0 && a->cap >= 1, "gcs_bins: n_raw=%lld cap=%lld", (long long)a->n_raw, (long long)a->cap);
  GCS_REQUIRE(ctx, a->n_bins >= 1 && a->n_bins <= kMaxBins, "gcs_bins: n_bins=%d not in [1,%d]", a->n_bins, kMaxBins);
  GCS_REQUIRE(ctx, a->tau > 0.0, "gcs_bins: tau must be > 0 (got %g)", a->tau);
  GCS_REQUIRE(ctx, a->precision >= 0 && a->precision <= 2, "gcs_bins: precision=%d", a->precision);
  GCS_REQUIRE(ctx, a->pts && a->t && a->w && a->scan_t0 && a->scan_t1 && a->xi && a->bin_dirs && a->cert,
              "gcs_bins: a required device pointer is NULL");
  g->n_raw_total = a->n_raw_total > 0 ? a->n_raw_total : a->n_raw;
  g->cap_total = a->cap_total > 0 ? a->cap_total : a->cap;
  g->stride = g->n_raw_total <= g->cap_total ? 1 : ceil_div64(g->n_raw_total, g->cap_total);  // point_budget.py:160
  GCS_REQUIRE(ctx, a->shard_row0 % g->stride == 0, "gcs_bins: shard_row0=%lld must be a multiple of stride=%lld",
              (long long)a->shard_row0, (long long)g->stride);
  g->n_sel = ceil_div64(a->n_raw, g->stride);
  GCS_REQUIRE(ctx, g->n_sel <= a->cap, "gcs_bins: %lld selected rows do not fit cap=%lld", (long long)g->n_sel,
              (long long)a->cap);
  g->U = a->n_scans * a->n_hyp;
  int64_t n_tiles = ceil_div64(a->cap, kScanThreads);
  // CTAs per unit: fill whole waves of (SMs x resident CTAs) and keep >= 4 tiles per CTA when the work allows it,
  // so the per-CTA epilogue (3 smem reduction rounds + a 10 KB partial) amortises.
  {
    const int64_t slots = (int64_t)ctx->sm_count * scan_ctas_per_sm(pick_q(a->n_bins));
    int64_t cmax = n_tiles < 64 ? n_tiles : 64;
    if (cmax < 1) cmax = 1;
    int64_t best = 1;
    double best_score = -1.0;
    for (int64_t c = 1; c <= cmax; ++c) {
      const int64_t total = (int64_t)g->U * c;
      const double waves = (double)total / (double)slots;
      double eff = waves / (double)ceil_div64(total, slots);          // fraction of the last wave that is busy
      const double tiles_per_cta = (double)n_tiles / (double)c;
      const double imbalance = tiles_per_cta / (double)ceil_div64(n_tiles, c);  // uneven tile split inside a unit
      double score = eff * imbalance;
      if (tiles_per_cta < 4.0) score *= 0.25 + 0.1875 * tiles_per_cta;  // epilogue amortisation
      if (score > best_score + 1e-9) { best_score = score; best = c; }
    }
    g->ctas_per_unit = (int)best;
  }
  g->tc_parts = a->precision == GCS_PREC_TC ? bin_scan_tc_parts(ctx->sm_count, g->U, a->cap) : 0;
  g->raw_len = raw_sums_len(a->n_bins);
  g->part_len = g->raw_len + kNMax;
  // ~4 CTAs per SM over the whole batch, >= 1024 rows each, at most 64 per scan: fat chunks keep the loads in flight
  int64_t chunks = ceil_div64(4 * (int64_t)ctx->sm_count, a->n_scans);
  const int64_t max_by_rows = ceil_div64(a->n_raw > 0 ? a->n_raw : 1, 1024);
  if (chunks > max_by_rows) chunks = max_by_rows;
  if (chunks > 64) chunks = 64;
  if (chunks < 1) chunks = 1;
  g->mass_chunks = (int)chunks;
  g->mass_rows_per_chunk = ceil_div64(a->n_raw > 0 ? a->n_raw : 1, chunks);
  // keep stride alignment of chunk boundaries irrelevant: selection uses the absolute local row index
  return GCS_OK;
}

static int run_mass(gcs_ctx* ctx, cudaStream_t st, const gcs_bins_args* a, const BinsGeom& g, double* mass_out,
                    double* ws_partial) {
  dim3 grid(g.mass_chunks, a->n_scans);
  mass_partial_kernel<<<grid, 256, 0, st>>>(a->w, a->n_raw, g.stride, g.mass_rows_per_chunk, ws_partial);
  GCS_LAUNCH_CHECK(ctx);
  int tot = a->n_scans * kNMass;
  mass_final_kernel<<<(tot + 127) / 128, 128, 0, st>>>(ws_partial, g.mass_chunks, a->n_scans, mass_out);
  GCS_LAUNCH_CHECK(ctx);
  return GCS_OK;
}

// workspace layout for the fused path: [mass partial | mass | partials | raw_sums | raw_max]
struct BinsWs { double* mass_partial; double* mass; double* partial; double* raw_sums; double* raw_max; };
static int bins_ws(gcs_ctx* ctx, const gcs_bins_args* a, const BinsGeom& g, BinsWs* w) {
  uint64_t n_mp = (uint64_t)a->n_scans * g.mass_chunks * kNMass;
  uint64_t n_m = (uint64_t)a->n_scans * kNMass;
  const int parts = g.ctas_per_unit > g.tc_parts ? g.ctas_per_unit : g.tc_parts;
  uint64_t n_p = (uint64_t)g.U * parts * g.part_len;
  uint64_t n_r = (uint64_t)g.U * g.raw_len;
  uint64_t n_x = (uint64_t)g.U * kNMax;
  uint64_t total = (n_mp + n_m + n_p + n_r + n_x + 16) * sizeof(double);
  int rc = gcs_ws_reserve(ctx, total);
  if (rc != GCS_OK) return rc;
  double* p = (double*)ctx->ws;
  w->mass_partial = p; p += n_mp;
  w->mass = p; p += n_m;
  w->partial = p; p += n_p;
  w->raw_sums = p; p += n_r;
  w->raw_max = p;
  return GCS_OK;
}

int gcs_bins_finalize_impl(gcs_ctx* ctx, cudaStream_t st, const gcs_bins_args* a, int64_t cap_total,
                           const double* mass, const double* raw_sums, const double* raw_max);  // gcs_bins_final.cu

static int accumulate_impl(gcs_ctx* ctx, cudaStream_t st, const gcs_bins_args* a, const BinsGeom& g, const BinsWs& w,
                           const double* mass, double* raw_sums, double* raw_max) {
  BinScanParams P;
  P.pts = a->pts; P.t = a->t; P.w = a->w; P.ring = a->ring; P.tag = a->tag;
  P.n_raw = a->n_raw; P.cap = a->cap; P.n_sel = g.n_sel; P.stride = g.stride;
  P.n_scans = a->n_scans; P.n_hyp = a->n_hyp; P.n_bins = a->n_bins;
  P.t0s = a->scan_t0; P.t1s = a->scan_t1; P.xi = a->xi; P.bin_dirs = a->bin_dirs;
  P.origin[0] = a->origin[0]; P.origin[1] = a->origin[1]; P.origin[2] = a->origin[2];
  P.inv_tau = 1.0 / a->tau;
  SoftmaxShift sh = softmax_shift(P.inv_tau, a->bin_norm_max > 0.0 ? a->bin_norm_max : 1.0);
  P.shift = sh.shift; P.use_true_max = sh.use_true_max;
  P.eps_mass = a->eps_mass;
  P.mass = mass;
  P.rs_pts = a->rs_pts; P.rs_t = a->rs_t; P.rs_w = a->rs_w; P.rs_ring = a->rs_ring; P.rs_tag = a->rs_tag;
  if (P.rs_pts) GCS_REQUIRE(ctx, P.rs_t && P.rs_w && P.rs_ring && P.rs_tag, "gcs_bins: rs_* outputs must be all set or all NULL");
  P.dk_pts = a->dk_pts; P.dk_w = a->dk_w; P.resp = a->resp;
  P.partial = w.partial; P.part_len = g.part_len;
  dim3 grid(g.ctas_per_unit, g.U);
  const int Q = pick_q(a->n_bins);
  // GCS_PREC_TC needs the constant softmax shift and no materialised responsibilities; otherwise it degrades to MIXED
  const bool use_tc = a->precision == GCS_PREC_TC && bin_scan_tc_supported(P);
  int n_parts = g.ctas_per_unit;
  if (use_tc) {
    n_parts = g.tc_parts;
    GCS_CHECK_CUDA(ctx, cudaMemsetAsync(P.partial, 0, (size_t)g.U * n_parts * g.part_len * sizeof(double), st));
  }
  gcs_timing_begin(ctx, st, GCS_TIME_BIN_SCAN);
  if (use_tc) GCS_CHECK_CUDA(ctx, launch_bin_scan_tc(ctx->sm_count, st, P, n_parts));
  else if (a->precision == GCS_PREC_F64) GCS_CHECK_CUDA(ctx, launch_scan<0>(Q, grid, st, P));
  else GCS_CHECK_CUDA(ctx, launch_scan<1>(Q, grid, st, P));
  gcs_timing_end(ctx, st, GCS_TIME_BIN_SCAN);
  GCS_LAUNCH_CHECK(ctx);
  reduce_partials_kernel<<<dim3((g.raw_len + kNMax + 63) / 64, g.U), 256, 0, st>>>(w.partial, n_parts, g.part_len,
                                                                                    a->n_bins, kNF, raw_sums, raw_max);
  GCS_LAUNCH_CHECK(ctx);
  return GCS_OK;
}

extern "C" {

int gcs_bins_raw_sums_len(int n_bins) { return raw_sums_len(n_bins); }

int gcs_bins_mass(gcs_ctx* ctx, void* stream, const gcs_bins_args* a, double* mass) {
  if (!ctx) return GCS_EINVAL;
  GCS_CHECK_CUDA(ctx, cudaSetDevice(ctx->device));
  BinsGeom g;
  int rc = bins_geometry(ctx, a, &g);
  if (rc) return rc;
  GCS_REQUIRE(ctx, mass != nullptr, "gcs_bins_mass: mass is NULL");
  BinsWs w;
  rc = bins_ws(ctx, a, g, &w);
  if (rc) return rc;
  return run_mass(ctx, (cudaStream_t)stream, a, g, mass, w.mass_partial);
}

int gcs_bins_accumulate(gcs_ctx* ctx, void* stream, const gcs_bins_args* a, const double* mass, double* raw_sums,
                        double* raw_max) {
  if (!ctx) return GCS_EINVAL;
  GCS_CHECK_CUDA(ctx, cudaSetDevice(ctx->device));
  BinsGeom g;
  int rc = bins_geometry(ctx, a, &g);
  if (rc) return rc;
  GCS_REQUIRE(ctx, raw_sums && raw_max, "gcs_bins_accumulate: raw_sums/raw_max is NULL");
  BinsWs w;
  rc = bins_ws(ctx, a, g, &w);
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  if (!mass) {
    rc = run_mass(ctx, st, a, g, w.mass, w.mass_partial);
    if (rc) return rc;
    mass = w.mass;
  }
  return accumulate_impl(ctx, st, a, g, w, mass, raw_sums, raw_max);
}

int gcs_bins_finalize(gcs_ctx* ctx, void* stream, const gcs_bins_args* a, const double* mass, const double* raw_sums,
                      const double* raw_max) {
  if (!ctx) return GCS_EINVAL;
  GCS_CHECK_CUDA(ctx, cudaSetDevice(ctx->device));
  BinsGeom g;
  int rc = bins_geometry(ctx, a, &g);
  if (rc) return rc;
  GCS_REQUIRE(ctx, mass && raw_sums && raw_max, "gcs_bins_finalize: mass/raw_sums/raw_max is NULL");
  return gcs_bins_finalize_impl(ctx, (cudaStream_t)stream, a, g.cap_total, mass, raw_sums, raw_max);
}

int gcs_lidar_evidence_bins(gcs_ctx* ctx, void* stream, const gcs_bins_args* a) {
  if (!ctx) return GCS_EINVAL;
  GCS_CHECK_CUDA(ctx, cudaSetDevice(ctx->device));
  BinsGeom g;
  int rc = bins_geometry(ctx, a, &g);
  if (rc) return rc;
  BinsWs w;
  rc = bins_ws(ctx, a, g, &w);
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  rc = run_mass(ctx, st, a, g, w.mass, w.mass_partial);
  if (rc) return rc;
  rc = accumulate_impl(ctx, st, a, g, w, w.mass, w.raw_sums, w.raw_max);
  if (rc) return rc;
  return gcs_bins_finalize_impl(ctx, st, a, g.cap_total, w.mass, w.raw_sums, w.raw_max);
}

// ---- a1 ---------------------------------------------------------------------------------------------
int gcs_point_budget_resample(gcs_ctx* ctx, void* stream, const double* pts, const double* t, const double* w,
                              const uint8_t* ring, const uint8_t* tag, int64_t n_raw, int64_t cap, double eps_mass,
                              double* out_pts, double* out_t, double* out_w, uint8_t* out_ring, uint8_t* out_tag,
                              double* cert) {
  if (!ctx) return GCS_EINVAL;
  GCS_CHECK_CUDA(ctx, cudaSetDevice(ctx->device));
  GCS_REQUIRE(ctx, n_raw >= 0 && cap >= 1, "point_budget_resample: n_raw=%lld cap=%lld", (long long)n_raw, (long long)cap);
  GCS_REQUIRE(ctx, (n_raw == 0 || (pts && t && w)) && out_pts && out_t && out_w && out_ring && out_tag && cert,
              "point_budget_resample: NULL pointer");
  cudaStream_t st = (cudaStream_t)stream;
  int64_t stride = n_raw <= cap ? 1 : ceil_div64(n_raw, cap);
  int64_t n_sel = ceil_div64(n_raw, stride);
  int64_t chunks = ceil_div64(n_raw > 0 ? n_raw : 1, 2048);
  if (chunks > 64) chunks = 64;
  int64_t rpc = ceil_div64(n_raw > 0 ? n_raw : 1, chunks);
  int rc = gcs_ws_reserve(ctx, (uint64_t)(chunks + 1) * kNMass * sizeof(double));
  if (rc) return rc;
  double* part = (double*)ctx->ws;
  double* mass = part + chunks * kNMass;
  mass_partial_kernel<<<dim3((unsigned)chunks, 1), 256, 0, st>>>(w, n_raw, stride, rpc, part);
  GCS_LAUNCH_CHECK(ctx);
  mass_final_kernel<<<1, 128, 0, st>>>(part, (int)chunks, 1, mass);
  GCS_LAUNCH_CHECK(ctx);
  resample_cert_kernel<<<1, 32, 0, st>>>(mass, cap, eps_mass, cert);
  GCS_LAUNCH_CHECK(ctx);
  resample_gather_kernel<<<(unsigned)ceil_div64(cap, 256), 256, 0, st>>>(pts, t, w, ring, tag, n_sel, cap, stride, mass,
                                                                          eps_mass, out_pts, out_t, out_w, out_ring, out_tag);
  GCS_LAUNCH_CHECK(ctx);
  return GCS_OK;
}

// ---- a2 ---------------------------------------------------------------------------------------------
static int deskew_launch(gcs_ctx* ctx, cudaStream_t st, const double* pts, const double* t, const double* w, int64_t n,
                         const double* xi_host, const double* xi_dev, int n_units, double t0, double t1, double* out_pts,
                         double* out_w, double* cert) {
  int blocks = (int)ceil_div64(n > 0 ? n : 1, 256);
  int maxb = ctx->sm_count * 8;
  if (blocks > maxb) blocks = maxb;
  int rc = gcs_ws_reserve(ctx, (uint64_t)n_units * blocks * 2 * sizeof(double));
  if (rc) return rc;
  double* part = (double*)ctx->ws;
  const double z6[6] = {0, 0, 0, 0, 0, 0};
  const double* xi = xi_host ? xi_host : z6;
  // virtual blocks per CTA: one for a single unit (latency), more when the batch fills the device anyway
  int vbp = 1;
  if (n_units > 1) {
    const int64_t want_ctas = (int64_t)ctx->sm_count * 8;
    vbp = (int)(((int64_t)blocks * n_units + want_ctas - 1) / want_ctas);
    if (vbp < 1) vbp = 1;
    if (vbp > 16) vbp = 16;
  }
  const int ctas = (blocks + vbp - 1) / vbp;
  deskew_kernel<<<dim3(ctas, n_units), 256, 0, st>>>(pts, t, w, n, xi[0], xi[1], xi[2], xi[3], xi[4], xi[5], xi_dev, t0, t1,
                                                     out_pts, out_w, part, blocks, vbp);
  GCS_LAUNCH_CHECK(ctx);
  sum_pairs_kernel<<<n_units, 64, 0, st>>>(part, blocks, 2, cert, GCS_DK_NCERT);
  GCS_LAUNCH_CHECK(ctx);
  return GCS_OK;
}

int gcs_deskew_constant_twist(gcs_ctx* ctx, void* stream, const double* pts, const double* t, const double* w, int64_t n,
                              const double* xi, double t0, double t1, double* out_pts, double* out_w, double* cert) {
  if (!ctx) return GCS_EINVAL;
  GCS_CHECK_CUDA(ctx, cudaSetDevice(ctx->device));
  GCS_REQUIRE(ctx, n >= 0 && xi && cert && (n == 0 || (pts && t && w && out_pts && out_w)), "deskew_constant_twist: bad args");
  return deskew_launch(ctx, (cudaStream_t)stream, pts, t, w, n, xi, nullptr, 1, t0, t1, out_pts, out_w, cert);
}

int gcs_deskew_constant_twist_batched(gcs_ctx* ctx, void* stream, const double* pts, const double* t, const double* w,
                                      int64_t n, const double* xi_dev, int32_t n_units, double t0, double t1,
                                      double* out_pts, double* out_w, double* cert) {
  if (!ctx) return GCS_EINVAL;
  GCS_CHECK_CUDA(ctx, cudaSetDevice(ctx->device));
  GCS_REQUIRE(ctx, n >= 1 && n_units >= 1 && n_units <= 65535 && xi_dev && cert && pts && t && w && out_pts && out_w,
              "deskew_constant_twist_batched: bad args");
  return deskew_launch(ctx, (cudaStream_t)stream, pts, t, w, n, nullptr, xi_dev, n_units, t0, t1, out_pts, out_w, cert);
}

// ---- a3 ---------------------------------------------------------------------------------------------
int gcs_ray_directions(gcs_ctx* ctx, void* stream, const double* pts, int64_t n, const double* origin, double eps,
                       double* out) {
  if (!ctx) return GCS_EINVAL;
  GCS_CHECK_CUDA(ctx, cudaSetDevice(ctx->device));
  GCS_REQUIRE(ctx, n >= 0 && origin && (n == 0 || (pts && out)), "ray_directions: bad args");
  if (n == 0) return GCS_OK;
  ray_dirs_kernel<<<(unsigned)ceil_div64(n, 256), 256, 0, (cudaStream_t)stream>>>(pts, n, origin[0], origin[1], origin[2],
                                                                                   eps, out);
  GCS_LAUNCH_CHECK(ctx);
  return GCS_OK;
}

// ---- a4 ---------------------------------------------------------------------------------------------
int gcs_bin_soft_assign(gcs_ctx* ctx, void* stream, const double* dirs, int64_t n, const double* bin_dirs, int n_bins,
                        double tau, double eps_mass, int precision, double* out_resp, double* cert) {
  if (!ctx) return GCS_EINVAL;
  GCS_CHECK_CUDA(ctx, cudaSetDevice(ctx->device));
  GCS_REQUIRE(ctx, n >= 1 && dirs && bin_dirs && out_resp && cert, "bin_soft_assign: bad args");
  GCS_REQUIRE(ctx, n_bins >= 1 && n_bins <= kMaxBins, "bin_soft_assign: n_bins=%d not in [1,%d]", n_bins, kMaxBins);
  GCS_REQUIRE(ctx, tau > 0.0, "bin_soft_assign: tau must be > 0");
  cudaStream_t st = (cudaStream_t)stream;
  int blocks = (int)ceil_div64(n, 16 * 8);
  int maxb = ctx->sm_count * 8;
  if (blocks > maxb) blocks = maxb;
  int rc = gcs_ws_reserve(ctx, (uint64_t)blocks * 4 * sizeof(double));
  if (rc) return rc;
  double* part = (double*)ctx->ws;
  // stand-alone operator takes arbitrary bin vectors: always use the per-point maximum (exactly jax.nn.softmax)
  const double inv_tau = 1.0 / tau;
  const int Q = pick_q(n_bins);
#define GCS_SA(QQ, PP) soft_assign_kernel<QQ, PP><<<blocks, 256, 0, st>>>(dirs, n, bin_dirs, n_bins, inv_tau, 0.0, 1, out_resp, part)
  if (precision == GCS_PREC_F64) {
    if (Q == 1) GCS_SA(1, 0); else if (Q == 2) GCS_SA(2, 0); else if (Q == 3) GCS_SA(3, 0); else GCS_SA(4, 0);
  } else {
    if (Q == 1) GCS_SA(1, 1); else if (Q == 2) GCS_SA(2, 1); else if (Q == 3) GCS_SA(3, 1); else GCS_SA(4, 1);
  }
#undef GCS_SA
  GCS_LAUNCH_CHECK(ctx);
  soft_assign_cert_kernel<<<1, 32, 0, st>>>(part, blocks, (double)n, n_bins, eps_mass, cert);
  GCS_LAUNCH_CHECK(ctx);
  return GCS_OK;
}

// ---- a6 ---------------------------------------------------------------------------------------------
int gcs_kappa_from_resultant_batch(gcs_ctx* ctx, void* stream, const double* R_bar, int64_t n, double eps_r, double d,
                                   double r0, double tau, double* out) {
  if (!ctx) return GCS_EINVAL;
  GCS_CHECK_CUDA(ctx, cudaSetDevice(ctx->device));
  GCS_REQUIRE(ctx, n >= 0 && (n == 0 || (R_bar && out)), "kappa_from_resultant_batch: bad args");
  if (n == 0) return GCS_OK;
  kappa_batch_kernel<<<(unsigned)ceil_div64(n, 128), 128, 0, (cudaStream_t)stream>>>(R_bar, n, eps_r, d, r0, tau, out);
  GCS_LAUNCH_CHECK(ctx);
  return GCS_OK;
}


int gcs_bins_reduce_gathered(gcs_ctx* ctx, void* stream, const double* gathered, int32_t world, int64_t n_sum, int64_t n_max,
                             double* out_sum, double* out_max) {
  if (!ctx) return GCS_EINVAL;
  GCS_CHECK_CUDA(ctx, cudaSetDevice(ctx->device));
  GCS_REQUIRE(ctx, gathered && world >= 1 && n_sum >= 0 && n_max >= 0 && n_sum + n_max >= 1 && (n_sum == 0 || out_sum) &&
                       (n_max == 0 || out_max), "bins_reduce_gathered: bad args");
  reduce_gathered_kernel<<<(unsigned)ceil_div64(n_sum + n_max, 256), 256, 0, (cudaStream_t)stream>>>(gathered, world, n_sum, n_max,
                                                                                                    out_sum, out_max);
  GCS_LAUNCH_CHECK(ctx);
  return GCS_OK;
}

}  // extern "C"

// ---- a5 (host part lives here because it launches moments_from_resp_kernel) ---------------------------
int gcs_stats_from_raw_impl(gcs_ctx* ctx, cudaStream_t st, const double* raw_sums, int n_units, int n_bins, double eps_psd,
                            double eps_mass, const gcs_bin_stats* out, double* cert_st);  // gcs_bins_final.cu

extern "C" int gcs_scan_bin_moment_match(gcs_ctx* ctx, void* stream, const double* pts, const double* point_cov,
                                         const double* w, const double* resp, const double* point_lambda,
                                         const double* origin, int64_t n, int n_bins, double eps_psd, double eps_mass,
                                         const gcs_bin_stats* out, double* cert) {
  if (!ctx) return GCS_EINVAL;
  GCS_CHECK_CUDA(ctx, cudaSetDevice(ctx->device));
  GCS_REQUIRE(ctx, n >= 1 && pts && w && resp && origin && out && cert, "scan_bin_moment_match: bad args");
  GCS_REQUIRE(ctx, n_bins >= 1 && n_bins <= kMaxBins, "scan_bin_moment_match: n_bins=%d not in [1,%d]", n_bins, kMaxBins);
  cudaStream_t st = (cudaStream_t)stream;
  int blocks = (int)ceil_div64(n, 8 * 16);
  int maxb = ctx->sm_count * 4;
  if (blocks > maxb) blocks = maxb;
  const int raw_len = raw_sums_len(n_bins), part_len = raw_len + kNMax;
  uint64_t need = ((uint64_t)blocks * part_len + raw_len + kNMax + 8) * sizeof(double);
  int rc = gcs_ws_reserve(ctx, need);
  if (rc) return rc;
  double* part = (double*)ctx->ws;
  double* raw = part + (uint64_t)blocks * part_len;
  double* rmax = raw + raw_len;
  const int Q = pick_q(n_bins);
#define GCS_MM(QQ, CC) moments_from_resp_kernel<QQ, CC><<<blocks, 128, 0, st>>>(pts, point_cov, w, resp, point_lambda, n, n_bins, origin[0], origin[1], origin[2], eps_mass, part, part_len)
  if (point_cov) {
    if (Q == 1) GCS_MM(1, true); else if (Q == 2) GCS_MM(2, true); else if (Q == 3) GCS_MM(3, true); else GCS_MM(4, true);
  } else {
    if (Q == 1) GCS_MM(1, false); else if (Q == 2) GCS_MM(2, false); else if (Q == 3) GCS_MM(3, false); else GCS_MM(4, false);
  }
#undef GCS_MM
  GCS_LAUNCH_CHECK(ctx);
  reduce_partials_kernel<<<dim3((raw_len + kNMax + 63) / 64, 1), 256, 0, st>>>(part, blocks, part_len, n_bins,
                                                                                point_cov ? kNFCov : kNF, raw, rmax);
  GCS_LAUNCH_CHECK(ctx);
  return gcs_stats_from_raw_impl(ctx, st, raw, 1, n_bins, eps_psd, eps_mass, out, cert);
}
