// gcs_bins_tc.cu -- tensor-core variant of the fused bin kernel (GCS_PREC_TC).
//
// The per-bin moment accumulation  M[b, f] = sum_i  (w_i / Z_i) e_ib  *  phi_f(i)   (48 bins x 19 features, all points i
// of a scan) is a [bins x points] . [points x features] contraction.  Here it runs on the 5th-generation tensor cores:
//
//   producer warps (8 per CTA)  one thread per point: gather + constant-twist deskew + window weight + ray direction
//            in float64 (the deskewed cloud is an operator output and stays float64-exact); logits, MUFU ex2 in
//            float32.  Every e_ib and every scaled feature (w_i/Z_i) phi_f(i) is split into two tf32 terms
//            x = hi + lo (22 significant bits) and stored K-major (K = point) into the warp's shared-memory operand
//            tile in the 128-byte-swizzled layout of gcs_tc.cuh -- a warp writes one value per lane into one row: no
//            bank conflicts.  A = [e_hi ; e_lo] (2 x 16Q rows), B = [phi_hi ; phi_lo] (38 rows).
//   MMA      lane 0 of the producer warp issues 4 x tcgen05.mma.kind::tf32 (M=128, N=48, K=8) per 32-point tile:
//            D[row, col] += sum_k A[row, k] B[col, k] in the warp's own TMEM accumulator; all four hi/lo cross
//            products land in one instruction.  Padding rows/columns of the 128 x 48 tile read neighbouring shared
//            memory; they only pollute accumulator rows/columns nobody reads.
//   epilogue warps (Q per CTA, one per 32 TMEM lanes) drain every accumulator after `flush` tiles (tcgen05.ld) into
//            float64 registers: the float32 accumulator (which truncates) never sums more than flush x 4 MMA steps.
//
// A CTA is persistent over a contiguous range of 32-point tiles of the flattened (unit, tile) space -- every SM gets
// the same amount of work whatever the batch shape -- and writes one partial per unit segment it touched, in the
// layout of bin_scan_kernel, so reduce_partials_kernel and the finalize kernel are shared with the other precisions.
// Accumulation order is fixed (per-warp tile order, epilogue drains warps in index order): bit-identical reruns.
#include <stdlib.h>

#include "gcs_bins.cuh"
#include "gcs_tc.cuh"

namespace gcs {

namespace {

constexpr int kProd = 8;            // producer warps per CTA
constexpr int kAccStride = 64;      // TMEM columns between per-warp accumulators (8 x 64 = 512 = all of TMEM)
constexpr int kMmaN = 48;           // >= 2 * kNF, multiple of 16
constexpr int kBRows = 40;          // B tile rows kept in shared memory (38 used)
constexpr double kLog2e = 1.4426950408889634;
constexpr double kLn2 = 0.6931471805599453;

template <int Q>
struct TcCfg {
  static constexpr int kBinsPad = 16 * Q;   // rows of the e_hi block == first row of the e_lo block
  static constexpr int kRows = 32 * Q;      // rows of A that carry data
  static constexpr int kEpi = Q;            // epilogue warps (32 TMEM lanes each)
  static constexpr int kThreads = 32 * (kProd + kEpi);
  static constexpr int kABytes = kRows * tc::kRowBytes;
  static constexpr int kBBytes = kBRows * tc::kRowBytes;
  static constexpr int kStageBytes = kABytes + kBBytes;            // 17 KB (Q=3) / 21 KB (Q=4): multiples of 1 KB
  static constexpr int kStagesBytes = kProd * kStageBytes + 1024;  // + tail read by the last B tile's padding rows
};

struct TcMisc {
  float4 bins[kMaxBins];
  uint64_t bar_stage[kProd];   // MMAs that read the warp's operand tile have completed
  uint64_t bar_full[kProd];    // the warp's accumulator holds a finished round
  uint64_t bar_empty[kProd];   // the epilogue has drained it
  uint32_t tmem;
  double ex[kProd][8];
};

struct TcGeom {
  int64_t tiles_per_unit;   // ceil(cap / 32)
  int64_t total_tiles;      // U * tiles_per_unit
  int n_cta, n_parts, flush;
};

__device__ __forceinline__ int64_t cta_tile0(const TcGeom& G, int c) { return (int64_t)c * G.total_tiles / G.n_cta; }
// CTA whose range contains tile g
__device__ __forceinline__ int cta_of_tile(const TcGeom& G, int64_t g) {
  return (int)(((g + 1) * G.n_cta - 1) / G.total_tiles);
}

template <int Q>
__global__ void __launch_bounds__(TcCfg<Q>::kThreads, 1) bin_scan_tc_kernel(const BinScanParams P, const TcGeom G) {
  using C = TcCfg<Q>;
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  // 1 KB alignment by pointer arithmetic on the __shared__ symbol (keeps the shared address space: STS/LDS, not generic)
  unsigned char* stages = smem_raw + ((1024u - (tc::smem_u32(smem_raw) & 1023u)) & 1023u);
  TcMisc& mi = *reinterpret_cast<TcMisc*>(stages + C::kStagesBytes);
  const int tid = threadIdx.x, wid = tid >> 5, lane = tid & 31;
  const int nb = P.n_bins;

  // ---- one-time setup
  for (int k = tid; k < C::kStagesBytes / 16; k += C::kThreads) reinterpret_cast<uint4*>(stages)[k] = make_uint4(0, 0, 0, 0);
  {
    const double sc = P.inv_tau * kLog2e;
    const float c2 = (float)(P.shift * kLog2e);
    for (int b = tid; b < kMaxBins; b += C::kThreads) {
      float4 v = make_float4(0.f, 0.f, 0.f, -1.0e30f);   // bins past n_bins: e = 2^(-1e30) = 0
      if (b < nb) v = make_float4((float)(P.bin_dirs[3 * b] * sc), (float)(P.bin_dirs[3 * b + 1] * sc),
                                  (float)(P.bin_dirs[3 * b + 2] * sc), -c2);
      mi.bins[b] = v;
    }
  }
  if (tid == 0) {
    for (int w = 0; w < kProd; ++w) {
      tc::mbar_init(&mi.bar_stage[w], 1);
      tc::mbar_init(&mi.bar_full[w], 1);
      tc::mbar_init(&mi.bar_empty[w], C::kEpi);
    }
    tc::mbar_init_fence();
  }
  if (wid == 0) tc::tmem_alloc(&mi.tmem, 512);
  tc::fence_smem_to_async();
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem = mi.tmem;

  // ---- persistent range of this CTA
  const int cta = blockIdx.x;
  const int64_t g_begin = cta_tile0(G, cta), g_end = cta_tile0(G, cta + 1);

  // per-thread state that survives across unit segments
  uint32_t n_stage_uses = 0, n_rounds = 0;        // producer: uses of the operand tile / finished accumulator rounds
  uint32_t n_drained[kProd];                      // epilogue: rounds drained per producer warp
#pragma unroll
  for (int w = 0; w < kProd; ++w) n_drained[w] = 0;
  // lane-dependent byte offsets of this lane's element inside a row, for the 8 row phases of the swizzle
  uint32_t off[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) off[j] = (uint32_t)((((lane >> 2) ^ j) << 4) | ((lane & 3) << 2));

  for (int64_t g0 = g_begin; g0 < g_end;) {
    const int u = (int)(g0 / G.tiles_per_unit);
    const int64_t unit_t0 = (int64_t)u * G.tiles_per_unit;
    const int64_t g1 = (unit_t0 + G.tiles_per_unit < g_end) ? unit_t0 + G.tiles_per_unit : g_end;
    const int64_t lt0 = g0 - unit_t0, lt1 = g1 - unit_t0;   // local tile range inside unit u
    const int s = u / P.n_hyp, h = u - s * P.n_hyp;

    if (wid < kProd) {
      // =============================== producer warp ===============================
      unsigned char* sA = stages + wid * C::kStageBytes;
      unsigned char* sB = sA + C::kABytes;
      const uint32_t aA = tc::smem_u32(sA), aB = tc::smem_u32(sB);
      const double t0 = P.t0s[s], t1 = P.t1s[s];
      const double inv_denom = 1.0 / fmax(t1 - t0, 1e-12);
      const double inv_sig = window_inv_sigma(t0, t1);
      double xi[6];
#pragma unroll
      for (int k = 0; k < 6; ++k) xi[k] = P.xi[(int64_t)u * 6 + k];
      const double mass_scale = P.mass[s * kNMass + kMassAll] / (P.mass[s * kNMass + kMassSel] + P.eps_mass);
      const double* pts = P.pts + (int64_t)s * P.n_raw * 3;
      const double* tp = P.t + (int64_t)s * P.n_raw;
      const double* wp = P.w + (int64_t)s * P.n_raw;
      const uint8_t* rp = P.ring ? P.ring + (int64_t)s * P.n_raw : nullptr;
      const uint8_t* gp = P.tag ? P.tag + (int64_t)s * P.n_raw : nullptr;
      double ent_dot = 0.0, ent_log = 0.0, mx_resp = 0.0, sum_wdk = 0.0, sum_wrs = 0.0, n_rows = 0.0;
      uint32_t in_round = 0;

      // software pipeline: the raw rows of the warp's next tile are requested before this tile is processed
      double nx[5] = {0.0, 0.0, 0.0, 0.0, 0.0};
      uint8_t nrg = 0, ntg = 0;
      auto fetch = [&](int64_t tile) {
        const int64_t ii = tile * tc::kTileK + lane;
        nx[0] = nx[1] = nx[2] = nx[3] = nx[4] = 0.0; nrg = 0; ntg = 0;
        if (tile < lt1 && ii < P.n_sel) {
          const int64_t j = ii * P.stride;
          nx[0] = pts[3 * j]; nx[1] = pts[3 * j + 1]; nx[2] = pts[3 * j + 2];
          nx[3] = tp[j]; nx[4] = wp[j];
          if (rp) nrg = rp[j];
          if (gp) ntg = gp[j];
        }
      };
      fetch(lt0 + wid);
      for (int64_t lt = lt0 + wid; lt < lt1; lt += kProd) {
        const int64_t i = lt * tc::kTileK + lane;
        const bool row = i < P.cap;
        const double p[3] = {nx[0], nx[1], nx[2]}, tt = nx[3], ww = nx[4];
        const uint8_t rg = nrg, tg = ntg;
        fetch(lt + kProd);
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
        const double r0 = p0[0] - P.origin[0], r1 = p0[1] - P.origin[1], r2 = p0[2] - P.origin[2];
        const double nrm = sqrt(r0 * r0 + r1 * r1 + r2 * r2);
        const double invn = 1.0 / (nrm + P.eps_mass);
        const double d0 = r0 * invn, d1 = r1 * invn, d2 = r2 * invn;
        const float f0 = (float)d0, f1 = (float)d1, f2 = (float)d2;

        // the tensor core may still be reading this warp's previous tile
        if (n_stage_uses > 0) tc::mbar_wait(&mi.bar_stage[wid], (n_stage_uses - 1) & 1);

        // ---- A = [e_hi ; e_lo]: softmax numerators in log2 units, unnormalised (1/Z goes into B)
        float sum = 0.f, dot = 0.f, emax = 0.f;
#pragma unroll
        for (int b = 0; b < C::kBinsPad; ++b) {
          const float4 bb = mi.bins[b];
          const float l = fmaf(f0, bb.x, fmaf(f1, bb.y, fmaf(f2, bb.z, bb.w)));
          float e;
          asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(l));
          sum += e;
          dot = fmaf(e, l, dot);
          emax = fmaxf(emax, e);
          float hi, lo;
          tc::split_tf32(e, hi, lo);
          *reinterpret_cast<float*>(sA + b * tc::kRowBytes + off[b & 7]) = hi;
          *reinterpret_cast<float*>(sA + (C::kBinsPad + b) * tc::kRowBytes + off[b & 7]) = lo;
        }
        const double ssum = (double)sum;
        const double inv = 1.0 / ssum;
        if (row) {
          ent_dot = fma(inv * kLn2, (double)dot, ent_dot);
          ent_log += log(ssum);
          mx_resp = fmax(mx_resp, (double)emax * inv);
        }
        // ---- B = [phi_hi ; phi_lo], phi = (w / Z) * (1, d, d d^T, p, p p^T)
        {
          const double sc = row ? w_dk * inv : 0.0;
          const double sd0 = sc * d0, sd1 = sc * d1, sd2 = sc * d2;
          const double sp0 = sc * p0[0], sp1 = sc * p0[1], sp2 = sc * p0[2];
          const double v[kNF] = {sc, sd0, sd1, sd2, sd0 * d0, sd0 * d1, sd0 * d2, sd1 * d1, sd1 * d2, sd2 * d2,
                                 sp0, sp1, sp2, sp0 * p0[0], sp0 * p0[1], sp0 * p0[2], sp1 * p0[1], sp1 * p0[2], sp2 * p0[2]};
#pragma unroll
          for (int f = 0; f < kNF; ++f) {
            float hi, lo;
            tc::split_tf32((float)v[f], hi, lo);
            *reinterpret_cast<float*>(sB + f * tc::kRowBytes + off[f & 7]) = hi;
            *reinterpret_cast<float*>(sB + (kNF + f) * tc::kRowBytes + off[(kNF + f) & 7]) = lo;
          }
        }
        tc::fence_smem_to_async();
        __syncwarp();
        const bool last = (in_round + 1 == (uint32_t)G.flush) || (lt + kProd >= lt1);
        if (lane == 0) {
          if (in_round == 0 && n_rounds > 0) tc::mbar_wait(&mi.bar_empty[wid], (n_rounds - 1) & 1);
          tc::fence_after_sync();
          const uint64_t da = tc::smem_desc_sw128(aA), db = tc::smem_desc_sw128(aB);
          const uint32_t idesc = tc::idesc_tf32(128, kMmaN);
          const uint32_t d_tmem = tmem + wid * kAccStride;
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) tc::mma_tf32_ss(d_tmem, da + 2 * ks, db + 2 * ks, idesc, (in_round | ks) > 0);
          tc::mma_commit(&mi.bar_stage[wid]);
          if (last) tc::mma_commit(&mi.bar_full[wid]);
        }
        ++n_stage_uses;
        if (last) { in_round = 0; ++n_rounds; } else { ++in_round; }
        __syncwarp();
      }
      // per-warp sums of the scalar certificates (fixed shuffle tree)
      double v;
      v = warp_sum(ent_dot); if (lane == 0) mi.ex[wid][kExEntDot] = v;
      v = warp_sum(ent_log); if (lane == 0) mi.ex[wid][kExEntLog] = v;
      v = warp_sum(sum_wdk); if (lane == 0) mi.ex[wid][kExSumWdk] = v;
      v = warp_sum(sum_wrs); if (lane == 0) mi.ex[wid][kExSumWrs] = v;
      v = warp_sum(n_rows);  if (lane == 0) mi.ex[wid][kExCount] = v;
      v = warp_max(mx_resp); if (lane == 0) mi.ex[wid][5] = v;
    }

    // =============================== epilogue warps ===============================
    double acc[2 * kNF];
    if (wid >= kProd) {
#pragma unroll
      for (int c = 0; c < 2 * kNF; ++c) acc[c] = 0.0;
      const int q = wid - kProd;
      const int n_seg = (int)(lt1 - lt0);
      const uint32_t lane_base = (uint32_t)(32 * q) << 16;
      int rounds[kProd];
#pragma unroll
      for (int w = 0; w < kProd; ++w) {
        const int tiles_w = n_seg > w ? (n_seg - w + kProd - 1) / kProd : 0;
        rounds[w] = (tiles_w + G.flush - 1) / G.flush;
      }
      for (int r = 0;; ++r) {
        bool any = false;
#pragma unroll
        for (int w = 0; w < kProd; ++w) {
          if (r < rounds[w]) {
            any = true;
            tc::mbar_wait(&mi.bar_full[w], n_drained[w] & 1);
            ++n_drained[w];
            tc::fence_after_sync();
            uint32_t a0[16], a1[16], a2[4], a3[2];
            const uint32_t addr = tmem + lane_base + w * kAccStride;
            tc::tmem_ld_x16(addr, a0);
            tc::tmem_ld_x16(addr + 16, a1);
            tc::tmem_ld_x4(addr + 32, a2);
            tc::tmem_ld_x2(addr + 36, a3);
            tc::tmem_ld_wait();
            tc::fence_before_sync();
            __syncwarp();
            if (lane == 0) tc::mbar_arrive(&mi.bar_empty[w]);
#pragma unroll
            for (int c = 0; c < 16; ++c) acc[c] += (double)__uint_as_float(a0[c]);
#pragma unroll
            for (int c = 0; c < 16; ++c) acc[16 + c] += (double)__uint_as_float(a1[c]);
#pragma unroll
            for (int c = 0; c < 4; ++c) acc[32 + c] += (double)__uint_as_float(a2[c]);
#pragma unroll
            for (int c = 0; c < 2; ++c) acc[36 + c] += (double)__uint_as_float(a3[c]);
          }
        }
        if (!any) break;
      }
    }

    // =============================== combine the segment ===============================
    __syncthreads();   // every MMA of the segment has completed (the epilogue waited for all of them)
    double* red = reinterpret_cast<double*>(stages);   // operand tiles are idle now
    if (wid >= kProd) {
      const int R = 32 * (wid - kProd) + lane;
#pragma unroll
      for (int f = 0; f < kNF; ++f) red[R * kNF + f] = acc[f] + acc[kNF + f];
    }
    __syncthreads();
    {
      const int slot = cta - cta_of_tile(G, unit_t0);
      double* part = P.partial + ((int64_t)u * G.n_parts + slot) * P.part_len;
      for (int idx = tid; idx < nb * kNF; idx += C::kThreads) {
        const int b = idx / kNF, f = idx - b * kNF;
        part[b * kRowLen + f] = red[b * kNF + f] + red[(C::kBinsPad + b) * kNF + f];
      }
      if (tid < kNExtras + kNMax) {
        double* ex = part + nb * kRowLen;
        double r = 0.0;
        if (tid <= kExCount) {
          for (int w = 0; w < kProd; ++w) r += mi.ex[w][tid];
          ex[tid] = r;
        } else if (tid < kNExtras) {
          ex[tid] = 0.0;
        } else if (tid == kNExtras + kMxResp) {
          for (int w = 0; w < kProd; ++w) r = fmax(r, mi.ex[w][5]);
          ex[tid] = r;
        } else {
          ex[tid] = 0.0;
        }
      }
    }
    __syncthreads();
    // the scratch area must read as finite floats again where padding rows alias it (any bit pattern is harmless for
    // the rows that matter, but keep NaN payloads out of the accumulator columns nobody reads)
    for (int k = tid; k < (C::kRows * kNF * 8 + 15) / 16; k += C::kThreads) reinterpret_cast<uint4*>(stages)[k] = make_uint4(0, 0, 0, 0);
    tc::fence_smem_to_async();
    __syncthreads();
    g0 = g1;
  }

  tc::fence_before_sync();
  __syncthreads();
  if (wid == 0) tc::tmem_free(tmem, 512);
}

int tc_flush_tiles() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("GCS_TC_FLUSH");
    v = e ? atoi(e) : 4;
    if (v < 1) v = 1;
    if (v > 64) v = 64;
  }
  return v;
}

TcGeom make_geom(int sm_count, int n_units, int64_t cap, int n_parts) {
  TcGeom G;
  G.tiles_per_unit = (cap + tc::kTileK - 1) / tc::kTileK;
  G.total_tiles = G.tiles_per_unit * n_units;
  int64_t n_cta = (G.total_tiles + kProd - 1) / kProd;
  if (n_cta > sm_count) n_cta = sm_count;
  if (n_cta < 1) n_cta = 1;
  G.n_cta = (int)n_cta;
  G.n_parts = n_parts;
  G.flush = tc_flush_tiles();
  return G;
}

template <int Q>
cudaError_t launch_q(cudaStream_t st, const BinScanParams& P, const TcGeom& G) {
  using C = TcCfg<Q>;
  static bool attr_set = false;
  const int smem = C::kStagesBytes + (int)sizeof(TcMisc) + 1024;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(bin_scan_tc_kernel<Q>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return e;
    attr_set = true;
  }
  bin_scan_tc_kernel<Q><<<G.n_cta, C::kThreads, smem, st>>>(P, G);
  return cudaSuccess;
}

}  // namespace

// a unit is touched by at most ceil(n_cta / U) + 1 CTAs
int bin_scan_tc_parts(int sm_count, int n_units, int64_t cap) {
  TcGeom G = make_geom(sm_count, n_units, cap, 0);
  return (G.n_cta + n_units - 1) / n_units + 1;
}

bool bin_scan_tc_supported(const BinScanParams& P) {
  return P.resp == nullptr && !P.use_true_max && P.n_bins <= kMaxBins;
}

cudaError_t launch_bin_scan_tc(int sm_count, cudaStream_t st, const BinScanParams& P, int n_parts) {
  const int U = P.n_scans * P.n_hyp;
  TcGeom G = make_geom(sm_count, U, P.cap, n_parts);
  return P.n_bins <= 48 ? launch_q<3>(st, P, G) : launch_q<4>(st, P, G);
}

}  // namespace gcs
