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
// Operand formats.  The default keeps every operand as two fp16 terms (kind::f16, K = 16 per MMA, 64-byte rows,
// SWIZZLE_64B): x = hi + lo / 2^11 with power-of-two pre-scales per operand class (kShift*), undone exactly on the
// float64 partials.  Half the operand bytes of the tf32 form, so every producer warp owns TWO operand tiles and starts
// its next tile while the tensor core still reads the previous one.  The tf32 form (kind::tf32, K = 8, 128-byte rows,
// SWIZZLE_128B, one tile per warp) is kept for sharp kernels (tau < 0.05) where fp16's exponent range is too small.
//
// A CTA is persistent over a contiguous range of 32-point tiles of the flattened (unit, tile) space -- every SM gets
// the same amount of work whatever the batch shape -- and writes one partial per unit segment it touched, in the
// layout of bin_scan_kernel, so reduce_partials_kernel and the finalize kernel are shared with the other precisions.
// Accumulation order is fixed (per-warp tile order, epilogue drains warps in index order): bit-identical reruns.
#include <stdio.h>
#include <stdlib.h>

#include "gcs_bins.cuh"
#include "gcs_tc.cuh"

namespace gcs {

namespace {

constexpr int kMaxProd = 12;        // producer warps per CTA (upper bound; see TcCfg)
constexpr int kMmaN = 40;           // >= 2 * kNF, multiple of 8 (N = 40 at M = 128 is a legal tcgen05 shape: tools/tc_probe.cu)
constexpr int kAccStride = kMmaN;   // TMEM columns between per-warp accumulators (12 x 40 = 480 <= 512)
constexpr int kBRows = 40;          // B tile rows kept in shared memory (38 used)
constexpr double kLog2e = 1.4426950408889634;
constexpr double kLn2 = 0.6931471805599453;

// operand scaling of the 16-bit variant (all powers of two, undone exactly on the float64 partials):
//   e' = e * 2^kShiftE (folded into the logit offset), every "lo" term is the residual * 2^kShiftLo, and the three feature
//   classes (1, d, d d^T), p, p p^T carry 2^kShiftD, 2^kShiftP, 2^kShiftPP.  With these, fp16 operands stay finite for
//   w / Z < 2000 and |p| < 500 m; outside that range the moments come out as inf / NaN (never silently wrong).
constexpr int kShiftE = 14, kShiftLo = 11, kShiftD = 5, kShiftP = -2, kShiftPP = -12;

template <int Q, bool H>
struct TcCfg {
  // producer warps: as many 17 KB / 21 KB operand tiles as fit in shared memory, in whole warpgroups
  static constexpr int kProd = Q <= 3 ? 12 : 8;
  static constexpr int kRegProd = Q <= 3 ? 136 : 184;   // setmaxnreg targets (launch allocation 128 per thread)
  static constexpr int kRegEpi = 104;
  static constexpr int kBinsPad = 16 * Q;   // rows of the e_hi block == first row of the e_lo block
  static constexpr int kRows = 32 * Q;      // rows of A that carry data
  // Warpgroup after the producers: epilogue warps, one per 32 TMEM lanes that carry data (Q of them).  For Q = 3 its
  // fourth warp is the MMA issuer; for Q = 4 the issuer is the first warp of one more (otherwise idle) warpgroup.
  static constexpr int kEpi = Q;
  static constexpr int kIssuerWarp = Q <= 3 ? kProd + 3 : kProd + 4;
  static constexpr int kThreads = Q <= 3 ? 32 * (kProd + 4) : 32 * (kProd + 8);
  static constexpr int kRowB = H ? tc::kRowBytes16 : tc::kRowBytes;   // bytes of one operand row (32 points)
  static constexpr int kNBuf = H ? 2 : 1;                             // operand tiles per producer warp
  static constexpr int kMmaPerTile = H ? 2 : 4;                       // K = 16 (f16) / 8 (tf32) points per MMA
  static constexpr int kABytes = kRows * kRowB;
  static constexpr int kBBytes = kBRows * kRowB;
  static constexpr int kStageBytes = kABytes + kBBytes;   // tf32: 17 / 21 KB (multiples of 1 KB); f16: 8.5 / 10.5 KB (of 512 B)
  static constexpr int kStagesBytes = kProd * kNBuf * kStageBytes + 1024;   // + tail read by the last B tile's padding rows
};

struct TcMisc {
  float4 bins2[kMaxBins * 2];  // per bin: (bx, bx, by, by), (bz, bz, -c2, -c2), pre-scaled by log2(e)/tau
  uint64_t bar_tile[kMaxProd][2];    // the producer warp has written (and fenced) its operand tile (buffer 0 / 1)
  uint64_t bar_stage[kMaxProd][2];   // MMAs that read the warp's operand tile (buffer 0 / 1) have completed
  uint64_t bar_full[kMaxProd];    // the warp's accumulator holds a finished round
  uint64_t bar_empty[kMaxProd];   // the epilogue has drained it
  uint32_t tmem;
  double ex[kMaxProd][8];
  TwistCtx tw[kMaxProd];        // per producer warp: hoisted invariants of the current unit (reloaded every tile)
  WindowCtx win[kMaxProd];
};

struct TcGeom {
  int64_t tiles_per_unit;   // ceil(cap / 32)
  int64_t total_tiles;      // U * tiles_per_unit
  int n_cta, n_parts, flush;
  int dbg;   // 0, or 1 + slot of the timing hook
};

__device__ __forceinline__ int64_t cta_tile0(const TcGeom& G, int c) { return (int64_t)c * G.total_tiles / G.n_cta; }
// CTA whose range contains tile g
__device__ __forceinline__ int cta_of_tile(const TcGeom& G, int64_t g) {
  return (int)(((g + 1) * G.n_cta - 1) / G.total_tiles);
}

// segment of one unit handled by this CTA
struct TcSeg { int u, s, h; int64_t unit_t0, lt0, lt1; };

__device__ __forceinline__ void bar_all(int n_threads) { asm volatile("bar.sync 1, %0;" ::"r"(n_threads) : "memory"); }

__device__ __forceinline__ bool next_segment(const TcGeom& G, int n_hyp, int64_t& g0, int64_t g_end, TcSeg& sg) {
  if (g0 >= g_end) return false;
  sg.u = (int)(g0 / G.tiles_per_unit);
  sg.unit_t0 = (int64_t)sg.u * G.tiles_per_unit;
  const int64_t g1 = (sg.unit_t0 + G.tiles_per_unit < g_end) ? sg.unit_t0 + G.tiles_per_unit : g_end;
  sg.lt0 = g0 - sg.unit_t0; sg.lt1 = g1 - sg.unit_t0;
  sg.s = sg.u / n_hyp; sg.h = sg.u - sg.s * n_hyp;
  g0 = g1;
  return true;
}

// Segment combine, executed by all threads of the CTA after the epilogue has written its row sums to `red`.
template <int Q, bool H>
__device__ __forceinline__ void write_partial(const BinScanParams& P, const TcGeom& G, const TcSeg& sg, const TcMisc& mi,
                                              const double* red, int cta, int tid) {
  using C = TcCfg<Q, H>;
  constexpr int kProd = C::kProd;
  const int nb = P.n_bins;
  const int slot = cta - cta_of_tile(G, sg.unit_t0);
  double* part = P.partial + ((int64_t)sg.u * G.n_parts + slot) * P.part_len;
  for (int idx = tid; idx < nb * kNF; idx += C::kThreads) {
    const int b = idx / kNF, f = idx - b * kNF;
    if (H) {
      const double sc_f = ldexp(1.0, -kShiftE - (f < 10 ? kShiftD : (f < 13 ? kShiftP : kShiftPP)));
      part[b * kRowLen + f] = fma(red[(C::kBinsPad + b) * kNF + f], ldexp(1.0, -kShiftLo), red[b * kNF + f]) * sc_f;
    } else {
      part[b * kRowLen + f] = red[b * kNF + f] + red[(C::kBinsPad + b) * kNF + f];
    }
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

// Barrier schedule of a segment (all kThreads threads, named barrier 1):
//   [producers: tiles + MMAs | epilogue: drains]  B1  [epilogue: row sums -> red]  B2  [all: partial]  B3  [all: re-zero red]  B4
template <int Q, bool H>
__device__ __forceinline__ void segment_tail(const BinScanParams& P, const TcGeom& G, const TcSeg& sg, TcMisc& mi,
                                             unsigned char* stages, int cta, int tid, const double* acc_or_null) {
  using C = TcCfg<Q, H>;
  constexpr int kProd = C::kProd;
  double* red = reinterpret_cast<double*>(stages);   // operand tiles are idle: every MMA of the segment has completed
  bar_all(C::kThreads);
  if (acc_or_null) {
    const int R = tid - 32 * kProd;
#pragma unroll
    for (int f = 0; f < kNF; ++f) red[R * kNF + f] = acc_or_null[f];
  }
  bar_all(C::kThreads);
  write_partial<Q, H>(P, G, sg, mi, red, cta, tid);
  bar_all(C::kThreads);
  // padding rows of the operand tiles alias this scratch: keep it free of NaN bit patterns
  for (int k = tid; k < (128 * kNF * 8 + 15) / 16; k += C::kThreads) reinterpret_cast<uint4*>(stages)[k] = make_uint4(0, 0, 0, 0);
  tc::fence_smem_to_async();
  bar_all(C::kThreads);
}

template <int Q, bool H>
__device__ __forceinline__ void producer_role(const BinScanParams& P, const TcGeom& G, TcMisc& mi, unsigned char* stages,
                                              uint32_t tmem, int cta, int tid) {
  using C = TcCfg<Q, H>;
  constexpr int kProd = C::kProd;
  const int wid = tid >> 5, lane = tid & 31;
  uint32_t n_stage_uses = 0;   // uses of the operand tile
  // lane-dependent byte offsets of this lane's element inside a row, for the 8 row phases of the swizzle
  uint32_t off[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) off[j] = (uint32_t)((((lane >> 2) ^ j) << 4) | ((lane & 3) << 2));
  // same for the pair layout (points 2j, 2j+1 of the tile, j = lane & 15): one 8-byte store per row
  uint32_t off2[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) off2[j] = (uint32_t)(((((lane & 15) >> 1) ^ j) << 4) | ((lane & 1) << 3));
  // 16-bit rows (64 B, SWIZZLE_64B: chunk ^= (row >> 1) & 3): element = point (2 B) / point pair (4 B)
  uint32_t offh[4], off2h[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    offh[j] = (uint32_t)((((lane >> 3) ^ j) << 4) | ((lane & 7) << 1));
    off2h[j] = (uint32_t)(((((lane & 15) >> 2) ^ j) << 4) | ((lane & 3) << 2));
  }
  int64_t g0 = cta_tile0(G, cta);
  const int64_t g_end = cta_tile0(G, cta + 1);
  TcSeg sg;
  while (next_segment(G, P.n_hyp, g0, g_end, sg)) {
    const int u = sg.u, s = sg.s, h = sg.h;
    const int64_t lt0 = sg.lt0, lt1 = sg.lt1;
    unsigned char* const sA0 = stages + wid * C::kNBuf * C::kStageBytes;
    const double t0 = P.t0s[s], t1 = P.t1s[s];
    const double inv_denom = 1.0 / fmax(t1 - t0, 1e-12);
    __syncwarp();
    if (lane == 0) { mi.win[wid] = make_window_ctx(t0, t1); mi.tw[wid] = make_twist_ctx(P.xi + (int64_t)u * 6); }
    __syncwarp();
    const double mass_scale = P.mass[s * kNMass + kMassAll] / (P.mass[s * kNMass + kMassSel] + P.eps_mass);
    const double* pts = P.pts + (int64_t)s * P.n_raw * 3;
    const double* tp = P.t + (int64_t)s * P.n_raw;
    const double* wp = P.w + (int64_t)s * P.n_raw;
    const uint8_t* rp = P.ring ? P.ring + (int64_t)s * P.n_raw : nullptr;
    const uint8_t* gp = P.tag ? P.tag + (int64_t)s * P.n_raw : nullptr;
    double ent_dot = 0.0, ent_log = 0.0, mx_resp = 0.0, sum_wdk = 0.0, sum_wrs = 0.0, n_rows = 0.0;
    // pair layout of the soft-assign stage: lane = (half, j); the lane evaluates bins [half*NB2, half*NB2 + NB2) for
    // the two points 2j, 2j+1 of the tile (packed f32x2 math, 8-byte operand stores)
    constexpr int NB2 = C::kBinsPad / 2;
    const int half = lane >> 4, pj = lane & 15;
    const float4* bins_lane = mi.bins2 + half * 2;   // entry i of this half at bins_lane[4 * i], [4 * i + 1]

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
      // ---------------- stage 1 (lane = point): resample gather, deskew, window weight, ray direction (float64)
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
      deskew_point_ctx(p, alpha, mi.tw[wid], p0);
      const double w_dk = w_rs * window_weight_ctx(tt, mi.win[wid]);
      if (row) {
        const int64_t o = (int64_t)u * P.cap + i;
        if (P.dk_pts) { P.dk_pts[3 * o] = p0[0]; P.dk_pts[3 * o + 1] = p0[1]; P.dk_pts[3 * o + 2] = p0[2]; }
        if (P.dk_w) P.dk_w[o] = w_dk;
        sum_wdk += w_dk; sum_wrs += w_rs; n_rows += 1.0;
      }
      const double r0 = p0[0] - P.origin[0], r1 = p0[1] - P.origin[1], r2 = p0[2] - P.origin[2];
      const double rr = fma(r0, r0, fma(r1, r1, r2 * r2));
      const double y = fast_rsqrt(fmax(rr, 1e-300));
      const double invn = fma(-P.eps_mass * y, y, y);          // 1 / (|r| + eps) to first order in eps / |r|
      const float f0 = (float)(r0 * invn), f1 = (float)(r1 * invn), f2 = (float)(r2 * invn);

      // ---------------- stage 2 (pair layout): A = [e_hi ; e_lo], softmax numerators in log2 units, unnormalised
      const float2 g0 = make_float2(__shfl_sync(0xffffffffu, f0, 2 * pj), __shfl_sync(0xffffffffu, f0, 2 * pj + 1));
      const float2 g1 = make_float2(__shfl_sync(0xffffffffu, f1, 2 * pj), __shfl_sync(0xffffffffu, f1, 2 * pj + 1));
      const float2 g2 = make_float2(__shfl_sync(0xffffffffu, f2, 2 * pj), __shfl_sync(0xffffffffu, f2, 2 * pj + 1));
      // the tensor core may still be reading the operand tile this one goes into
      const uint32_t buf = H ? (n_stage_uses & 1u) : 0u;
      unsigned char* const sA = sA0 + buf * C::kStageBytes;
      unsigned char* const sB = sA + C::kABytes;
      // tf32: the lane's bins are [half * NB2, half * NB2 + NB2); f16: bins 2 i + half (rows of the two halves then fall
      // into different banks: a 64-byte row covers half of them)
      unsigned char* const sA_lane = sA + (H ? half * tc::kRowBytes16 : half * NB2 * tc::kRowBytes);
      if (n_stage_uses >= (uint32_t)C::kNBuf)
        tc::mbar_wait(&mi.bar_stage[wid][buf], ((n_stage_uses / C::kNBuf) - 1) & 1);
      float2 sum = make_float2(0.f, 0.f), dot = sum;
      float mx0 = 0.f, mx1 = 0.f;
      // groups of kGrp bins: table loads of the next group are issued before the operand stores of this one (the
      // compiler cannot move shared-memory loads across those stores by itself), math of a group is independent
      constexpr int kGrp = 4;
      static_assert(NB2 % kGrp == 0, "bin half must be a multiple of the group size");
      float4 tab[2][kGrp][2];
#pragma unroll
      for (int k = 0; k < kGrp; ++k) { tab[0][k][0] = bins_lane[4 * k]; tab[0][k][1] = bins_lane[4 * k + 1]; }
#pragma unroll
      for (int g = 0; g < NB2 / kGrp; ++g) {
        const int cur = g & 1;
        if (g + 1 < NB2 / kGrp) {
#pragma unroll
          for (int k = 0; k < kGrp; ++k) {
            tab[cur ^ 1][k][0] = bins_lane[4 * ((g + 1) * kGrp + k)];
            tab[cur ^ 1][k][1] = bins_lane[4 * ((g + 1) * kGrp + k) + 1];
          }
        }
        float2 l[kGrp], e[kGrp], hi[kGrp], lo[kGrp];
#pragma unroll
        for (int k = 0; k < kGrp; ++k) {
          const float4 ta = tab[cur][k][0], tb = tab[cur][k][1];
          l[k] = tc::fma2(g0, make_float2(ta.x, ta.y),
                          tc::fma2(g1, make_float2(ta.z, ta.w), tc::fma2(g2, make_float2(tb.x, tb.y), make_float2(tb.z, tb.w))));
        }
#pragma unroll
        for (int k = 0; k < kGrp; ++k) e[k] = make_float2(tc::ex2f(l[k].x), tc::ex2f(l[k].y));
#pragma unroll
        for (int k = 0; k < kGrp; ++k) {
          sum = tc::add2(sum, e[k]);
          dot = tc::fma2(e[k], l[k], dot);
          mx0 = fmaxf(mx0, e[k].x); mx1 = fmaxf(mx1, e[k].y);
          if (!H) {
            hi[k] = make_float2(tc::tf32_hi(e[k].x), tc::tf32_hi(e[k].y));
            lo[k] = tc::sub2(e[k], hi[k]);
          }
        }
        if (!H) {
#pragma unroll
          for (int k = 0; k < kGrp; ++k) {
            const int b = g * kGrp + k;
            *reinterpret_cast<float2*>(sA_lane + b * tc::kRowBytes + off2[b & 7]) = hi[k];
            *reinterpret_cast<float2*>(sA_lane + (C::kBinsPad + b) * tc::kRowBytes + off2[b & 7]) = lo[k];
          }
        } else {
          // e' = hi + lo / 2^kShiftLo, both fp16: one packed conversion per point pair, residual exact in float32
          uint32_t h2[kGrp], l2[kGrp];
#pragma unroll
          for (int k = 0; k < kGrp; ++k) {
            h2[k] = tc::pack_f16x2(e[k].x, e[k].y);
            const float2 r = tc::sub2(e[k], tc::unpack_f16x2(h2[k]));
            const float2 rs = tc::mul2(r, make_float2((float)(1 << kShiftLo), (float)(1 << kShiftLo)));
            l2[k] = tc::pack_f16x2(rs.x, rs.y);
          }
#pragma unroll
          for (int k = 0; k < kGrp; ++k) {
            const int i = g * kGrp + k;   // row 2 i + half; lo rows kBinsPad further down (same swizzle phase)
            const int ro = (i >> 2) * tc::kGroupBytes16 + ((2 * i) & 7) * tc::kRowBytes16;
            *reinterpret_cast<uint32_t*>(sA_lane + ro + off2h[i & 3]) = h2[k];
            *reinterpret_cast<uint32_t*>(sA_lane + C::kBinsPad * tc::kRowBytes16 + ro + off2h[i & 3]) = l2[k];
          }
        }
      }
      // both bin halves -> every lane holds the full row sums of its two points; then back to lane = point
      sum.x += __shfl_xor_sync(0xffffffffu, sum.x, 16); sum.y += __shfl_xor_sync(0xffffffffu, sum.y, 16);
      dot.x += __shfl_xor_sync(0xffffffffu, dot.x, 16); dot.y += __shfl_xor_sync(0xffffffffu, dot.y, 16);
      mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 16)); mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 16));
      const int src = lane >> 1;
      const bool odd = lane & 1;
      const float sa = __shfl_sync(0xffffffffu, sum.x, src), sb = __shfl_sync(0xffffffffu, sum.y, src);
      const float da_ = __shfl_sync(0xffffffffu, dot.x, src), db_ = __shfl_sync(0xffffffffu, dot.y, src);
      const float ma = __shfl_sync(0xffffffffu, mx0, src), mb = __shfl_sync(0xffffffffu, mx1, src);
      const float sumf = odd ? sb : sa, dotf = odd ? db_ : da_, emax = odd ? mb : ma;

      // ---------------- stage 3 (lane = point): B = [phi_hi ; phi_lo], phi = (w / Z) (1, d, d d^T, p, p p^T)
      const double ssum = (double)sumf;
      double inv;
      {
        float rf;
        asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(rf) : "f"(sumf));
        inv = (double)rf;
        inv = fma(inv, fma(-ssum, inv, 1.0), inv);
        inv = fma(inv, fma(-ssum, inv, 1.0), inv);
      }
      if (row) {
        ent_dot = fma(inv * kLn2, (double)dotf, ent_dot);
        ent_log = fma((double)__log2f(sumf), kLn2, ent_log);
        mx_resp = fmax(mx_resp, (double)emax * inv);
      }
      {
        const float sc0 = row ? (float)(w_dk * inv) : 0.f;
        // 16-bit operands: inv is 1 / (2^kShiftE Z); the feature classes carry their own power-of-two scales
        const float sc = H ? sc0 * (float)(1 << (kShiftE + kShiftD)) : sc0;
        const float scp = H ? sc0 * (float)(1 << (kShiftE + kShiftP)) : sc0;
        const float scpp = H ? sc0 * (float)(1 << (kShiftE + kShiftPP)) : sc0;
        const float q0 = (float)p0[0], q1 = (float)p0[1], q2 = (float)p0[2];
        const float sd0 = sc * f0, sd1 = sc * f1, sd2 = sc * f2;
        const float sp0 = scp * q0, sp1 = scp * q1, sp2 = scp * q2;
        const float sq0 = scpp * q0, sq1 = scpp * q1, sq2 = scpp * q2;
        const float v[kNF] = {sc, sd0, sd1, sd2, sd0 * f0, sd0 * f1, sd0 * f2, sd1 * f1, sd1 * f2, sd2 * f2,
                              sp0, sp1, sp2, sq0 * q0, sq0 * q1, sq0 * q2, sq1 * q1, sq1 * q2, sq2 * q2};
        if (!H) {
#pragma unroll
          for (int f = 0; f < kNF; ++f) {
            float hi, lo;
            tc::split_tf32(v[f], hi, lo);
            *reinterpret_cast<float*>(sB + f * tc::kRowBytes + off[f & 7]) = hi;
            *reinterpret_cast<float*>(sB + (kNF + f) * tc::kRowBytes + off[(kNF + f) & 7]) = lo;
          }
        } else {
#pragma unroll
          for (int f = 0; f < kNF; ++f) {
            const __half hi = __float2half_rn(v[f]);
            const __half lo = __float2half_rn((v[f] - __half2float(hi)) * (float)(1 << kShiftLo));
            constexpr int r1 = kNF;
            *reinterpret_cast<__half*>(sB + tc::row_base16(f) + offh[(f >> 1) & 3]) = hi;
            *reinterpret_cast<__half*>(sB + tc::row_base16(r1 + f) + offh[((r1 + f) >> 1) & 3]) = lo;
          }
        }
      }
      tc::fence_smem_to_async();
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(&mi.bar_tile[wid][buf]);   // the issuer warp takes it from here
      ++n_stage_uses;
    }
    // per-warp sums of the scalar certificates (fixed shuffle tree)
    double v;
    v = warp_sum(ent_dot); if (lane == 0) mi.ex[wid][kExEntDot] = v;
    v = warp_sum(ent_log); if (lane == 0) mi.ex[wid][kExEntLog] = v;
    v = warp_sum(sum_wdk); if (lane == 0) mi.ex[wid][kExSumWdk] = v;
    v = warp_sum(sum_wrs); if (lane == 0) mi.ex[wid][kExSumWrs] = v;
    v = warp_sum(n_rows);  if (lane == 0) mi.ex[wid][kExCount] = v;
    v = warp_max(mx_resp); if (lane == 0) mi.ex[wid][5] = v;
    segment_tail<Q, H>(P, G, sg, mi, stages, cta, tid, nullptr);
  }
}

template <int Q, bool H>
__device__ __forceinline__ void epilogue_role(const BinScanParams& P, const TcGeom& G, TcMisc& mi, unsigned char* stages,
                                              uint32_t tmem, int cta, int tid) {
  constexpr int kProd = TcCfg<Q, H>::kProd;
  const int wid = tid >> 5, lane = tid & 31;
  uint32_t n_drained[kProd];   // rounds drained per producer warp
#pragma unroll
  for (int w = 0; w < kProd; ++w) n_drained[w] = 0;
  int64_t g0 = cta_tile0(G, cta);
  const int64_t g_end = cta_tile0(G, cta + 1);
  TcSeg sg;
  while (next_segment(G, P.n_hyp, g0, g_end, sg)) {
    const int64_t lt0 = sg.lt0, lt1 = sg.lt1;
    double acc[kNF];   // per TMEM lane (= operand row): hi-feature column + lo-feature column
#pragma unroll
    for (int c = 0; c < kNF; ++c) acc[c] = 0.0;
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
          // column f holds row x phi_hi[f], column 19 + f row x phi_lo[f] (2^-11 of the former): one float32 add, then
          // the float64 accumulation
          float v[2 * kNF];
#pragma unroll
          for (int c = 0; c < 16; ++c) { v[c] = __uint_as_float(a0[c]); v[16 + c] = __uint_as_float(a1[c]); }
#pragma unroll
          for (int c = 0; c < 4; ++c) v[32 + c] = __uint_as_float(a2[c]);
          v[36] = __uint_as_float(a3[0]); v[37] = __uint_as_float(a3[1]);
#pragma unroll
          for (int f = 0; f < kNF; ++f)
            acc[f] += (double)(H ? fmaf(v[kNF + f], 1.0f / (float)(1 << kShiftLo), v[f]) : v[f] + v[kNF + f]);
        }
      }
      if (!any) break;
    }
    segment_tail<Q, H>(P, G, sg, mi, stages, cta, tid, acc);
  }
}

// MMA issuer: one thread feeds the tensor core for the whole CTA; producers never block on the tensor-core queue.  Lane w
// of the issuer warp watches the barriers of producer warp w (one mbarrier.test_wait polls all of them at once) and
// lane 0 issues whichever tiles are ready -- no fixed order across warps, so a late warp does not hold up the MMAs (and
// hence the operand-tile release) of the others.  Each accumulator still sees only its own warp's tiles, in sequence,
// and the epilogue drains in a fixed order: results do not depend on the issue order.
template <int Q, bool H>
__device__ __forceinline__ void issuer_role(const BinScanParams& P, const TcGeom& G, TcMisc& mi, unsigned char* stages,
                                            uint32_t tmem, int cta, int tid) {
  using C = TcCfg<Q, H>;
  constexpr int kProd = C::kProd;
  const int lane = tid & 31;
  int64_t g0 = cta_tile0(G, cta);
  const int64_t g_end = cta_tile0(G, cta + 1);
  const uint32_t idesc = H ? tc::idesc_f16(128, kMmaN) : tc::idesc_tf32(128, kMmaN);
  const uint32_t a0 = tc::smem_u32(stages);
  TcSeg sg;
  // lane w (< kProd) keeps the tile / round counters of producer warp w and polls its barriers; lane 0 issues
  uint32_t n_done = 0;      // tiles of this lane's warp issued so far: operand buffer n_done % kNBuf, phase n_done / kNBuf
  uint32_t par_empty = 0;   // phase of bar_empty to test next (the warp's previous round)
  bool have_round = false;                // a round of this lane's warp has been handed to the epilogue
  while (next_segment(G, P.n_hyp, g0, g_end, sg)) {
    {
      const int n_seg = (int)(sg.lt1 - sg.lt0);
      const int my_tiles = (lane < kProd && n_seg > lane) ? (n_seg - lane + kProd - 1) / kProd : 0;
      int t = 0, in_round = 0;
      for (;;) {
        const bool pending = t < my_tiles;
        bool ready = false;
        if (pending) {
          ready = tc::mbar_test_wait(&mi.bar_tile[lane][n_done % C::kNBuf], (n_done / C::kNBuf) & 1u);
          if (ready && in_round == 0 && have_round) ready = tc::mbar_test_wait(&mi.bar_empty[lane], par_empty);
        }
        const unsigned rdy = __ballot_sync(0xffffffffu, ready);
        if (rdy == 0u) {
          if (__ballot_sync(0xffffffffu, pending) == 0u) break;
          continue;
        }
        const bool last = (in_round + 1 == G.flush) || (t + 1 >= my_tiles);
        const unsigned first_m = __ballot_sync(0xffffffffu, ready && in_round == 0);
        const unsigned last_m = __ballot_sync(0xffffffffu, ready && last);
        const unsigned buf_m = H ? __ballot_sync(0xffffffffu, ready && (n_done & 1u)) : 0u;   // operand buffer of the tile
        tc::fence_after_sync();
        if (lane == 0) {
          unsigned m = rdy;
          while (m) {
            const int w = __ffs(m) - 1;
            m &= m - 1;
            const uint32_t buf = (buf_m >> w) & 1u;
            const uint32_t sa = a0 + (w * C::kNBuf + buf) * C::kStageBytes;
            const uint64_t da = H ? tc::smem_desc_sw64(sa) : tc::smem_desc_sw128(sa);
            const uint64_t db = H ? tc::smem_desc_sw64(sa + C::kABytes) : tc::smem_desc_sw128(sa + C::kABytes);
            const uint32_t d_tmem = tmem + w * kAccStride;
            const uint32_t acc0 = ((first_m >> w) & 1u) ^ 1u;
            // one MMA consumes 32 bytes of every operand row (8 tf32 / 16 fp16 points): descriptor start + 2 (x 16 B)
#pragma unroll
            for (int ks = 0; ks < C::kMmaPerTile; ++ks) {
              if (H) tc::mma_f16_ss(d_tmem, da + 2 * ks, db + 2 * ks, idesc, ks ? 1u : acc0);
              else tc::mma_tf32_ss(d_tmem, da + 2 * ks, db + 2 * ks, idesc, ks ? 1u : acc0);
            }
            tc::mma_commit(&mi.bar_stage[w][buf]);
            if ((last_m >> w) & 1u) tc::mma_commit(&mi.bar_full[w]);
          }
        }
        if (ready) {
          ++t;
          ++n_done;
          if (last) { in_round = 0; if (have_round) par_empty ^= 1u; have_round = true; } else ++in_round;
        }
        __syncwarp();
      }
    }
    __syncwarp();
    segment_tail<Q, H>(P, G, sg, mi, stages, cta, tid, nullptr);
  }
}

// warps of the CTA that have no role in a configuration still take part in the segment barriers
template <int Q, bool H>
__device__ __forceinline__ void idle_role(const BinScanParams& P, const TcGeom& G, TcMisc& mi, unsigned char* stages, int cta,
                                          int tid) {
  int64_t g0 = cta_tile0(G, cta);
  const int64_t g_end = cta_tile0(G, cta + 1);
  TcSeg sg;
  while (next_segment(G, P.n_hyp, g0, g_end, sg)) segment_tail<Q, H>(P, G, sg, mi, stages, cta, tid, nullptr);
}

// Timing hook (GCS_TC_TIMES=1): first CTA entry and last CTA exit of the first 32 launches on %globaltimer, printed
// after the 24th launch -- tells the launch / drain overhead apart from the time the CTAs really run (8 us of 530).
__device__ unsigned long long g_dbg_t[64];
template <int Q, bool H>
__global__ void __launch_bounds__(TcCfg<Q, H>::kThreads, 1) bin_scan_tc_kernel(const BinScanParams P, const TcGeom G) {
  if (G.dbg && threadIdx.x == 0) { unsigned long long g; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(g)); atomicMin(&g_dbg_t[2 * (G.dbg - 1)], g); }
  using C = TcCfg<Q, H>;
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  // 1 KB alignment by pointer arithmetic on the __shared__ symbol (keeps the shared address space: STS/LDS, not generic)
  unsigned char* stages = smem_raw + ((1024u - (tc::smem_u32(smem_raw) & 1023u)) & 1023u);
  TcMisc& mi = *reinterpret_cast<TcMisc*>(stages + C::kStagesBytes);
  constexpr int kProd = C::kProd;
  const int tid = threadIdx.x, wid = tid >> 5;
  const int nb = P.n_bins;

  // ---- one-time setup
  for (int k = tid; k < C::kStagesBytes / 16; k += C::kThreads) reinterpret_cast<uint4*>(stages)[k] = make_uint4(0, 0, 0, 0);
  {
    const double sc = P.inv_tau * kLog2e;
    const float c2 = (float)(P.shift * kLog2e);
    for (int b = tid; b < C::kBinsPad; b += C::kThreads) {
      float x = 0.f, y = 0.f, z = 0.f, w = -1.0e30f;   // bins past n_bins: e = 2^(-1e30) = 0
      if (b < nb) {
        x = (float)(P.bin_dirs[3 * b] * sc); y = (float)(P.bin_dirs[3 * b + 1] * sc); z = (float)(P.bin_dirs[3 * b + 2] * sc);
        w = H ? (float)kShiftE - c2 : -c2;   // 16-bit operands: e' = 2^kShiftE e
      }
      // interleave the two bin halves (lanes 0-15 / 16-31 read entry i of their half in the same instruction): the two
      // 16-byte reads of a warp then fall into different banks
      const int e = H ? b : (b % (C::kBinsPad / 2)) * 2 + b / (C::kBinsPad / 2);
      mi.bins2[2 * e] = make_float4(x, x, y, y);
      mi.bins2[2 * e + 1] = make_float4(z, z, w, w);
    }
  }
  if (tid == 0) {
    for (int w = 0; w < kProd; ++w) {
      tc::mbar_init(&mi.bar_tile[w][0], 1);
      tc::mbar_init(&mi.bar_tile[w][1], 1);
      tc::mbar_init(&mi.bar_stage[w][0], 1);
      tc::mbar_init(&mi.bar_stage[w][1], 1);
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

  // ---- roles.  Register re-allocation: the producer warpgroups take what the other warpgroups give up.
  if (wid < kProd) {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(C::kRegProd));
    producer_role<Q, H>(P, G, mi, stages, tmem, blockIdx.x, tid);
  } else if (wid < kProd + 4) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(C::kRegEpi));
    if (wid < kProd + C::kEpi) epilogue_role<Q, H>(P, G, mi, stages, tmem, blockIdx.x, tid);
    else issuer_role<Q, H>(P, G, mi, stages, tmem, blockIdx.x, tid);
  } else {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
    if (wid == C::kIssuerWarp) issuer_role<Q, H>(P, G, mi, stages, tmem, blockIdx.x, tid);
    else idle_role<Q, H>(P, G, mi, stages, blockIdx.x, tid);
  }

  tc::fence_before_sync();
  __syncthreads();
  if (wid == 0) tc::tmem_free(tmem, 512);
  if (G.dbg && threadIdx.x == 0) { unsigned long long g; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(g)); atomicMax(&g_dbg_t[2 * (G.dbg - 1) + 1], g); }
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

TcGeom make_geom(int sm_count, int n_units, int64_t cap, int n_parts, int n_prod) {
  const int kProd = n_prod;
  TcGeom G;
  G.tiles_per_unit = (cap + tc::kTileK - 1) / tc::kTileK;
  G.total_tiles = G.tiles_per_unit * n_units;
  int64_t n_cta = (G.total_tiles + kProd - 1) / kProd;
  if (n_cta > sm_count) n_cta = sm_count;
  if (n_cta < 1) n_cta = 1;
  G.n_cta = (int)n_cta;
  G.n_parts = n_parts;
  G.flush = tc_flush_tiles();
  G.dbg = getenv("GCS_TC_TIMES") ? 1 : 0;
  return G;
}

// Operand format.  fp16 hi/lo operands need bounded magnitudes (see kShift*): w / Z stays small as long as every ray is
// within a few tau (in cosine) of some bin, which holds for any atlas that covers the sphere when tau >= 0.05 (the
// 48-bin Fibonacci atlas: Z >= exp(-0.045 / tau)).  Sharper kernels use the tf32 operands, whose exponent range is that
// of float32.  GCS_TC_OPERANDS=tf32 / f16 forces one variant (A/B runs).
bool tc_use_f16(double inv_tau) {
  static int forced = -2;
  if (forced == -2) {
    const char* e = getenv("GCS_TC_OPERANDS");
    forced = !e ? -1 : (e[0] == 't' ? 0 : 1);
  }
  return forced >= 0 ? forced == 1 : inv_tau <= 20.0;
}

template <int Q, bool H>
cudaError_t launch_q(cudaStream_t st, const BinScanParams& P, const TcGeom& G) {
  using C = TcCfg<Q, H>;
  const int smem = C::kStagesBytes + (int)sizeof(TcMisc) + 1024;
  {
    cudaError_t e = gcs_smem_attr_once((const void*)bin_scan_tc_kernel<Q, H>, smem);
    if (e != cudaSuccess) return e;
  }
  static int n_dbg = 0;
  TcGeom G2 = G;
  if (G.dbg) {
    if (n_dbg == 0) {
      unsigned long long z[64];
      for (int i = 0; i < 32; ++i) { z[2 * i] = ~0ull; z[2 * i + 1] = 0ull; }
      cudaMemcpyToSymbol(g_dbg_t, z, sizeof(z));
    }
    G2.dbg = n_dbg < 32 ? n_dbg + 1 : 0;
  }
  bin_scan_tc_kernel<Q, H><<<G.n_cta, C::kThreads, smem, st>>>(P, G2);
  if (G.dbg && ++n_dbg == 24) {
    cudaDeviceSynchronize();
    unsigned long long t[64];
    cudaMemcpyFromSymbol(t, g_dbg_t, sizeof(t));
    for (int i = 0; i < 24; ++i)
      fprintf(stderr, "tc launch %d: first CTA entry -> last CTA exit %.1f us; gap to next entry %.1f us\n", i,
              (t[2 * i + 1] - t[2 * i]) * 1e-3, i < 23 ? ((double)t[2 * i + 2] - (double)t[2 * i + 1]) * 1e-3 : 0.0);
  }
  return cudaSuccess;
}

}  // namespace

// a unit is touched by at most ceil(n_cta / U) + 1 CTAs
int bin_scan_tc_parts(int sm_count, int n_units, int64_t cap) {
  TcGeom G = make_geom(sm_count, n_units, cap, 0, 8);   // fewest producer warps of any instantiation: most CTAs
  return (G.n_cta + n_units - 1) / n_units + 1;
}

bool bin_scan_tc_supported(const BinScanParams& P) {
  return P.resp == nullptr && !P.use_true_max && P.n_bins <= kMaxBins;
}

cudaError_t launch_bin_scan_tc(int sm_count, cudaStream_t st, const BinScanParams& P, int n_parts) {
  const int U = P.n_scans * P.n_hyp;
  if (tc_use_f16(P.inv_tau)) {
    if (P.n_bins <= 48) return launch_q<3, true>(st, P, make_geom(sm_count, U, P.cap, n_parts, TcCfg<3, true>::kProd));
    return launch_q<4, true>(st, P, make_geom(sm_count, U, P.cap, n_parts, TcCfg<4, true>::kProd));
  }
  if (P.n_bins <= 48) return launch_q<3, false>(st, P, make_geom(sm_count, U, P.cap, n_parts, TcCfg<3, false>::kProd));
  return launch_q<4, false>(st, P, make_geom(sm_count, U, P.cap, n_parts, TcCfg<4, false>::kProd));
}

}  // namespace gcs
