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
// Operand formats.  The default (H = true) keeps every operand as two fp16 terms (kind::f16, K = 16 per MMA) with
// power-of-two pre-scales per operand class (kShift*), undone exactly on the float64 partials, in an MN-MAJOR,
// unswizzled layout (tools/tc_probe_mn.cu): the 8-element chunks of one point (one K row) are 16 bytes each, so a lane
// stores its point's 96 A values and 40 B values with 12 + 5 STS.128 instead of 134 two-byte stores, the eight lanes
// of a quarter warp fill one 128-byte core matrix (conflict-free), and no per-row swizzle offsets are needed.  A tile is
// 64 points: every lane owns the two consecutive points 2 lane, 2 lane + 1 (16-byte global loads / stores of the pair,
// packed f32x2 arithmetic across the pair) and writes them to K rows lane and 32 + lane (the order inside a tile is
// irrelevant to the contraction).  The deskew increment p0 - p is evaluated in float32 and added to the float64 point
// (error <= 2^-24 |p0 - p|); sweep fraction, window weight and all sums stay float64.
// The tf32 form (H = false; kind::tf32, K = 8, K-major 128-byte rows, SWIZZLE_128B, 32-point tiles, all-float64
// geometry) is kept for sharp kernels (tau < 0.05) where fp16's exponent range is too small.
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
// MN-major fp16 tiles: bytes between two groups of 8 points (K groups) of the B tile (5 chunks of 8 features)
constexpr int kKgB = 5 * 128;
#ifndef GCS_TC_SLEEP
#define GCS_TC_SLEEP 64
#endif
constexpr unsigned kIssuerSleepNs = GCS_TC_SLEEP;
// fp16 (MN-major) producer: lane 0 of every producer warp issues the MMAs of its own tile (no hop through the issuer
// warp's polling loop: the operand tile is released -- bar_stage -- a poll period earlier).  0: the issuer warp issues.
#ifndef GCS_TC_SELF_ISSUE
#define GCS_TC_SELF_ISSUE 0
#endif
// where the raw rows of the warp's next tile are loaded into registers: 0 behind the operand stores of stage 3, 1 ahead of
// stage 3 (the loads then have the B-operand packing to complete in)
// two MMA issuers per CTA (Q <= 3): see issuer_role
#ifndef GCS_TC_DUAL_ISSUE
#define GCS_TC_DUAL_ISSUE 0
#endif
// 1: wait for the operand tile's release only at the first operand store (behind the first bin group's arithmetic)
#ifndef GCS_TC_LATE_WAIT
#define GCS_TC_LATE_WAIT 0
#endif
#ifndef GCS_TC_FETCH_AT
#define GCS_TC_FETCH_AT 1
#endif
constexpr int kLoColH = 20;   // first "lo" feature column of the fp16 B operand (columns 19 and 39 are zero)

template <int Q, bool H>
struct TcCfg {
  // producer warps: as many 17 KB / 21 KB operand tiles as fit in shared memory, in whole warpgroups
  static constexpr int kProd = Q <= 3 ? 12 : 8;
#ifndef GCS_TC_REG_PROD
#define GCS_TC_REG_PROD 136
#define GCS_TC_REG_EPI 104
#endif
  static constexpr int kRegProd = Q <= 3 ? GCS_TC_REG_PROD : 184;   // setmaxnreg targets (launch allocation 128 per thread)
  static constexpr int kRegEpi = Q <= 3 ? GCS_TC_REG_EPI : 104;
  static constexpr int kBinsPad = 16 * Q;   // rows of the e_hi block == first row of the e_lo block
  static constexpr int kRows = 32 * Q;      // rows of A that carry data
  // Warpgroup after the producers: epilogue warps, one per 32 TMEM lanes that carry data (Q of them).  For Q = 3 its
  // fourth warp is the MMA issuer; for Q = 4 the issuer is the first warp of one more (otherwise idle) warpgroup.
  static constexpr int kEpi = Q;
  static constexpr int kIssuerWarp = Q <= 3 ? kProd + 3 : kProd + 4;
  static constexpr int kThreads = Q <= 3 ? 32 * (kProd + 4) : 32 * (kProd + 8);
  static constexpr int kTilePts = H ? 64 : 32;                        // points per operand tile
  static constexpr int kNBuf = 1;                                     // operand tiles per producer warp
  static constexpr int kMmaPerTile = 4;                               // K = 16 (f16, 64 points) / 8 (tf32, 32 points) per MMA
  static constexpr int kKgA = kRows * 16;                             // fp16: bytes between K groups of the A tile
  static constexpr int kABytes = H ? 8 * kKgA : kRows * tc::kRowBytes;
  static constexpr int kBBytes = H ? 8 * kKgB : kBRows * tc::kRowBytes;
  static constexpr int kStageBytes = kABytes + kBBytes;   // 17 / 21 KB in both forms (multiples of 1 KB)
  static constexpr int kStagesBytes = kProd * kNBuf * kStageBytes + 1024;   // + tail read by the last tile's padding rows
  static constexpr int kLoCol = H ? kLoColH : kNF;                    // first "lo" feature column of the accumulator
};

// float32 copy of the hoisted twist invariants, every scalar duplicated into a pair (one 8-byte load feeds an f32x2 operand)
struct TwistCtxF {
  float2 rho[3], phi[3], nphi[3], c1[3], c2[3], th1sq, nth1sq, norg[3];
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
  TwistCtxF twf[kMaxProd];
  WindowCtx win[kMaxProd];
  // per producer warp: base pointers of the current segment.  Kept here instead of in ten registers across the tile loop
  // (read back with one LDS.64 each where a tile needs them): the registers go to the soft-assign stage.
  struct SegPtrs { const double* pts; const double* tp; const double* wp; double* dkp; double* dkw; const uint8_t* rp; const uint8_t* gp; } seg[kMaxProd];
};

struct TcGeom {
  int64_t tiles_per_unit;   // ceil(cap / tile_pts)
  int64_t total_tiles;      // U * tiles_per_unit
  int n_cta, n_parts, flush;
  int dbg;   // 0, or 1 + slot of the timing hook
};

// Tiles of producer warp w's FIRST round in a segment; every later round has G.flush tiles.  The twelve producer warps run
// in lock-step (same start, same work per tile): with equal rounds they all hand their accumulators to the epilogue at the
// same moment, the drains serialise, and every warp waits for its turn at every round boundary (a tenth of the producers'
// stall samples sat on the operand-tile barrier behind it).  Staggered first rounds spread the drains evenly over a round;
// the first rounds grow with w, so the epilogue's fixed (round, warp) drain order is also the order of completion.
#ifndef GCS_TC_STAGGER
#define GCS_TC_STAGGER 0
#endif
__device__ __forceinline__ int first_round_tiles(const TcGeom& G, int w, int n_prod) {
  return GCS_TC_STAGGER ? 1 + (w * G.flush) / n_prod : G.flush;
}

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
    if (H) {   // e_lo rows carry the unscaled residual (the 2^-kShiftLo of the feature "lo" columns is applied by the epilogue)
      const double sc_f = ldexp(1.0, -kShiftE - (f < 10 ? kShiftD : (f < 13 ? kShiftP : kShiftPP)));
      part[b * kRowLen + f] = (red[b * kNF + f] + red[(C::kBinsPad + b) * kNF + f]) * sc_f;
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
//   [producers: tiles + MMAs | epilogue: drains]  B1  [epilogue: row sums -> red]  B2  [all: partial]  B3  ([all: re-zero red]  B4: tf32 only)
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
  if (H) return;   // fp16 tiles are rewritten completely by every tile; the rows that alias this scratch feed unread accumulator rows
  // padding rows of the operand tiles alias this scratch: keep it free of NaN bit patterns
  for (int k = tid; k < (128 * kNF * 8 + 15) / 16; k += C::kThreads) reinterpret_cast<uint4*>(stages)[k] = make_uint4(0, 0, 0, 0);
  tc::fence_smem_to_async();
  bar_all(C::kThreads);
}

// ---------------------------------------------------------------------------------------------------------------------
// Producer, tf32 operands (H = false): 32-point tiles, K-major SWIZZLE_128B, all-float64 geometry.  One lane = one point
// in stages 1 and 3; stage 2 uses a pair layout (lane = bin half x point pair).
// ---------------------------------------------------------------------------------------------------------------------
template <int Q>
__device__ __forceinline__ void producer_role_tf32(const BinScanParams& P, const TcGeom& G, TcMisc& mi, unsigned char* stages,
                                                   uint32_t tmem, int cta, int tid) {
  constexpr bool H = false;
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
      unsigned char* const sA = sA0;
      unsigned char* const sB = sA + C::kABytes;
      unsigned char* const sA_lane = sA + half * NB2 * tc::kRowBytes;   // the lane's bins are [half * NB2, half * NB2 + NB2)
      // the tensor core may still be reading the operand tile this one goes into
      if (n_stage_uses >= 1u) tc::mbar_wait(&mi.bar_stage[wid][0], (n_stage_uses - 1) & 1);
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
          hi[k] = make_float2(tc::tf32_hi(e[k].x), tc::tf32_hi(e[k].y));
          lo[k] = tc::sub2(e[k], hi[k]);
        }
#pragma unroll
        for (int k = 0; k < kGrp; ++k) {
          const int b = g * kGrp + k;
          *reinterpret_cast<float2*>(sA_lane + b * tc::kRowBytes + off2[b & 7]) = hi[k];
          *reinterpret_cast<float2*>(sA_lane + (C::kBinsPad + b) * tc::kRowBytes + off2[b & 7]) = lo[k];
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
        const float sc = row ? (float)(w_dk * inv) : 0.f;
        const float q0 = (float)p0[0], q1 = (float)p0[1], q2 = (float)p0[2];
        const float sd0 = sc * f0, sd1 = sc * f1, sd2 = sc * f2;
        const float sp0 = sc * q0, sp1 = sc * q1, sp2 = sc * q2;
        const float v[kNF] = {sc, sd0, sd1, sd2, sd0 * f0, sd0 * f1, sd0 * f2, sd1 * f1, sd1 * f2, sd2 * f2,
                              sp0, sp1, sp2, sp0 * q0, sp0 * q1, sp0 * q2, sp1 * q1, sp1 * q2, sp2 * q2};
#pragma unroll
        for (int f = 0; f < kNF; ++f) {
          float hi, lo;
          tc::split_tf32(v[f], hi, lo);
          *reinterpret_cast<float*>(sB + f * tc::kRowBytes + off[f & 7]) = hi;
          *reinterpret_cast<float*>(sB + (kNF + f) * tc::kRowBytes + off[(kNF + f) & 7]) = lo;
        }
      }
      tc::fence_smem_to_async();
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(&mi.bar_tile[wid][0]);   // the issuer warp takes it from here
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

// ---------------------------------------------------------------------------------------------------------------------
// Producer, fp16 operands (H = true): 64-point tiles, MN-major unswizzled operand tiles, two points per lane.
//   A tile: byte (K group kg, M block mb, row r, element) = kg * kKgA + mb * 128 + r * 16 + 2 * (m & 7)
//   B tile: the same with kKgB; feature columns [hi 0..18, 0, lo 0..18, 0]
// Lane l owns the points 2 l, 2 l + 1 of the tile and writes them to K rows l and 32 + l.
// ---------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ float2 bc2(float x) { return make_float2(x, x); }
__device__ __forceinline__ float max3f(float a, float b, float c) {
  float d;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
  return d;
}
__device__ __forceinline__ float rsqrt_approx(float x) {
  float y;
  asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float rcp_approx(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// 16-byte accesses of a pair of consecutive doubles
__device__ __forceinline__ double2 ldg_d2(const double* p) { return __ldg(reinterpret_cast<const double2*>(p)); }
__device__ __forceinline__ void stg_d2(double* p, double a, double b) { *reinterpret_cast<double2*>(p) = make_double2(a, b); }


template <int Q>
__device__ __forceinline__ void producer_role_mn(const BinScanParams& P, const TcGeom& G, TcMisc& mi, unsigned char* stages,
                                                 uint32_t tmem, int cta, int tid) {
  constexpr bool H = true;
  using C = TcCfg<Q, H>;
  constexpr int kProd = C::kProd;
  constexpr int kTile = C::kTilePts;
  const int wid = tid >> 5, lane = tid & 31;
  uint32_t n_stage_uses = 0;
#if GCS_TC_SELF_ISSUE
  // round bookkeeping of this warp's accumulator (what the issuer warp keeps per lane otherwise)
  uint32_t par_empty = 0;
  bool have_round = false;
#endif
  unsigned char* const sA = stages + wid * C::kStageBytes;
  unsigned char* const sB = sA + C::kABytes;
  // K rows lane (point a) and 32 + lane (point b): K group lane >> 3 (+ 4), row lane & 7 of the core matrices
  unsigned char* const sA_a = sA + (lane >> 3) * C::kKgA + (lane & 7) * 16;
  unsigned char* const sA_b = sA_a + 4 * C::kKgA;
  unsigned char* const sB_a = sB + (lane >> 3) * kKgB + (lane & 7) * 16;
  unsigned char* const sB_b = sB_a + 4 * kKgB;
  const TwistCtxF& tf = mi.twf[wid];
  int64_t g0 = cta_tile0(G, cta);
  const int64_t g_end = cta_tile0(G, cta + 1);
  TcSeg sg;
  while (next_segment(G, P.n_hyp, g0, g_end, sg)) {
    const int u = sg.u, s = sg.s, h = sg.h;
    const int64_t lt0 = sg.lt0, lt1 = sg.lt1;
    const double t0 = P.t0s[s], t1 = P.t1s[s];
    const double inv_denom = 1.0 / fmax(t1 - t0, 1e-12);
    __syncwarp();
    if (lane == 0) {
      mi.win[wid] = make_window_ctx(t0, t1);
      const TwistCtx c = make_twist_ctx(P.xi + (int64_t)u * 6);
      mi.tw[wid] = c;
      TwistCtxF& f = mi.twf[wid];
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        f.rho[k] = bc2((float)c.rho[k]); f.phi[k] = bc2((float)c.phi[k]); f.nphi[k] = bc2(-(float)c.phi[k]);
        f.c1[k] = bc2((float)c.c1[k]); f.c2[k] = bc2((float)c.c2[k]); f.norg[k] = bc2(-(float)P.origin[k]);
      }
      f.th1sq = bc2((float)c.th1sq); f.nth1sq = bc2(-(float)c.th1sq);
    }
    __syncwarp();
    const double mass_scale = P.mass[s * kNMass + kMassAll] / (P.mass[s * kNMass + kMassSel] + P.eps_mass);
    const double* pts = P.pts + (int64_t)s * P.n_raw * 3;
    const double* tp = P.t + (int64_t)s * P.n_raw;
    const double* wp = P.w + (int64_t)s * P.n_raw;
    // ring / tag only pass through to the resampled rows: not read at all when those are not materialised
    const uint8_t* rp = (P.ring && P.rs_pts) ? P.ring + (int64_t)s * P.n_raw : nullptr;
    const uint8_t* gp = (P.tag && P.rs_pts) ? P.tag + (int64_t)s * P.n_raw : nullptr;
    double* const dkp0 = P.dk_pts ? P.dk_pts + (int64_t)u * P.cap * 3 : nullptr;
    double* const dkw0 = P.dk_w ? P.dk_w + (int64_t)u * P.cap : nullptr;
    // 16-byte paths: consecutive raw rows (stride 1) and 16-byte aligned scan / unit bases (pair index is even)
    const bool vec_in = P.stride == 1 && (((uintptr_t)pts | (uintptr_t)tp | (uintptr_t)wp) & 15) == 0 &&
                        (!rp || ((uintptr_t)rp & 1) == 0) && (!gp || ((uintptr_t)gp & 1) == 0);
    const bool vec_out = (((uintptr_t)dkp0 | (uintptr_t)dkw0) & 15) == 0;
    __syncwarp();
    if (lane == 0) {
      mi.seg[wid].pts = pts; mi.seg[wid].tp = tp; mi.seg[wid].wp = wp; mi.seg[wid].dkp = dkp0; mi.seg[wid].dkw = dkw0;
      mi.seg[wid].rp = rp; mi.seg[wid].gp = gp;
    }
    __syncwarp();
    const bool has_rt = rp != nullptr || gp != nullptr;
    const volatile TcMisc::SegPtrs& sp = mi.seg[wid];
    double ent_dot = 0.0, ent_log = 0.0, sum_wdk = 0.0, sum_wrs = 0.0;
    float mx_resp = 0.f;
    int n_rows = 0;

    // software pipeline: the raw rows of the warp's next tile are requested before this tile is processed
    double nx[10];
    // ring / tag of the pair as loaded (ring a | ring b << 8, tag a | tag b << 8): kept in registers of their own and only
    // combined where they are used -- any arithmetic on a prefetched value at the prefetch site waits for the load there
    // (8 % of all stall samples sat on the shift that used to merge them)
    unsigned short nring = 0, ntag = 0;
    auto fetch = [&](int64_t tile) {
      const int64_t ia = tile * kTile + 2 * lane;
#pragma unroll
      for (int k = 0; k < 10; ++k) nx[k] = 0.0;
      nring = 0; ntag = 0;
      if (tile >= lt1) return;
      if (vec_in && ia + 1 < P.n_sel) {
        const double* pts_ = sp.pts;
        const double* tp_ = sp.tp;
        const double* wp_ = sp.wp;
        const double2 a0 = ldg_d2(pts_ + 3 * ia), a1 = ldg_d2(pts_ + 3 * ia + 2), a2 = ldg_d2(pts_ + 3 * ia + 4);
        const double2 tt = ldg_d2(tp_ + ia), ww = ldg_d2(wp_ + ia);
        nx[0] = a0.x; nx[1] = a0.y; nx[2] = a1.x; nx[3] = tt.x; nx[4] = ww.x;
        nx[5] = a1.y; nx[6] = a2.x; nx[7] = a2.y; nx[8] = tt.y; nx[9] = ww.y;
        if (has_rt) {
          const uint8_t* rp_ = sp.rp;
          const uint8_t* gp_ = sp.gp;
          if (rp_) nring = __ldg(reinterpret_cast<const unsigned short*>(rp_ + ia));
          if (gp_) ntag = __ldg(reinterpret_cast<const unsigned short*>(gp_ + ia));
        }
      } else {
        const double* pts_ = sp.pts;
        const double* tp_ = sp.tp;
        const double* wp_ = sp.wp;
        const uint8_t* rp_ = sp.rp;
        const uint8_t* gp_ = sp.gp;
#pragma unroll
        for (int q = 0; q < 2; ++q) {
          if (ia + q < P.n_sel) {
            const int64_t j = (ia + q) * P.stride;
            nx[5 * q] = pts_[3 * j]; nx[5 * q + 1] = pts_[3 * j + 1]; nx[5 * q + 2] = pts_[3 * j + 2];
            nx[5 * q + 3] = tp_[j]; nx[5 * q + 4] = wp_[j];
            if (rp_) nring |= (unsigned short)((unsigned)rp_[j] << (8 * q));
            if (gp_) ntag |= (unsigned short)((unsigned)gp_[j] << (8 * q));
          }
        }
      }
    };
    // The raw rows of the warp's next tile are pulled into L2 at the top of a tile (no registers held) and loaded into
    // registers only behind the operand stores of this tile, where the register-hungry soft-assign stage is over: with
    // the loads at the top, 20 registers stayed live across the whole tile and loop invariants were spilled.
    auto prefetch_l2 = [&](int64_t tile) {
      if (tile >= lt1 || !vec_in) return;
      const int64_t ia = tile * kTile + 2 * lane;
      if (ia + 1 >= P.n_sel) return;
      const double* pts_ = sp.pts;
      const double* tp_ = sp.tp;
      const double* wp_ = sp.wp;
      asm volatile("prefetch.global.L2 [%0];" ::"l"(pts_ + 3 * ia));
      asm volatile("prefetch.global.L2 [%0];" ::"l"(pts_ + 3 * ia + 4));
      asm volatile("prefetch.global.L2 [%0];" ::"l"(tp_ + ia));
      asm volatile("prefetch.global.L2 [%0];" ::"l"(wp_ + ia));
    };
    fetch(lt0 + wid);
#if GCS_TC_SELF_ISSUE
    int in_round = 0, round_len = first_round_tiles(G, wid, kProd);
#endif
    for (int64_t lt = lt0 + wid; lt < lt1; lt += kProd) {
#if defined(GCS_TC_DRY) && !GCS_TC_SELF_ISSUE
      // experiment: the hand-off protocol alone (no operand work) -- the rate the issuer warp sustains
      if (n_stage_uses >= 1u) tc::mbar_wait(&mi.bar_stage[wid][0], (n_stage_uses - 1) & 1);
      tc::fence_smem_to_async();
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(&mi.bar_tile[wid][0]);
      ++n_stage_uses;
      continue;
#endif
      // ================= stage 1: resample gather, sweep fraction + window weight (float64), deskew increment (float32)
      const int64_t ia = lt * kTile + 2 * lane;
      const bool row_a = ia < P.cap, row_b = ia + 1 < P.cap;
      const double pa[3] = {nx[0], nx[1], nx[2]}, pb[3] = {nx[5], nx[6], nx[7]};
      const double tta = nx[3], ttb = nx[8];
      const double w_rs_a = nx[4] * mass_scale, w_rs_b = nx[9] * mass_scale;
      const unsigned short ring2 = nring, tag2 = ntag;
      prefetch_l2(lt + kProd);
      if (h == 0 && P.rs_pts && row_a) {
        const int64_t o = (int64_t)s * P.cap + ia;
        P.rs_pts[3 * o] = pa[0]; P.rs_pts[3 * o + 1] = pa[1]; P.rs_pts[3 * o + 2] = pa[2];
        P.rs_t[o] = tta; P.rs_w[o] = w_rs_a; P.rs_ring[o] = (uint8_t)ring2; P.rs_tag[o] = (uint8_t)tag2;
        if (row_b) {
          P.rs_pts[3 * o + 3] = pb[0]; P.rs_pts[3 * o + 4] = pb[1]; P.rs_pts[3 * o + 5] = pb[2];
          P.rs_t[o + 1] = ttb; P.rs_w[o + 1] = w_rs_b; P.rs_ring[o + 1] = (uint8_t)(ring2 >> 8); P.rs_tag[o + 1] = (uint8_t)(tag2 >> 8);
        }
      }
      const double dta = tta - t0, dtb = ttb - t0;   // float64: epoch stamps (1.67e9 s) lose the sweep to float32
      const double alpha_a = dta * inv_denom, alpha_b = dtb * inv_denom;
      // Window weight sigmoid(a) sigmoid(S - a) = x / (c + x (1 + c + x)), x = e^-a = 2^(n + f): the exponent argument is
      // reduced in float64 (n = rint, |f| <= 1/2), 2^f, the quotient and its Newton step in float32 -- relative error
      // 3e-7 on an operator output whose stated tolerance is 1e-5 -- and the result returns to float64 for the product with
      // the (float64) resampled weight.  The all-float64 form (exp2 polynomial + reciprocal: 30 FP64-pipe instructions per
      // point on a pipe a quarter as wide) was a tenth of the producer's stall samples.
      double w_dk_a, w_dk_b;
      {
        const WindowCtx& wc = mi.win[wid];
        const double magic = 6755399441055744.0;   // 1.5 * 2^52: x + magic - magic = rint(x)
        const double ea = fmin(fmax(dta * wc.inv_sig * -1.4426950408889634, -120.0), 60.0);
        const double eb = fmin(fmax(dtb * wc.inv_sig * -1.4426950408889634, -120.0), 60.0);
        const double na = (ea + magic) - magic, nb = (eb + magic) - magic;
        const float2 fr = make_float2((float)(ea - na), (float)(eb - nb));
        float2 x = make_float2(tc::ex2f(fr.x), tc::ex2f(fr.y));
        // scale by 2^n through the exponent field (n in [-120, 60]: x stays a normal float32, x^2 stays finite)
        x.x = __int_as_float(__float_as_int(x.x) + ((int)na << 23));
        x.y = __int_as_float(__float_as_int(x.y) + ((int)nb << 23));
        const float2 cf = bc2((float)wc.c), opc = bc2((float)wc.one_plus_c);
        const float2 den = tc::fma2(x, tc::add2(opc, x), cf);
        float2 r = make_float2(rcp_approx(den.x), rcp_approx(den.y));
        r = tc::fma2(r, tc::fma2(tc::mul2(den, bc2(-1.0f)), r, bc2(1.0f)), r);
        const float2 ww = tc::mul2(x, r);
        w_dk_a = w_rs_a * fma((double)ww.x, 1.0 - kWeightFloor, kWeightFloor);
        w_dk_b = w_rs_b * fma((double)ww.y, 1.0 - kWeightFloor, kWeightFloor);
      }
      const float2 a = make_float2((float)alpha_a, (float)alpha_b);
      const float2 a2 = tc::mul2(a, a);
      const float2 th2 = tc::mul2(a2, tf.th1sq);
      double p0a[3], p0b[3];
      float2 q[3];   // deskewed point, float32, (point a, point b) per component
      if (th2.x < 0.25f && th2.y < 0.25f) {
        // Maclaurin series of sin(th)/th, (1-cos th)/th^2, (th-sin th)/th^3 in th^2 (truncation < 3e-9 for th^2 < 0.25)
        float2 S = tc::fma2(th2, bc2(2.7557319e-06f), bc2(-1.9841270e-04f));
        float2 Cc = tc::fma2(th2, bc2(-2.4801587e-05f), bc2(1.3888889e-03f));
        float2 Gg = tc::fma2(th2, bc2(-2.7557319e-06f), bc2(1.9841270e-04f));
        S = tc::fma2(S, th2, bc2(8.3333333e-03f));  Cc = tc::fma2(Cc, th2, bc2(-4.1666667e-02f)); Gg = tc::fma2(Gg, th2, bc2(-8.3333333e-03f));
        S = tc::fma2(S, th2, bc2(-1.6666667e-01f)); Cc = tc::fma2(Cc, th2, bc2(0.5f));            Gg = tc::fma2(Gg, th2, bc2(1.6666667e-01f));
        S = tc::fma2(S, th2, bc2(1.0f));
        const float2 aG = tc::mul2(a, Gg);
        const float2 na = tc::mul2(a, bc2(-1.0f));
        const float2 pf[3] = {make_float2((float)pa[0], (float)pb[0]), make_float2((float)pa[1], (float)pb[1]),
                              make_float2((float)pa[2], (float)pb[2])};
        float2 d1[3];
#pragma unroll
        for (int k = 0; k < 3; ++k) {
          const float2 uu = tc::fma2(aG, tf.c2[k], tc::mul2(Cc, tf.c1[k]));
          d1[k] = tc::mul2(na, tc::fma2(a, uu, tf.rho[k]));     // -t(a)
          q[k] = tc::add2(pf[k], d1[k]);
        }
        const float2 x0 = tc::fma2(tf.phi[1], q[2], tc::mul2(tf.nphi[2], q[1]));
        const float2 x1 = tc::fma2(tf.phi[2], q[0], tc::mul2(tf.nphi[0], q[2]));
        const float2 x2 = tc::fma2(tf.phi[0], q[1], tc::mul2(tf.nphi[1], q[0]));
        const float2 pq = tc::fma2(tf.phi[0], q[0], tc::fma2(tf.phi[1], q[1], tc::mul2(tf.phi[2], q[2])));
        const float2 nsa = tc::mul2(S, na), ca = tc::mul2(Cc, a2);
        const float2 xs[3] = {x0, x1, x2};
#pragma unroll
        for (int k = 0; k < 3; ++k) {
          const float2 yy = tc::fma2(tf.phi[k], pq, tc::mul2(tf.nth1sq, q[k]));
          const float2 rot = tc::fma2(ca, yy, tc::mul2(nsa, xs[k]));
          const float2 dl = tc::add2(d1[k], rot);               // p0 - p, float32
          q[k] = tc::add2(q[k], rot);
          p0a[k] = pa[k] + (double)dl.x; p0b[k] = pb[k] + (double)dl.y;
        }
      } else {   // large angles (garbage sweep fractions of zero-stamped padded rows): float64 closed forms
        deskew_point_ctx(pa, alpha_a, mi.tw[wid], p0a);
        deskew_point_ctx(pb, alpha_b, mi.tw[wid], p0b);
#pragma unroll
        for (int k = 0; k < 3; ++k) q[k] = make_float2((float)p0a[k], (float)p0b[k]);
      }
      double* const dkp = sp.dkp;
      double* const dkw = sp.dkw;
      if (row_b && vec_out) {
        if (dkp) {
          stg_d2(dkp + 3 * ia, p0a[0], p0a[1]); stg_d2(dkp + 3 * ia + 2, p0a[2], p0b[0]); stg_d2(dkp + 3 * ia + 4, p0b[1], p0b[2]);
        }
        if (dkw) stg_d2(dkw + ia, w_dk_a, w_dk_b);
      } else if (row_a) {
        if (dkp) {
          dkp[3 * ia] = p0a[0]; dkp[3 * ia + 1] = p0a[1]; dkp[3 * ia + 2] = p0a[2];
          if (row_b) { dkp[3 * ia + 3] = p0b[0]; dkp[3 * ia + 4] = p0b[1]; dkp[3 * ia + 5] = p0b[2]; }
        }
        if (dkw) { dkw[ia] = w_dk_a; if (row_b) dkw[ia + 1] = w_dk_b; }
      }
      if (row_a) { sum_wdk += w_dk_a; sum_wrs += w_rs_a; ++n_rows; }
      if (row_b) { sum_wdk += w_dk_b; sum_wrs += w_rs_b; ++n_rows; }
      // ray direction d = r / (|r| + eps), r = p0 - origin
      float2 f[3];
      {
        const float2 r0 = tc::add2(q[0], tf.norg[0]), r1 = tc::add2(q[1], tf.norg[1]), r2 = tc::add2(q[2], tf.norg[2]);
        const float2 rr = tc::fma2(r0, r0, tc::fma2(r1, r1, tc::mul2(r2, r2)));
        float2 y = make_float2(rsqrt_approx(fmaxf(rr.x, 1e-36f)), rsqrt_approx(fmaxf(rr.y, 1e-36f)));
        const float2 nh = tc::mul2(rr, bc2(-0.5f));
        y = tc::mul2(y, tc::fma2(nh, tc::mul2(y, y), bc2(1.5f)));                 // one Newton step
        const float2 invn = tc::fma2(tc::mul2(y, bc2(-(float)P.eps_mass)), y, y);   // 1 / (|r| + eps) to first order
        f[0] = tc::mul2(r0, invn); f[1] = tc::mul2(r1, invn); f[2] = tc::mul2(r2, invn);
      }

      // ================= stage 2: A = [e_hi ; e_lo] for both points; bins in pairs (b, b + 1) per f32x2 register
      const float2 ga0 = bc2(f[0].x), ga1 = bc2(f[1].x), ga2 = bc2(f[2].x);
      const float2 gb0 = bc2(f[0].y), gb1 = bc2(f[1].y), gb2 = bc2(f[2].y);
      // the tensor core may still be reading this warp's operand tile
#if !GCS_TC_LATE_WAIT
      if (n_stage_uses >= 1u) tc::mbar_wait(&mi.bar_stage[wid][0], (n_stage_uses - 1) & 1);
#endif
      float2 sum_a = make_float2(0.f, 0.f), dot_a = sum_a, sum_b = sum_a, dot_b = sum_a;
      float mx_a = 0.f, mx_b = 0.f;
      constexpr int kGroups = C::kBinsPad / 8;   // 8 bins = 4 bin pairs = one 16-byte chunk of hi and one of lo per point
      // The bin table of a group (4 bin pairs, 32 registers) is loaded at the head of the group.  A second buffer that
      // prefetched the next group cost 32 more registers: loop invariants were spilled, and their reloads -- local
      // memory behind an L1 that streams the point data -- were a tenth of the kernel's stall samples.
#pragma unroll
      for (int g = 0; g < kGroups; ++g) {
#if GCS_TC_FETCH_AT >= 2
        if (g == kGroups - (GCS_TC_FETCH_AT - 1)) fetch(lt + kProd);
#endif
        float4 tab[4][2];
#pragma unroll
        for (int k = 0; k < 4; ++k) { tab[k][0] = mi.bins2[2 * (4 * g + k)]; tab[k][1] = mi.bins2[2 * (4 * g + k) + 1]; }
        uint32_t ha[4], la[4], hb[4], lb[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const float4 ta = tab[k][0], tb = tab[k][1];
          const float2 mx = make_float2(ta.x, ta.y), my = make_float2(ta.z, ta.w), mz = make_float2(tb.x, tb.y),
                       cc = make_float2(tb.z, tb.w);
          const float2 l_a = tc::fma2(ga0, mx, tc::fma2(ga1, my, tc::fma2(ga2, mz, cc)));
          const float2 l_b = tc::fma2(gb0, mx, tc::fma2(gb1, my, tc::fma2(gb2, mz, cc)));
          const float2 e_a = make_float2(tc::ex2f(l_a.x), tc::ex2f(l_a.y));
          const float2 e_b = make_float2(tc::ex2f(l_b.x), tc::ex2f(l_b.y));
          sum_a = tc::add2(sum_a, e_a); dot_a = tc::fma2(e_a, l_a, dot_a); mx_a = max3f(mx_a, e_a.x, e_a.y);
          sum_b = tc::add2(sum_b, e_b); dot_b = tc::fma2(e_b, l_b, dot_b); mx_b = max3f(mx_b, e_b.x, e_b.y);
          // e' = hi + lo, both fp16: hi = rn(e'), lo = rn(e' - hi) (exact residual in float32)
          ha[k] = tc::pack_f16x2(e_a.x, e_a.y);
          hb[k] = tc::pack_f16x2(e_b.x, e_b.y);
          const float2 r_a = tc::sub2(e_a, tc::unpack_f16x2(ha[k]));
          const float2 r_b = tc::sub2(e_b, tc::unpack_f16x2(hb[k]));
          la[k] = tc::pack_f16x2(r_a.x, r_a.y);
          lb[k] = tc::pack_f16x2(r_b.x, r_b.y);
        }
#if GCS_TC_LATE_WAIT
        if (g == 0 && n_stage_uses >= 1u) tc::mbar_wait(&mi.bar_stage[wid][0], (n_stage_uses - 1) & 1);
#endif
        // M block g holds e_hi of bins 8g..8g+7, M block kBinsPad/8 + g their e_lo
        *reinterpret_cast<uint4*>(sA_a + g * 128) = make_uint4(ha[0], ha[1], ha[2], ha[3]);
        *reinterpret_cast<uint4*>(sA_a + (kGroups + g) * 128) = make_uint4(la[0], la[1], la[2], la[3]);
        *reinterpret_cast<uint4*>(sA_b + g * 128) = make_uint4(hb[0], hb[1], hb[2], hb[3]);
        *reinterpret_cast<uint4*>(sA_b + (kGroups + g) * 128) = make_uint4(lb[0], lb[1], lb[2], lb[3]);
      }

#if GCS_TC_FETCH_AT == 1
      fetch(lt + kProd);
#endif
      // ================= stage 3: B = [phi_hi, 0, phi_lo, 0], phi = (w / Z) (1, d, d d^T, p, p p^T)
      const float2 ssum = make_float2(sum_a.x + sum_a.y, sum_b.x + sum_b.y);
      float2 inv = make_float2(rcp_approx(ssum.x), rcp_approx(ssum.y));
      inv = tc::fma2(inv, tc::fma2(tc::mul2(ssum, bc2(-1.0f)), inv, bc2(1.0f)), inv);   // one Newton step: 1 / (2^kShiftE Z)
      {
        const float da_ = (dot_a.x + dot_a.y) * inv.x, db_ = (dot_b.x + dot_b.y) * inv.y;
        const float lg_a = __log2f(ssum.x), lg_b = __log2f(ssum.y);
        ent_dot += (double)((row_a ? da_ : 0.f) + (row_b ? db_ : 0.f));
        ent_log += (double)((row_a ? lg_a : 0.f) + (row_b ? lg_b : 0.f));
        mx_resp = max3f(mx_resp, row_a ? mx_a * inv.x : 0.f, row_b ? mx_b * inv.y : 0.f);
      }
      {
        const float2 sc0 = tc::mul2(make_float2(row_a ? (float)w_dk_a : 0.f, row_b ? (float)w_dk_b : 0.f), inv);
        const float2 sc = tc::mul2(sc0, bc2((float)(1 << (kShiftE + kShiftD))));
        const float2 scp = tc::mul2(sc0, bc2((float)(1 << (kShiftE + kShiftP))));
        const float2 scpp = tc::mul2(sc0, bc2((float)(1 << (kShiftE + kShiftPP))));
        const float2 sd0 = tc::mul2(sc, f[0]), sd1 = tc::mul2(sc, f[1]), sd2 = tc::mul2(sc, f[2]);
        const float2 sq0 = tc::mul2(scpp, q[0]), sq1 = tc::mul2(scpp, q[1]), sq2 = tc::mul2(scpp, q[2]);
        // 20 columns: the 19 features and a zero
        const float2 v[20] = {sc, sd0, sd1, sd2, tc::mul2(sd0, f[0]), tc::mul2(sd0, f[1]), tc::mul2(sd0, f[2]),
                              tc::mul2(sd1, f[1]), tc::mul2(sd1, f[2]), tc::mul2(sd2, f[2]),
                              tc::mul2(scp, q[0]), tc::mul2(scp, q[1]), tc::mul2(scp, q[2]),
                              tc::mul2(sq0, q[0]), tc::mul2(sq0, q[1]), tc::mul2(sq0, q[2]),
                              tc::mul2(sq1, q[1]), tc::mul2(sq1, q[2]), tc::mul2(sq2, q[2]), make_float2(0.f, 0.f)};
        uint32_t ua[20], ub[20];   // 10 hi words + 10 lo words per point (two features per word)
#pragma unroll
        for (int j = 0; j < 10; ++j) {
          ua[j] = tc::pack_f16x2(v[2 * j].x, v[2 * j + 1].x);
          ub[j] = tc::pack_f16x2(v[2 * j].y, v[2 * j + 1].y);
          const float2 h_a = tc::unpack_f16x2(ua[j]), h_b = tc::unpack_f16x2(ub[j]);
          const float2 r0 = tc::mul2(tc::sub2(v[2 * j], make_float2(h_a.x, h_b.x)), bc2((float)(1 << kShiftLo)));
          const float2 r1 = tc::mul2(tc::sub2(v[2 * j + 1], make_float2(h_a.y, h_b.y)), bc2((float)(1 << kShiftLo)));
          ua[10 + j] = tc::pack_f16x2(r0.x, r1.x);
          ub[10 + j] = tc::pack_f16x2(r0.y, r1.y);
        }
#pragma unroll
        for (int c = 0; c < 5; ++c) {
          *reinterpret_cast<uint4*>(sB_a + c * 128) = make_uint4(ua[4 * c], ua[4 * c + 1], ua[4 * c + 2], ua[4 * c + 3]);
          *reinterpret_cast<uint4*>(sB_b + c * 128) = make_uint4(ub[4 * c], ub[4 * c + 1], ub[4 * c + 2], ub[4 * c + 3]);
        }
      }
#if !GCS_TC_FETCH_AT
      fetch(lt + kProd);
#endif
      tc::fence_smem_to_async();
      __syncwarp();
#if GCS_TC_SELF_ISSUE
      {
        // first tile of a round overwrites the accumulator: the epilogue must have drained the previous round
        const bool last = (in_round + 1 == round_len) || (lt + kProd >= lt1);
        if (in_round == 0 && have_round) tc::mbar_wait(&mi.bar_empty[wid], par_empty);
        if (lane == 0) {
          tc::fence_after_sync();
          const uint32_t idesc = tc::idesc_f16(128, kMmaN) | tc::kIdescMnMajorA | tc::kIdescMnMajorB;
          const uint32_t sa = tc::smem_u32(sA);
          const uint64_t da = tc::smem_desc_mn(sa, C::kKgA, 128);
          const uint64_t db = tc::smem_desc_mn(sa + C::kABytes, kKgB, 128);
          const uint32_t d_tmem = tmem + wid * kAccStride;
#pragma unroll
          for (int ks = 0; ks < C::kMmaPerTile; ++ks)
            tc::mma_f16_ss(d_tmem, da + (uint64_t)(ks * ((2 * C::kKgA) >> 4)), db + (uint64_t)(ks * ((2 * kKgB) >> 4)), idesc,
                           (ks || in_round) ? 1u : 0u);
          tc::mma_commit(&mi.bar_stage[wid][0]);
          if (last) tc::mma_commit(&mi.bar_full[wid]);
        }
        if (last) { in_round = 0; round_len = G.flush; if (have_round) par_empty ^= 1u; have_round = true; } else ++in_round;
      }
#else
      if (lane == 0) tc::mbar_arrive(&mi.bar_tile[wid][0]);   // the issuer warp takes it from here
#endif
      ++n_stage_uses;
    }
    // per-warp sums of the scalar certificates (fixed shuffle tree)
    double v;
    v = warp_sum(ent_dot * kLn2); if (lane == 0) mi.ex[wid][kExEntDot] = v;
    v = warp_sum(ent_log * kLn2); if (lane == 0) mi.ex[wid][kExEntLog] = v;
    v = warp_sum(sum_wdk); if (lane == 0) mi.ex[wid][kExSumWdk] = v;
    v = warp_sum(sum_wrs); if (lane == 0) mi.ex[wid][kExSumWrs] = v;
    v = warp_sum((double)n_rows);  if (lane == 0) mi.ex[wid][kExCount] = v;
    v = warp_max((double)mx_resp); if (lane == 0) mi.ex[wid][5] = v;
    segment_tail<Q, H>(P, G, sg, mi, stages, cta, tid, nullptr);
  }
}

template <int Q, bool H>
__device__ __forceinline__ void producer_role(const BinScanParams& P, const TcGeom& G, TcMisc& mi, unsigned char* stages,
                                              uint32_t tmem, int cta, int tid) {
  if (H) producer_role_mn<Q>(P, G, mi, stages, tmem, cta, tid);
  else producer_role_tf32<Q>(P, G, mi, stages, tmem, cta, tid);
}

// Rounds (accumulator hand-overs to the epilogue) of producer warp w in a segment of n_seg tiles.
__device__ __forceinline__ int rounds_of(const TcGeom& G, int n_seg, int w, int n_prod) {
  const int tiles_w = n_seg > w ? (n_seg - w + n_prod - 1) / n_prod : 0;
  const int first = first_round_tiles(G, w, n_prod);
  return tiles_w <= first ? (tiles_w > 0 ? 1 : 0) : 1 + (tiles_w - first + G.flush - 1) / G.flush;
}

// One accumulator drain of an epilogue warp: TMEM lanes lane_base.., accumulator of producer warp w, into acc (float64).
template <int Q, bool H>
__device__ __forceinline__ void drain_accumulator(TcMisc& mi, uint32_t tmem, uint32_t lane_base, int w, int lane, double* acc) {
  tc::fence_after_sync();
  uint32_t a0[16], a1[16], a2[4], a3[4];
  const uint32_t addr = tmem + lane_base + w * kAccStride;
  tc::tmem_ld_x16(addr, a0);
  tc::tmem_ld_x16(addr + 16, a1);
  tc::tmem_ld_x4(addr + 32, a2);
  tc::tmem_ld_x4(addr + 36, a3);
  tc::tmem_ld_wait();
  tc::fence_before_sync();
  __syncwarp();
  if (lane == 0) tc::mbar_arrive(&mi.bar_empty[w]);
  // column f holds row x phi_hi[f], column kLoCol + f row x phi_lo[f] (fp16 operands: 2^-11 of the former): one
  // float32 add, then the float64 accumulation
  float v[40];
#pragma unroll
  for (int c = 0; c < 16; ++c) { v[c] = __uint_as_float(a0[c]); v[16 + c] = __uint_as_float(a1[c]); }
#pragma unroll
  for (int c = 0; c < 4; ++c) { v[32 + c] = __uint_as_float(a2[c]); v[36 + c] = __uint_as_float(a3[c]); }
  constexpr int kLo = TcCfg<Q, H>::kLoCol;
#pragma unroll
  for (int f = 0; f < kNF; ++f)
    acc[f] += (double)(H ? fmaf(v[kLo + f], 1.0f / (float)(1 << kShiftLo), v[f]) : v[f] + v[kLo + f]);
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
    for (int w = 0; w < kProd; ++w) rounds[w] = rounds_of(G, n_seg, w, kProd);
    for (int r = 0;; ++r) {
      bool any = false;
#pragma unroll
      for (int w = 0; w < kProd; ++w) {
        if (r < rounds[w]) {
          any = true;
          tc::mbar_wait(&mi.bar_full[w], n_drained[w] & 1);
          ++n_drained[w];
          drain_accumulator<Q, H>(mi, tmem, lane_base, w, lane, acc);
        }
      }
      if (!any) break;
    }
    segment_tail<Q, H>(P, G, sg, mi, stages, cta, tid, acc);
  }
}

// MMA issuer for producer warps [w_lo, w_hi): producers never block on the tensor-core queue.  Lane l (< w_hi - w_lo)
// watches the barriers of producer warp w_lo + l -- one `mbarrier.test_wait` instruction polls them all -- and one elected
// lane issues whichever tiles are ready, no fixed order across warps, so a late warp does not hold up the MMAs (and hence
// the operand-tile release) of the others.  Each accumulator still sees only its own warp's tiles, in sequence, and the
// epilogue drains in a fixed order: results do not depend on the issue order.
//
// The hand-off is a serial resource (about 0.2 us per tile: four tcgen05.mma, one or two tcgen05.commit, the poll), so the
// producers are split between TWO issuers: the spare warp of the epilogue warpgroup and -- `kDrain` -- the first epilogue
// warp, which drains its TMEM lanes from the same polling loop (the same (round, warp) order as the blocking epilogue warps).
template <int Q, bool H, bool kDrain>
__device__ __forceinline__ void issuer_role(const BinScanParams& P, const TcGeom& G, TcMisc& mi, unsigned char* stages,
                                            uint32_t tmem, int cta, int tid, int w_lo, int w_hi) {
  using C = TcCfg<Q, H>;
  constexpr int kProd = C::kProd;
  const int lane = tid & 31;
  const int n_own = w_hi - w_lo;
  const int my_w = w_lo + (lane < n_own ? lane : 0);
  int64_t g0 = cta_tile0(G, cta);
  const int64_t g_end = cta_tile0(G, cta + 1);
  const uint32_t idesc = H ? (tc::idesc_f16(128, kMmaN) | tc::kIdescMnMajorA | tc::kIdescMnMajorB) : tc::idesc_tf32(128, kMmaN);
  const uint32_t a0 = tc::smem_u32(stages);
  const uint32_t tmem_u = __reduce_max_sync(0xffffffffu, tmem);   // the same value in every lane, now known to be uniform
  const uint32_t lane_base = kDrain ? (uint32_t)(32 * ((tid >> 5) - kProd)) << 16 : 0u;
  TcSeg sg;
  uint32_t n_done = 0;      // tiles of this lane's warp issued so far: operand buffer n_done % kNBuf, phase n_done / kNBuf
  uint32_t par_empty = 0;   // phase of bar_empty to test next (the warp's previous round)
  bool have_round = false;  // a round of this lane's warp has been handed to the epilogue
  uint32_t par_full = 0;    // kDrain: bit w = phase of bar_full[w] to wait for next
  while (next_segment(G, P.n_hyp, g0, g_end, sg)) {
    double acc[kNF];
#pragma unroll
    for (int c = 0; c < kNF; ++c) acc[c] = 0.0;
    {
      const int n_seg = (int)(sg.lt1 - sg.lt0);
      const int my_tiles = (lane < n_own && n_seg > my_w) ? (n_seg - my_w + kProd - 1) / kProd : 0;
      int t = 0, in_round = 0, round_len = first_round_tiles(G, my_w, kProd);
      // drain cursor (warp-uniform): next (round, warp) in the epilogue's fixed order
      int dr = 0, dw = 0, max_rounds = 0;
      bool drain_valid = false;
      if (kDrain) {
        for (int w = 0; w < kProd; ++w) max_rounds = max(max_rounds, rounds_of(G, n_seg, w, kProd));
        drain_valid = max_rounds > 0;
        while (drain_valid && dr >= rounds_of(G, n_seg, dw, kProd)) {
          if (++dw == kProd) { dw = 0; if (++dr >= max_rounds) drain_valid = false; }
        }
      }
      for (;;) {
        const bool pending = t < my_tiles;
        bool ready = false;
        if (pending) {
          ready = tc::mbar_test_wait(&mi.bar_tile[my_w][0], n_done & 1u);
          if (ready && in_round == 0 && have_round) ready = tc::mbar_test_wait(&mi.bar_empty[my_w], par_empty);
        }
        const unsigned rdy = __ballot_sync(0xffffffffu, ready);
        bool drained = false;
        if (kDrain && drain_valid) {
          const bool ok = tc::mbar_test_wait(&mi.bar_full[dw], (par_full >> dw) & 1u);
          if (__ballot_sync(0xffffffffu, ok) & 1u) {   // lane 0's view, the same for every lane
            drain_accumulator<Q, H>(mi, tmem, lane_base, dw, lane, acc);
            par_full ^= 1u << dw;
            drained = true;
            do {
              if (++dw == kProd) { dw = 0; if (++dr >= max_rounds) { drain_valid = false; break; } }
            } while (dr >= rounds_of(G, n_seg, dw, kProd));
          }
        }
        if (rdy == 0u) {
          if (drained) continue;
          if (__ballot_sync(0xffffffffu, pending) == 0u && !(kDrain && drain_valid)) break;
          if (kIssuerSleepNs) __nanosleep(kIssuerSleepNs);   // a producer needs microseconds per tile: do not spend its issue slots on polling
          continue;
        }
        const bool last = (in_round + 1 == round_len) || (t + 1 >= my_tiles);
        const unsigned first_m = __ballot_sync(0xffffffffu, ready && in_round == 0);
        const unsigned last_m = __ballot_sync(0xffffffffu, ready && last);
        tc::fence_after_sync();
        {
          // every lane walks the (warp-uniform) ready mask; one elected lane issues.  All operands are warp-uniform.
          unsigned m = rdy;
          while (m) {
            const int l = __ffs(m) - 1;
            m &= m - 1;
            const int w = w_lo + l;
            const uint32_t sa = a0 + w * C::kStageBytes;
            const uint64_t da = H ? tc::smem_desc_mn(sa, C::kKgA, 128) : tc::smem_desc_sw128(sa);
            const uint64_t db = H ? tc::smem_desc_mn(sa + C::kABytes, kKgB, 128) : tc::smem_desc_sw128(sa + C::kABytes);
            const uint32_t d_tmem = tmem_u + w * kAccStride;
            const uint32_t acc0 = ((first_m >> l) & 1u) ^ 1u;
            const bool fin = (last_m >> l) & 1u;
            // tf32: one MMA consumes 32 bytes of every operand row (8 points): descriptor start + 2 (x 16 B);
            // fp16: one MMA consumes two K groups (16 points): descriptor start + 2 K-group strides
#ifdef GCS_TC_DRY_MMAS
            constexpr int kIssue = GCS_TC_DRY_MMAS;
#else
            constexpr int kIssue = C::kMmaPerTile;
#endif
            if (tc::elect_one()) {
#pragma unroll
              for (int ks = 0; ks < kIssue; ++ks) {
                if (H) tc::mma_f16_ss(d_tmem, da + (uint64_t)(ks * ((2 * C::kKgA) >> 4)), db + (uint64_t)(ks * ((2 * kKgB) >> 4)), idesc,
                                      ks ? 1u : acc0);
                else tc::mma_tf32_ss(d_tmem, da + 2 * ks, db + 2 * ks, idesc, ks ? 1u : acc0);
              }
              tc::mma_commit(&mi.bar_stage[w][0]);
              if (fin) tc::mma_commit(&mi.bar_full[w]);
            }
          }
        }
        if (ready) {
          ++t;
          ++n_done;
          if (last) { in_round = 0; round_len = G.flush; if (have_round) par_empty ^= 1u; have_round = true; } else ++in_round;
        }
        __syncwarp();
      }
    }
    __syncwarp();
    segment_tail<Q, H>(P, G, sg, mi, stages, cta, tid, kDrain ? acc : nullptr);
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
      if (H) {
        // per bin pair (b, b + 1): (x_b, x_b1, y_b, y_b1), (z_b, z_b1, w_b, w_b1) -- f32x2 operands over the two bins
        float* t = reinterpret_cast<float*>(mi.bins2 + 2 * (b >> 1));
        t[b & 1] = x; t[2 + (b & 1)] = y; t[4 + (b & 1)] = z; t[6 + (b & 1)] = w;
      } else {
        // interleave the two bin halves (lanes 0-15 / 16-31 read entry i of their half in the same instruction): the two
        // 16-byte reads of a warp then fall into different banks
        const int e = (b % (C::kBinsPad / 2)) * 2 + b / (C::kBinsPad / 2);
        mi.bins2[2 * e] = make_float4(x, x, y, y);
        mi.bins2[2 * e + 1] = make_float4(z, z, w, w);
      }
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
    constexpr bool kDual = GCS_TC_DUAL_ISSUE && Q <= 3;   // Q = 4: the issuer sits in a warpgroup of its own
    constexpr int kSplit = kDual ? kProd / 2 : kProd;     // the spare warp issues for [0, kSplit), epilogue warp 0 for the rest
    if (kDual && wid == kProd) issuer_role<Q, H, true>(P, G, mi, stages, tmem, blockIdx.x, tid, kSplit, kProd);
    else if (wid < kProd + C::kEpi) epilogue_role<Q, H>(P, G, mi, stages, tmem, blockIdx.x, tid);
    else if (H && GCS_TC_SELF_ISSUE) idle_role<Q, H>(P, G, mi, stages, blockIdx.x, tid);
    else issuer_role<Q, H, false>(P, G, mi, stages, tmem, blockIdx.x, tid, 0, kSplit);
  } else {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
    if (wid == C::kIssuerWarp && !(H && GCS_TC_SELF_ISSUE)) issuer_role<Q, H, false>(P, G, mi, stages, tmem, blockIdx.x, tid, 0, kProd);
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
    // 8 x 32 points between float64 drains: the float32 accumulator's truncation grows linearly with the interval --
    // deviation from the float64 kernels 7e-7 at 4, 1.4e-6 at 8, 2.9e-6 at 16 (tolerance 1e-5) -- while the drains (tcgen05.ld,
    // 19 conversions + 19 DADD per lane on the narrow float64 pipe) cost 5 % of the kernel at 4
    v = e ? atoi(e) : 8;
    if (v < 1) v = 1;
    if (v > 64) v = 64;
  }
  return v;
}

TcGeom make_geom(int sm_count, int n_units, int64_t cap, int n_parts, int n_prod, int tile_pts) {
  const int kProd = n_prod;
  TcGeom G;
  G.tiles_per_unit = (cap + tile_pts - 1) / tile_pts;
  G.total_tiles = G.tiles_per_unit * n_units;
  int64_t n_cta = (G.total_tiles + kProd - 1) / kProd;
  if (n_cta > sm_count) n_cta = sm_count;
  if (n_cta < 1) n_cta = 1;
  G.n_cta = (int)n_cta;
  G.n_parts = n_parts;
  G.flush = tc_flush_tiles() * tc::kTileK / tile_pts;   // GCS_TC_FLUSH counts 32-point tiles
  if (G.flush < 1) G.flush = 1;
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
  TcGeom G = make_geom(sm_count, n_units, cap, 0, 8, tc::kTileK);   // fewest producer warps, smallest tiles: most CTAs
  return (G.n_cta + n_units - 1) / n_units + 1;
}

bool bin_scan_tc_supported(const BinScanParams& P) {
  return P.resp == nullptr && !P.use_true_max && P.n_bins <= kMaxBins;
}

cudaError_t launch_bin_scan_tc(int sm_count, cudaStream_t st, const BinScanParams& P, int n_parts) {
  const int U = P.n_scans * P.n_hyp;
  if (tc_use_f16(P.inv_tau)) {
    if (P.n_bins <= 48) return launch_q<3, true>(st, P, make_geom(sm_count, U, P.cap, n_parts, TcCfg<3, true>::kProd, TcCfg<3, true>::kTilePts));
    return launch_q<4, true>(st, P, make_geom(sm_count, U, P.cap, n_parts, TcCfg<4, true>::kProd, TcCfg<4, true>::kTilePts));
  }
  if (P.n_bins <= 48) return launch_q<3, false>(st, P, make_geom(sm_count, U, P.cap, n_parts, TcCfg<3, false>::kProd, TcCfg<3, false>::kTilePts));
  return launch_q<4, false>(st, P, make_geom(sm_count, U, P.cap, n_parts, TcCfg<4, false>::kProd, TcCfg<4, false>::kTilePts));
}

}  // namespace gcs
