// gcs_bins.cuh -- layout constants shared by the bin-family kernels.
#pragma once
#include "gcs_common.cuh"

namespace gcs {

// Per-bin additive moments (one row of the "raw sums" block):
//   0      N        = sum w r
//   1..3   s_dir    = sum w r d
//   4..9   S        = sum w r d d^T   (xx,xy,xz,yy,yz,zz)
//   10..12 sum_p    = sum w r p
//   13..18 sum_ppT  = sum w r p p^T   (xx,xy,xz,yy,yz,zz)
//   19..24 sum_cov  = sum w r Sigma_pt (xx,xy,xz,yy,yz,zz)   (zero in the pipeline: fl/backend/pipeline.py:586-587)
constexpr int kNF = 19;       // features accumulated by the fused kernel
constexpr int kNFCov = 25;    // with per-point covariances (stand-alone ScanBinMomentMatch only)
constexpr int kRowLen = 25;   // row stride of the raw-sums block
constexpr int kMaxBins = 64;
// extras appended after kMaxBins... no: after n_bins*kRowLen
enum { kExEntDot = 0,  // sum_i sum_b r_ib (x_ib - c)
       kExEntLog,      // sum_i log(sum_b exp(x_ib - c))
       kExSumWdk,      // sum of deskewed weights   (DeskewConstantTwist cert numerator)
       kExSumWrs,      // sum of resampled weights  (DeskewConstantTwist cert denominator)
       kExCount,       // number of output rows this rank contributed (incl. padded rows)
       kNExtras = 8 };
enum { kMxResp = 0, kNMax = 2 };
// mass block per scan
enum { kMassAll = 0, kMassSel, kMassSelSq, kMassNSel, kNMass = 4 };

__host__ __device__ inline int raw_sums_len(int n_bins) { return n_bins * kRowLen + kNExtras; }

constexpr int kScanThreads = 128;  // fused kernel CTA size == points per tile
constexpr int kFeatStride = 20;    // doubles per staged point (w/sum, 18 features, 1/sum)
constexpr int kHalfF = 10;         // features per half-warp in the fused kernel's phase 2

// launch parameters of the fused per-point kernels (gcs_bins_scan.cu, gcs_bins_tc.cu)
struct BinScanParams {
  const double* pts; const double* t; const double* w; const uint8_t* ring; const uint8_t* tag;
  int64_t n_raw;     // local raw rows per scan
  int64_t cap;       // local output rows per scan
  int64_t n_sel;     // local selected rows = ceil(n_raw / stride)
  int64_t stride;
  int n_scans, n_hyp, n_bins;
  const double* t0s; const double* t1s; const double* xi; const double* bin_dirs;
  double origin[3];
  double inv_tau, shift, eps_mass;
  int use_true_max;
  const double* mass;  // (S, kNMass), already global
  double* rs_pts; double* rs_t; double* rs_w; uint8_t* rs_ring; uint8_t* rs_tag;
  double* dk_pts; double* dk_w; double* resp;
  double* partial;  // (U, ctas_per_unit, part_len)
  int part_len;
};

// Tensor-core variant (gcs_bins_tc.cu).  Returns the number of partial slots per unit it needs for n_units units of
// `cap` rows on a device with sm_count SMs; the caller zero-fills P.partial, the kernel fills it as (U, n_parts, part_len).
int bin_scan_tc_parts(int sm_count, int n_units, int64_t cap);
bool bin_scan_tc_supported(const BinScanParams& P);
cudaError_t launch_bin_scan_tc(int sm_count, cudaStream_t st, const BinScanParams& P, int n_parts);

}  // namespace gcs
