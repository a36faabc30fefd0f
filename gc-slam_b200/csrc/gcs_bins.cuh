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

}  // namespace gcs
