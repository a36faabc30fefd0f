// gcs_prims_surfel.cu -- LiDAR surfel extraction (a10) and MeasurementBatch builders.
//
//   S1 surfel_center      weighted centre of the unmasked points (two-level fixed-order reduction)
//   S2 surfel_cell_key    MA-hex 3-D cell of every point: floor in float64 exactly as hex_cell_3d_batch, Python-style mod
//   S3 surfel_rank_*      stable rank of every point inside its cell ("the 32 lowest original indices per cell"):
//                         per-chunk ordered counting with warp match + a per-cell exclusive scan over chunks.
//                         Equivalent to the reference's stable argsort by (masked, cell) without sorting anything.
//   S4 surfel_fit         one thread per cell: weighted centroid / covariance, Jacobi eigh, tangent basis, Wishart-
//                         regularised precision, vMF kappa
//   S5 surfel_select      valid cells in ascending cell order -> first n_surfel slots of the LiDAR slice (info form)
#include "gcs_common.cuh"

namespace gcs {

constexpr int kSurfThreads = 256;

struct SurfelGeom {
  int nc1, nc2, ncz, n_cells, max_occ, min_points;
  double h;
};
// Unit axis (blockIdx.y = hypothesis): element strides of the per-unit inputs (0 = shared by all units).  Work arrays
// and outputs are stacked per unit.
struct UnitStrides {
  int64_t pts, w, ts;
};

__device__ __forceinline__ bool surfel_point_ok(const double* p) {
  // jnp.all(jnp.abs(points) < 0.1 * GC_NONFINITE_SENTINEL)   (lidar_surfel_extraction.py:259-262); NaN -> false
  return (fabs(p[0]) < 1e5) && (fabs(p[1]) < 1e5) && (fabs(p[2]) < 1e5);
}

// ---- S1 ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kSurfThreads) surfel_center_partial_kernel(const double* __restrict__ pts,
                                                                             const double* __restrict__ w, int64_t n,
                                                                             int64_t per_block, double* __restrict__ part,
                                                                             UnitStrides U) {
  __shared__ double sred[4][kSurfThreads / 32];
  pts += blockIdx.y * U.pts; w += blockIdx.y * U.w; part += (int64_t)blockIdx.y * gridDim.x * 4;
  const int64_t i0 = (int64_t)blockIdx.x * per_block;
  const int64_t i1 = (i0 + per_block < n) ? i0 + per_block : n;
  double a[4] = {0, 0, 0, 0};
  for (int64_t i = i0 + threadIdx.x; i < i1; i += kSurfThreads) {
    const double p[3] = {pts[3 * i], pts[3 * i + 1], pts[3 * i + 2]};
    const double we = surfel_point_ok(p) ? w[i] : 0.0;
    a[0] += p[0] * we; a[1] += p[1] * we; a[2] += p[2] * we; a[3] += we;
  }
  for (int k = 0; k < 4; ++k) {
    double v = warp_sum(a[k]);
    if ((threadIdx.x & 31) == 0) sred[k][threadIdx.x >> 5] = v;
  }
  __syncthreads();
  if (threadIdx.x < 4) {
    double s = 0.0;
    for (int g = 0; g < kSurfThreads / 32; ++g) s += sred[threadIdx.x][g];
    part[blockIdx.x * 4 + threadIdx.x] = s;
  }
}
__global__ void surfel_center_final_kernel(const double* __restrict__ part, int n_parts, double eig_min,
                                           double* __restrict__ center) {
  if (threadIdx.x != 0) return;
  part += (int64_t)blockIdx.x * n_parts * 4; center += blockIdx.x * 4;
  double a[4] = {0, 0, 0, 0};
  for (int c = 0; c < n_parts; ++c)
    for (int k = 0; k < 4; ++k) a[k] += part[4 * c + k];
  const double ws = a[3] + eig_min;
  center[0] = a[0] / ws; center[1] = a[1] / ws; center[2] = a[2] / ws; center[3] = a[3];
}

// ---- S2 ------------------------------------------------------------------------------------------------
__device__ __forceinline__ int pymod(int a, int n) { int r = a % n; return r < 0 ? r + n : r; }

__global__ void __launch_bounds__(kSurfThreads) surfel_cell_key_kernel(const double* __restrict__ pts,
                                                                       const double* __restrict__ center, int64_t n,
                                                                       SurfelGeom G, int32_t* __restrict__ key,
                                                                       UnitStrides U) {
  const int64_t i = (int64_t)blockIdx.x * kSurfThreads + threadIdx.x;
  if (i >= n) return;
  pts += blockIdx.y * U.pts; center += blockIdx.y * 4; key += (int64_t)blockIdx.y * n;
  const double p[3] = {pts[3 * i], pts[3 * i + 1], pts[3 * i + 2]};
  int k = G.n_cells;  // masked points sort last and never enter a bucket
  if (surfel_point_ok(p)) {
    const double x = p[0] - center[0], y = p[1] - center[1], z = p[2] - center[2];
    const double s2 = x * 0.5 + y * (sqrt(3.0) * 0.5);  // ma_hex_web.py:233-235
    const int c1 = (int)floor(x / G.h), c2 = (int)floor(s2 / G.h), cz = (int)floor(z / G.h);
    k = pymod(c1, G.nc1) * (G.nc2 * G.ncz) + pymod(c2, G.nc2) * G.ncz + pymod(cz, G.ncz);
  }
  key[i] = k;
}

// ---- S3 ------------------------------------------------------------------------------------------------
// One WARP walks a contiguous chunk of points in index order, 32 at a time, keeping a running per-cell count in its own
// shared-memory table.  local_rank[i] = number of earlier points of the same chunk in the same cell;
// hist[chunk][cell] = points of this chunk per cell.  No barrier anywhere: a CTA is a single warp, seven of them fit the
// shared memory of an SM, and a batch of hypotheses brings a thousand independent chunks (the 1024-thread version let
// one warp at a time touch the table, 32 barriers per 1024 points: 239 us for 64 hypotheses).
__global__ void __launch_bounds__(32) surfel_rank_chunk_kernel(const int32_t* __restrict__ key, int64_t n,
                                                               int64_t per_chunk, int n_keys,
                                                               int32_t* __restrict__ local_rank,
                                                               int32_t* __restrict__ hist) {
  extern __shared__ int s_cnt[];  // n_keys ints
  key += (int64_t)blockIdx.y * n; local_rank += (int64_t)blockIdx.y * n; hist += (int64_t)blockIdx.y * gridDim.x * n_keys;
  const int lane = threadIdx.x;
  for (int k = lane; k < n_keys; k += 32) s_cnt[k] = 0;
  __syncwarp();
  const int64_t i0 = (int64_t)blockIdx.x * per_chunk;
  const int64_t i1 = (i0 + per_chunk < n) ? i0 + per_chunk : n;
  constexpr int kAhead = 8;   // the keys of eight 32-point steps are loaded together: one memory round trip per 256 points
  for (int64_t base0 = i0; base0 < i1; base0 += 32 * kAhead) {
    int kk[kAhead];
#pragma unroll
    for (int j = 0; j < kAhead; ++j) {
      const int64_t i = base0 + 32 * j + lane;
      kk[j] = i < i1 ? key[i] : -1 - lane;  // inactive lanes get unique negative keys
    }
#pragma unroll
    for (int j = 0; j < kAhead; ++j) {
      const int64_t i = base0 + 32 * j + lane;
      const bool on = i < i1;
      const int k = kk[j];
      const unsigned peers = __match_any_sync(0xffffffffu, k);
      const int before = __popc(peers & ((1u << lane) - 1u));
      const bool leader = (lane == 31 - __clz(peers));  // highest lane of the group updates the counter
      if (on) {
        const int basecnt = s_cnt[k];
        local_rank[i] = basecnt + before;
        __syncwarp(peers);
        if (leader) s_cnt[k] = basecnt + __popc(peers);
      }
      __syncwarp();
    }
  }
  for (int k = lane; k < n_keys; k += 32) hist[(int64_t)blockIdx.x * n_keys + k] = s_cnt[k];
}
// per key: exclusive scan over chunks (in place) + total
// ... and the list of the cells that reach the point budget (the only ones the plane fit has to look at; order = arrival
// order of the atomic: the fit of a cell does not depend on it), `valid` cleared for all others
__global__ void surfel_rank_scan_kernel(int32_t* __restrict__ hist, int n_chunks, int n_keys, int32_t* __restrict__ total,
                                        int min_points, int32_t* __restrict__ occ_list, int32_t* __restrict__ n_occ,
                                        uint8_t* __restrict__ cell_valid) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n_keys) return;
  hist += (int64_t)blockIdx.y * n_chunks * n_keys; total += (int64_t)blockIdx.y * n_keys;
  occ_list += (int64_t)blockIdx.y * (n_keys - 1); cell_valid += (int64_t)blockIdx.y * (n_keys - 1);
  int acc = 0;
  // eight chunk counts are loaded before any prefix is stored (the in-place store would otherwise order every load
  // behind it: one memory round trip per chunk)
  for (int c0 = 0; c0 < n_chunks; c0 += 8) {
    int v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = (c0 + j < n_chunks) ? hist[(int64_t)(c0 + j) * n_keys + k] : 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      if (c0 + j < n_chunks) hist[(int64_t)(c0 + j) * n_keys + k] = acc;
      acc += v[j];
    }
  }
  total[k] = acc;
  if (k < n_keys - 1) {           // key n_keys - 1 collects the masked points
    if (acc >= min_points && acc > 0) occ_list[atomicAdd(n_occ + blockIdx.y, 1)] = k;
    else cell_valid[k] = 0;
  }
}
__global__ void surfel_bucket_init_kernel(int32_t* __restrict__ bucket, int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) bucket[i] = -1;
}
__global__ void __launch_bounds__(kSurfThreads) surfel_bucket_fill_kernel(const int32_t* __restrict__ key,
                                                                          const int32_t* __restrict__ local_rank,
                                                                          const int32_t* __restrict__ hist, int64_t n,
                                                                          int64_t per_chunk, int n_keys, SurfelGeom G,
                                                                          int32_t* __restrict__ bucket, int n_chunks) {
  const int64_t i = (int64_t)blockIdx.x * kSurfThreads + threadIdx.x;
  if (i >= n) return;
  key += (int64_t)blockIdx.y * n; local_rank += (int64_t)blockIdx.y * n; hist += (int64_t)blockIdx.y * n_chunks * n_keys;
  bucket += (int64_t)blockIdx.y * G.n_cells * G.max_occ;
  const int k = key[i];
  if (k >= G.n_cells) return;
  const int chunk = (int)(i / per_chunk);
  const int rank = hist[(int64_t)chunk * n_keys + k] + local_rank[i];
  if (rank < G.max_occ) bucket[(int64_t)k * G.max_occ + rank] = (int32_t)i;
}

// ---- S4 ------------------------------------------------------------------------------------------------
struct CellFit {  // per cell, SoA in workspace
  double* centroid;  // (C,3) in original (un-centred) coordinates
  double* Sigma;     // (C,9)
  double* normal;    // (C,3)
  double* kappa;     // (C)
  double* w;         // (C)
  double* t;         // (C)
  uint8_t* valid;    // (C)
};

__device__ __forceinline__ CellFit cellfit_unit(CellFit F, int64_t u, int n_cells) {
  const int64_t o = u * n_cells;
  F.centroid += 3 * o; F.Sigma += 9 * o; F.normal += 3 * o; F.kappa += o; F.w += o; F.t += o; F.valid += o;
  return F;
}

__device__ __forceinline__ void normalize3(double* v, double eps) {
  const double inv = 1.0 / (sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]) + eps);
  v[0] *= inv; v[1] *= inv; v[2] *= inv;
}

// Plane fit of the listed cells (surfel_rank_scan_kernel) in two kernels:
//   surfel_moments_kernel  eight lanes per cell: weighted centroid, then the weighted scatter about it (two passes over the
//                          cell's <= 32 points, gathers four deep instead of 32 deep, partial sums in a fixed xor tree);
//   surfel_fit_kernel      one thread per cell: Jacobi eigh, tangent basis, Wishart-regularised precision, vMF kappa.
// The reference's third pass -- in-plane spreads sum w (d . e)^2 -- is e^T S e with the scatter S of the second pass
// (same number up to rounding), so the algebra kernel does not touch the points.  Groups / threads stride over the list:
// the grids do not depend on how many cells a scan fills, and the float64-heavy algebra runs in full warps.
constexpr int kFitLanes = 8;
constexpr int kCellMom = 12;   // wsum, tsum, mu (3), scatter s00 s01 s02 s11 s12 s22, cnt
__device__ __forceinline__ double group_sum8(double v, unsigned gmask) {
  v += __shfl_xor_sync(gmask, v, 4);
  v += __shfl_xor_sync(gmask, v, 2);
  v += __shfl_xor_sync(gmask, v, 1);
  return v;
}
__global__ void __launch_bounds__(128) surfel_moments_kernel(const double* __restrict__ pts, const double* __restrict__ ts,
                                                             const double* __restrict__ w, const double* __restrict__ center,
                                                             const int32_t* __restrict__ bucket,
                                                             const int32_t* __restrict__ total, SurfelGeom G, UnitStrides U,
                                                             int n_keys, const int32_t* __restrict__ occ_list,
                                                             const int32_t* __restrict__ n_occ, double* __restrict__ mom) {
  {
    const int64_t u = blockIdx.y;
    pts += u * U.pts; ts += u * U.ts; w += u * U.w; center += u * 4;
    bucket += u * G.n_cells * G.max_occ; total += u * n_keys;
    occ_list += u * G.n_cells; mom += u * G.n_cells * kCellMom;
  }
  const int n_list = n_occ[blockIdx.y];
  const int lane8 = threadIdx.x & (kFitLanes - 1);
  const unsigned gmask = 0xffu << (threadIdx.x & 24);
  const int n_groups = (gridDim.x * blockDim.x) / kFitLanes;
  const double eps = 1e-12;
  const double cx = center[0], cy = center[1], cz = center[2];
  for (int e = (blockIdx.x * blockDim.x + threadIdx.x) / kFitLanes; e < n_list; e += n_groups) {
    const int c = occ_list[e];
    const int cnt = total[c] < G.max_occ ? total[c] : G.max_occ;
    // pass 1: weighted centroid
    double wsum = 0.0, m0 = 0.0, m1 = 0.0, m2 = 0.0, tsum = 0.0;
    for (int o = lane8; o < cnt; o += kFitLanes) {
      const int i = bucket[(int64_t)c * G.max_occ + o];
      const double wi = w[i];
      m0 += (pts[3 * i] - cx) * wi; m1 += (pts[3 * i + 1] - cy) * wi; m2 += (pts[3 * i + 2] - cz) * wi;
      wsum += wi; tsum += ts[i];
    }
    wsum = group_sum8(wsum, gmask); m0 = group_sum8(m0, gmask); m1 = group_sum8(m1, gmask); m2 = group_sum8(m2, gmask);
    tsum = group_sum8(tsum, gmask);
    const double w_sum = wsum + eps;
    const double mu[3] = {m0 / w_sum, m1 / w_sum, m2 / w_sum};
    // pass 2: weighted scatter.  Absent slots gather point 0 with weight 0 in the reference
    // (lidar_surfel_extraction.py:115-121): they add exactly 0 to every weighted sum, so they are skipped here.
    double s00 = 0, s01 = 0, s02 = 0, s11 = 0, s12 = 0, s22 = 0;
    for (int o = lane8; o < cnt; o += kFitLanes) {
      const int i = bucket[(int64_t)c * G.max_occ + o];
      const double wi = w[i];
      const double d0 = (pts[3 * i] - cx) - mu[0], d1 = (pts[3 * i + 1] - cy) - mu[1], d2 = (pts[3 * i + 2] - cz) - mu[2];
      s00 += wi * d0 * d0; s01 += wi * d0 * d1; s02 += wi * d0 * d2;
      s11 += wi * d1 * d1; s12 += wi * d1 * d2; s22 += wi * d2 * d2;
    }
    s00 = group_sum8(s00, gmask); s01 = group_sum8(s01, gmask); s02 = group_sum8(s02, gmask);
    s11 = group_sum8(s11, gmask); s12 = group_sum8(s12, gmask); s22 = group_sum8(s22, gmask);
    if (lane8 == 0) {
      double* m = mom + (int64_t)e * kCellMom;
      m[0] = wsum; m[1] = tsum; m[2] = mu[0]; m[3] = mu[1]; m[4] = mu[2];
      m[5] = s00; m[6] = s01; m[7] = s02; m[8] = s11; m[9] = s12; m[10] = s22; m[11] = (double)cnt;
    }
  }
}

__global__ void __launch_bounds__(128) surfel_fit_kernel(const double* __restrict__ center, SurfelGeom G, gcs_surfel_cfg cfg,
                                                         CellFit F, const int32_t* __restrict__ occ_list,
                                                         const int32_t* __restrict__ n_occ, const double* __restrict__ mom) {
  {
    const int64_t u = blockIdx.y;
    center += u * 4; occ_list += u * G.n_cells; mom += u * G.n_cells * kCellMom;
    F = cellfit_unit(F, u, G.n_cells);
  }
  const int n_list = n_occ[blockIdx.y];
  const double eps = 1e-12, eig_min = cfg.eig_min;
  const double cx = center[0], cy = center[1], cz = center[2];
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < n_list; e += gridDim.x * blockDim.x) {
  const int c = occ_list[e];
  const double* m = mom + (int64_t)e * kCellMom;
  const double wsum = m[0], tsum = m[1];
  const double mu[3] = {m[2], m[3], m[4]};
  const double s00 = m[5], s01 = m[6], s02 = m[7], s11 = m[8], s12 = m[9], s22 = m[10];
  const int cnt = (int)m[11];
  const double w_sum = wsum + eps;
  Mat3 cov;
  cov(0, 0) = s00 / w_sum + eig_min; cov(1, 1) = s11 / w_sum + eig_min; cov(2, 2) = s22 / w_sum + eig_min;
  cov(0, 1) = cov(1, 0) = s01 / w_sum; cov(0, 2) = cov(2, 0) = s02 / w_sum; cov(1, 2) = cov(2, 1) = s12 / w_sum;
  double ev[3];
  Mat3 V;
  eigh3(cov, ev, V);
  double nrm[3] = {V(0, 0), V(1, 0), V(2, 0)};
  const double sgn = nrm[2] < 0.0 ? -1.0 : 1.0;  // deterministic sign (:130)
  nrm[0] *= sgn; nrm[1] *= sgn; nrm[2] *= sgn;
  normalize3(nrm, eps);
  double nn[3] = {nrm[0], nrm[1], nrm[2]};
  normalize3(nn, eps);  // _orthonormal_basis_from_normal normalises again (:72)
  double e1[3];
  if (fabs(nn[2]) < 0.9) { e1[0] = -nn[1]; e1[1] = nn[0]; e1[2] = 0.0; }
  else { e1[0] = -nn[2]; e1[1] = 0.0; e1[2] = nn[0]; }
  normalize3(e1, eps);
  double e2[3] = {nn[1] * e1[2] - nn[2] * e1[1], nn[2] * e1[0] - nn[0] * e1[2], nn[0] * e1[1] - nn[1] * e1[0]};
  normalize3(e2, eps);
  // in-plane spreads: sum w (d . e)^2 = e^T S e with the scatter S of the moments kernel
  auto quad = [&](const double* e) {
    return e[0] * (s00 * e[0] + s01 * e[1] + s02 * e[2]) + e[1] * (s01 * e[0] + s11 * e[1] + s12 * e[2]) +
           e[2] * (s02 * e[0] + s12 * e[1] + s22 * e[2]);
  };
  const double v1 = quad(e1), v2 = quad(e2);
  // An empty cell in the reference still gathers max_occ copies of point 0 with zero weight; all sums are 0 there too.
  const double var_e1 = v1 / w_sum + cfg.sensor_noise_var_per_axis;
  const double var_e2 = v2 / w_sum + cfg.sensor_noise_var_per_axis;
  const double sig_perp = fmax(ev[0], eig_min);
  const double var_perp = sig_perp + cfg.sensor_noise_var_per_axis;
  const double D[3] = {fmax(var_e1, eig_min), fmax(var_e2, eig_min), fmax(var_perp, eig_min)};
  const double* B[3] = {e1, e2, nrm};  // V = [e1 e2 normal]
  Mat3 Sigma;
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) {
      double a = 0.0;
      for (int k = 0; k < 3; ++k) a += B[k][i] * D[k] * B[k][j];
      Sigma(i, j) = a;
    }
  Mat3 S2;
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) S2(i, j) = 0.5 * (Sigma(i, j) + Sigma(j, i)) + ((i == j) ? eig_min : 0.0);
  Mat3 S3 = S2;
  S3(0, 0) += eig_min; S3(1, 1) += eig_min; S3(2, 2) += eig_min;
  Mat3 Lam = mat3_inv(S3);
  const double psi = fmax(cfg.wishart_psi_scale, eps);
  const double reg = cfg.wishart_nu / psi;
  Mat3 Lreg;
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) {
      const double sym = 0.5 * (Lam(i, j) + Lam(j, i));
      const double l1 = sym + ((i == j) ? reg : 0.0);
      Lreg(i, j) = l1;
    }
  Mat3 Lreg2;
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) Lreg2(i, j) = 0.5 * (Lreg(i, j) + Lreg(j, i)) + ((i == j) ? eig_min : 0.0);
  Mat3 Sr = mat3_inv(Lreg2);
  double kap = cfg.kappa_main_scale / sqrt(fmax(sig_perp, eig_min));
  kap = fmin(fmax(kap, cfg.kappa_min), cfg.kappa_max);
  const bool valid = (cnt >= G.min_points) && (wsum > 0.0);
  F.centroid[3 * c] = mu[0] + cx; F.centroid[3 * c + 1] = mu[1] + cy; F.centroid[3 * c + 2] = mu[2] + cz;
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) F.Sigma[9 * c + 3 * i + j] = 0.5 * (Sr(i, j) + Sr(j, i)) + ((i == j) ? eig_min : 0.0);
  F.normal[3 * c] = nrm[0]; F.normal[3 * c + 1] = nrm[1]; F.normal[3 * c + 2] = nrm[2];
  F.kappa[c] = kap; F.w[c] = wsum; F.t[c] = tsum / w_sum;  // unweighted stamp sum / weight sum, as written (:160)
  F.valid[c] = valid ? 1 : 0;
  }   // list entries of this thread
}


// ---- S5 ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024) surfel_select_kernel(CellFit F, SurfelGeom G, gcs_meas_batch B, double eps_lift,
                                                             int32_t* __restrict__ out_n_valid,
                                                             const int32_t* __restrict__ total,
                                                             int32_t* __restrict__ out_count, int n_keys) {
  __shared__ int s_scan[1024];
  __shared__ int s_total;
  const int tid = threadIdx.x;
  {
    const int64_t u = blockIdx.x;
    F = cellfit_unit(F, u, G.n_cells);
    B = meas_batch_unit(B, u);
    out_n_valid += u; total += u * n_keys;
    if (out_count) out_count += u * G.n_cells;
  }
  const int per = (G.n_cells + 1023) / 1024;
  const int c0 = tid * per, c1 = (c0 + per < G.n_cells) ? c0 + per : G.n_cells;
  int cnt = 0;
  for (int c = c0; c < c1; ++c) cnt += F.valid[c];
  // exclusive prefix of the 1024 per-thread counts: shuffle scan per warp + scan of the 32 warp totals (a single thread
  // walking the 1024 counts took 15 us of every scan's 38)
  const int lane = tid & 31, warp = tid >> 5;
  int inc = cnt;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int v = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += v;
  }
  if (lane == 31) s_scan[warp] = inc;
  __syncthreads();
  if (warp == 0) {
    const int w = s_scan[lane];
    int winc = w;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int v = __shfl_up_sync(0xffffffffu, winc, o);
      if (lane >= o) winc += v;
    }
    s_scan[32 + lane] = winc - w;          // exclusive offset of warp `lane`
    if (lane == 31) s_total = winc;
  }
  __syncthreads();
  int slot = inc - cnt + s_scan[32 + warp];
  // the cells that become surfels, in cell order (slot -> cell), then one thread per SURFEL: the valid cells are a third of
  // the grid and unevenly spread over the threads' cell ranges
  extern __shared__ int s_list[];          // n_surfel entries
  for (int c = c0; c < c1; ++c) {
    if (out_count) out_count[c] = total[c] < G.max_occ ? total[c] : G.max_occ;
    if (!F.valid[c]) continue;
    if (slot < B.n_surfel) s_list[slot] = c;
    ++slot;
  }
  __syncthreads();
  const int n_sel = s_total < B.n_surfel ? s_total : B.n_surfel;
  for (int sl = tid; sl < n_sel; sl += 1024) {
    const int c = s_list[sl];
    const int r = B.n_feat + sl;
    Mat3 S;
    for (int k = 0; k < 9; ++k) S.m[k] = F.Sigma[9 * c + k];
    S(0, 0) += eps_lift; S(1, 1) += eps_lift; S(2, 2) += eps_lift;
    Mat3 L = mat3_inv(S);  // measurement_batch_add_lidar_surfels (measurement_batch.py:299-301)
    const double mu[3] = {F.centroid[3 * c], F.centroid[3 * c + 1], F.centroid[3 * c + 2]};
    double th[3];
    mat3_vec(L, mu, th);
    for (int k = 0; k < 9; ++k) B.Lambdas[9 * r + k] = L.m[k];
    const double kap = F.kappa[c];
    const double nz = fmin(fmax(F.normal[3 * c + 2], -1.0), 1.0);
    const double g = 0.25 + 0.5 * (nz + 1.0) / 2.0;
    for (int k = 0; k < 3; ++k) {
      B.thetas[3 * r + k] = th[k];
      B.etas[9 * r + k] = kap * F.normal[3 * c + k];
      B.etas[9 * r + 3 + k] = 0.0;
      B.etas[9 * r + 6 + k] = 0.0;
      B.colors[3 * r + k] = g;
    }
    B.weights[r] = F.w[c];
    B.sources[r] = 1;
    B.source_indices[r] = sl;
    B.valid[r] = 1;
    B.timestamps[r] = F.t[c];
  }
  if (tid == 0) out_n_valid[0] = s_total < B.n_surfel ? s_total : B.n_surfel;
}

__global__ void camera_batch_kernel(const double* __restrict__ pos, const double* __restrict__ cov,
                                    const double* __restrict__ dir, const double* __restrict__ kap,
                                    const double* __restrict__ w, const double* __restrict__ ts,
                                    const double* __restrict__ col, int n, double eps_lift, gcs_meas_batch B) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  Mat3 S;
  for (int k = 0; k < 9; ++k) S.m[k] = cov[9 * i + k];
  S(0, 0) += eps_lift; S(1, 1) += eps_lift; S(2, 2) += eps_lift;
  Mat3 L = mat3_inv(S);
  const double mu[3] = {pos[3 * i], pos[3 * i + 1], pos[3 * i + 2]};
  double th[3];
  mat3_vec(L, mu, th);
  for (int k = 0; k < 9; ++k) B.Lambdas[9 * i + k] = L.m[k];
  for (int k = 0; k < 3; ++k) {
    B.thetas[3 * i + k] = th[k];
    B.etas[9 * i + k] = kap[i] * dir[3 * i + k];
    B.etas[9 * i + 3 + k] = 0.0;
    B.etas[9 * i + 6 + k] = 0.0;
    B.colors[3 * i + k] = col ? fmin(fmax(col[3 * i + k], 0.0), 1.0) : 0.5;
  }
  B.weights[i] = w[i];
  B.sources[i] = 0;
  B.source_indices[i] = i;
  B.valid[i] = 1;
  B.timestamps[i] = ts[i];
}

}  // namespace gcs

using namespace gcs;

static int64_t cdiv(int64_t a, int64_t b) { return (a + b - 1) / b; }

static int check_batch(gcs_ctx* ctx, const gcs_meas_batch* b, const char* who) {
  GCS_REQUIRE(ctx, b && b->Lambdas && b->thetas && b->etas && b->weights && b->sources && b->source_indices && b->valid &&
                       b->timestamps && b->colors, "%s: measurement batch pointer is NULL", who);
  GCS_REQUIRE(ctx, b->n_feat >= 0 && b->n_surfel >= 0 && b->n_feat + b->n_surfel >= 1, "%s: bad batch budget", who);
  GCS_REQUIRE(ctx, b->n_surfel <= 10240, "%s: n_surfel=%d exceeds the built budget of 10240 (the cell grid holds 8192)", who, b->n_surfel);
  return GCS_OK;
}

extern "C" {

int gcs_batch_from_camera_splats(gcs_ctx* ctx, void* stream, const double* positions, const double* covariances,
                                 const double* directions, const double* kappas, const double* weights,
                                 const double* timestamps, const double* colors, int32_t n, double eps_lift,
                                 const gcs_meas_batch* batch) {
  if (!ctx) return GCS_EINVAL;
  GCS_CHECK_CUDA(ctx, cudaSetDevice(ctx->device));
  int rc = check_batch(ctx, batch, "batch_from_camera_splats");
  if (rc) return rc;
  GCS_REQUIRE(ctx, n >= 0, "batch_from_camera_splats: n=%d", n);
  const int nv = n < batch->n_feat ? n : batch->n_feat;  // truncate to the N_FEAT budget
  if (nv == 0) return GCS_OK;
  GCS_REQUIRE(ctx, positions && covariances && directions && kappas && weights && timestamps, "batch_from_camera_splats: NULL input");
  camera_batch_kernel<<<(nv + 127) / 128, 128, 0, (cudaStream_t)stream>>>(positions, covariances, directions, kappas, weights,
                                                                          timestamps, colors, nv, eps_lift, *batch);
  GCS_LAUNCH_CHECK(ctx);
  return GCS_OK;
}

static int surfels_launch(gcs_ctx* ctx, cudaStream_t st, const double* pts, const double* timestamps, const double* weights,
                          int64_t n, int n_units, UnitStrides U, const gcs_surfel_cfg* cfg, const gcs_meas_batch* batch,
                          int32_t* out_n_valid, int32_t* out_bucket, int32_t* out_count, const char* who) {
  int rc = check_batch(ctx, batch, who);
  if (rc) return rc;
  GCS_REQUIRE(ctx, cfg && pts && timestamps && weights && out_n_valid && n >= 1, "%s: bad args", who);
  GCS_REQUIRE(ctx, cfg->n_cells_1 > 0 && cfg->n_cells_2 > 0 && cfg->n_cells_z > 0 && cfg->max_occupants > 0 &&
                       cfg->max_occupants <= 1024, "%s: bad grid config", who);
  GCS_REQUIRE(ctx, n < (1ll << 31), "%s: n too large for int32 point indices", who);
  SurfelGeom G;
  G.nc1 = cfg->n_cells_1; G.nc2 = cfg->n_cells_2; G.ncz = cfg->n_cells_z;
  G.n_cells = G.nc1 * G.nc2 * G.ncz;
  G.max_occ = cfg->max_occupants; G.min_points = cfg->min_points_per_voxel;
  G.h = cfg->voxel_size_m > 1e-12 ? cfg->voxel_size_m : 1e-12;
  const int n_keys = G.n_cells + 1;
  GCS_REQUIRE(ctx, (size_t)n_keys * sizeof(int) <= 160 * 1024, "%s: %d cells exceed the shared-memory counter", who, G.n_cells);
  // chunks of the per-cell ranking (one warp each; the ranks do not depend on the chunking): as many as the device holds
  // at once over all units, at least 512 points each
  const int warps_resident = ctx->sm_count * (int)(200 * 1024 / ((size_t)n_keys * sizeof(int)) > 1 ? 200 * 1024 / ((size_t)n_keys * sizeof(int)) : 1);
  int n_chunks = (int)cdiv(n, 512);
  int max_chunks = (warps_resident + n_units - 1) / n_units;
  if (max_chunks < 4) max_chunks = 4;
  if (n_chunks > max_chunks) n_chunks = max_chunks;
  const int64_t per_chunk = cdiv(cdiv(n, n_chunks), 32) * 32;
  n_chunks = (int)cdiv(n, per_chunk);
  const int n_cblocks = (int)(cdiv(n, 8192) < 256 ? cdiv(n, 8192) : 256);
  const int64_t per_cblock = cdiv(n, n_cblocks);
  // workspace carve-up: every array stacked per unit
  const size_t H = (size_t)n_units;
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t o = off; off += (bytes * H + 255) & ~(size_t)255; return o; };
  const size_t o_part = take((size_t)n_cblocks * 4 * 8), o_center = take(4 * 8), o_key = take((size_t)n * 4),
               o_lrank = take((size_t)n * 4), o_hist = take((size_t)n_chunks * n_keys * 4), o_total = take((size_t)n_keys * 4),
               o_bucket = take((size_t)G.n_cells * G.max_occ * 4), o_cen = take((size_t)G.n_cells * 3 * 8),
               o_sig = take((size_t)G.n_cells * 9 * 8), o_nrm = take((size_t)G.n_cells * 3 * 8), o_kap = take((size_t)G.n_cells * 8),
               o_w = take((size_t)G.n_cells * 8), o_t = take((size_t)G.n_cells * 8), o_val = take((size_t)G.n_cells),
               o_occ = take((size_t)G.n_cells * 4), o_nocc = take(4), o_mom = take((size_t)G.n_cells * kCellMom * 8);
  rc = gcs_ws_reserve(ctx, off);
  if (rc) return rc;
  char* ws = (char*)ctx->ws;
  double* part = (double*)(ws + o_part);
  double* center = (double*)(ws + o_center);
  int32_t* key = (int32_t*)(ws + o_key);
  int32_t* lrank = (int32_t*)(ws + o_lrank);
  int32_t* hist = (int32_t*)(ws + o_hist);
  int32_t* total = (int32_t*)(ws + o_total);
  int32_t* bucket = out_bucket ? out_bucket : (int32_t*)(ws + o_bucket);
  CellFit F;
  F.centroid = (double*)(ws + o_cen); F.Sigma = (double*)(ws + o_sig); F.normal = (double*)(ws + o_nrm);
  F.kappa = (double*)(ws + o_kap); F.w = (double*)(ws + o_w); F.t = (double*)(ws + o_t); F.valid = (uint8_t*)(ws + o_val);

  const unsigned Hu = (unsigned)n_units;
  surfel_center_partial_kernel<<<dim3(n_cblocks, Hu), kSurfThreads, 0, st>>>(pts, weights, n, per_cblock, part, U);
  GCS_LAUNCH_CHECK(ctx);
  surfel_center_final_kernel<<<Hu, 32, 0, st>>>(part, n_cblocks, cfg->eig_min, center);
  GCS_LAUNCH_CHECK(ctx);
  surfel_cell_key_kernel<<<dim3((unsigned)cdiv(n, kSurfThreads), Hu), kSurfThreads, 0, st>>>(pts, center, n, G, key, U);
  GCS_LAUNCH_CHECK(ctx);
  GCS_CHECK_CUDA(ctx, gcs_smem_attr_once((const void*)surfel_rank_chunk_kernel, 160 * 1024));
  surfel_rank_chunk_kernel<<<dim3(n_chunks, Hu), 32, (size_t)n_keys * sizeof(int), st>>>(key, n, per_chunk, n_keys, lrank, hist);
  GCS_LAUNCH_CHECK(ctx);
  int32_t* occ_list = (int32_t*)(ws + o_occ);
  int32_t* n_occ = (int32_t*)(ws + o_nocc);
  GCS_CHECK_CUDA(ctx, cudaMemsetAsync(n_occ, 0, sizeof(int32_t) * H, st));
  surfel_rank_scan_kernel<<<dim3((n_keys + 255) / 256, Hu), 256, 0, st>>>(hist, n_chunks, n_keys, total, G.min_points, occ_list, n_occ,
                                                                           F.valid);
  GCS_LAUNCH_CHECK(ctx);
  // the -1 fill of unused bucket slots only matters to a caller that receives the bucket (the reference's BucketResult);
  // the plane fit reads the first min(count, max_occ) slots of a cell, all of which bucket_fill writes
  if (out_bucket) {
    const int64_t nb = (int64_t)G.n_cells * G.max_occ * n_units;
    surfel_bucket_init_kernel<<<(unsigned)cdiv(nb, 256), 256, 0, st>>>(bucket, nb);
    GCS_LAUNCH_CHECK(ctx);
  }
  surfel_bucket_fill_kernel<<<dim3((unsigned)cdiv(n, kSurfThreads), Hu), kSurfThreads, 0, st>>>(key, lrank, hist, n, per_chunk, n_keys, G,
                                                                                            bucket, n_chunks);
  GCS_LAUNCH_CHECK(ctx);
  gcs_timing_begin(ctx, st, GCS_TIME_SURFEL_FIT);
  // groups of eight lanes / single threads stride over the occupied-cell list (a typical scan fills ~2,500 of 8,192 cells)
  double* mom = (double*)(ws + o_mom);
  // one list entry per group / thread for a single scan (latency), a quarter of that per unit for a batch (the list of a
  // typical scan is short and the batch fills the device anyway)
  const int spread = n_units >= 16 ? 4 : 1;
  const int mom_blocks = (G.n_cells * kFitLanes / spread + 127) / 128, fit_blocks = (G.n_cells / spread + 127) / 128;
  surfel_moments_kernel<<<dim3(mom_blocks, Hu), 128, 0, st>>>(pts, timestamps, weights, center, bucket, total, G, U, n_keys, occ_list,
                                                              n_occ, mom);
  GCS_LAUNCH_CHECK(ctx);
  surfel_fit_kernel<<<dim3(fit_blocks, Hu), 128, 0, st>>>(center, G, *cfg, F, occ_list, n_occ, mom);
  gcs_timing_end(ctx, st, GCS_TIME_SURFEL_FIT);
  GCS_LAUNCH_CHECK(ctx);
  surfel_select_kernel<<<Hu, 1024, (size_t)batch->n_surfel * sizeof(int), st>>>(F, G, *batch, cfg->eps_lift, out_n_valid, total, out_count,
                                                                                 n_keys);
  GCS_LAUNCH_CHECK(ctx);
  return GCS_OK;
}

int gcs_extract_lidar_surfels(gcs_ctx* ctx, void* stream, const double* pts, const double* timestamps,
                              const double* weights, int64_t n, const gcs_surfel_cfg* cfg, const gcs_meas_batch* batch,
                              int32_t* out_n_valid, int32_t* out_bucket, int32_t* out_count) {
  if (!ctx) return GCS_EINVAL;
  GCS_CHECK_CUDA(ctx, cudaSetDevice(ctx->device));
  const UnitStrides U = {0, 0, 0};
  return surfels_launch(ctx, (cudaStream_t)stream, pts, timestamps, weights, n, 1, U, cfg, batch, out_n_valid, out_bucket,
                        out_count, "extract_lidar_surfels");
}

int gcs_extract_lidar_surfels_batched(gcs_ctx* ctx, void* stream, const double* pts, const double* timestamps,
                                      const double* weights, int64_t n, int32_t n_units, int32_t timestamps_shared,
                                      const gcs_surfel_cfg* cfg, const gcs_meas_batch* batch, int32_t* out_n_valid) {
  if (!ctx) return GCS_EINVAL;
  GCS_CHECK_CUDA(ctx, cudaSetDevice(ctx->device));
  GCS_REQUIRE(ctx, n_units >= 1 && n_units <= 65535, "extract_lidar_surfels_batched: n_units=%d", n_units);
  const UnitStrides U = {3 * n, n, timestamps_shared ? 0 : n};
  return surfels_launch(ctx, (cudaStream_t)stream, pts, timestamps, weights, n, n_units, U, cfg, batch, out_n_valid, nullptr,
                        nullptr, "extract_lidar_surfels_batched");
}

}  // extern "C"
