// gcs_jacobi.cuh -- symmetric eigenproblem of a small (D <= 32) matrix by one CTA of 256 threads: parallel-ordered
// cyclic Jacobi in shared memory.  The D/2 disjoint (p, q) pairs of a round-robin step get their rotations from the
// current matrix, then all threads apply the column rotations (A <- A J, V <- V J) and the row rotations (A <- J^T A);
// D - 1 steps make a sweep; at most kHbSweeps sweeps, ended as soon as the off-diagonal mass is below 1e-30 of the
// squared Frobenius norm (quadratic convergence: 6-8 sweeps; the count depends on the data only, so reruns stay
// bit-identical).  Used by the PSD projections of the hypothesis combine and of the evidence fusion
// (domain_projection_psd_core, fl/common/primitives.py:80-123) and by the 6 x 6 pose-block eigvalsh of the fusion step.
#pragma once
#include "gcs_common.cuh"

namespace gcs {

constexpr int kHbThreads = 256;
constexpr int kHbMaxD = 32;
constexpr int kHbLd = kHbMaxD + 1;   // padded leading dimension (bank conflicts)
constexpr int kHbSweeps = 12;        // upper bound; the iteration stops when the off-diagonal mass is below float64 resolution

// fixed-order sum of one value per thread over the CTA (shuffle tree per warp, then over the warps); result in all threads
__device__ __forceinline__ double hb_block_sum(double v, double* sred) {
  v = warp_sum(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) sred[threadIdx.x >> 5] = v;
  __syncthreads();
  double t = 0.0;
#pragma unroll
  for (int w = 0; w < kHbThreads / 32; ++w) t += sred[w];
  return t;
}

// round-robin tournament on n (even) players: step s in [0, n-1), pair t in [0, n/2) -> (p, q)
__device__ __forceinline__ void rr_pair(int n, int s, int t, int& p, int& q) {
  const int m = n - 1;
  int a = (t == 0) ? m : (s + t) % m;
  int b = (s + m - t) % m;
  p = a < b ? a : b;
  q = a < b ? b : a;
}

// scratch of the rotation parameters of one round-robin step
struct JacobiScratch {
  double rc[kHbMaxD / 2], rs[kHbMaxD / 2];
  int rp[kHbMaxD / 2], rq[kHbMaxD / 2];
};

// A (D x D, leading dimension kHbLd, symmetric) -> eigenvalues on its diagonal; V must hold the identity on entry and
// holds the eigenvectors (columns) on exit.  All kHbThreads threads of the CTA must call.
__device__ inline void cta_jacobi_eigh(double* A, double* V, int D, JacobiScratch& J, double* sred) {
  const int tid = threadIdx.x;
  double* rc = J.rc; double* rs = J.rs; int* rp = J.rp; int* rq = J.rq;
  const int n = (D + 1) & ~1;      // players (a dummy one when D is odd)
  const int half = n / 2;
  for (int sweep = 0; sweep < kHbSweeps; ++sweep) {
    // convergence: off-diagonal sum of squares against the squared Frobenius norm (fixed-order reduction)
    double off = 0.0, fro = 0.0;
    for (int e = tid; e < D * D; e += kHbThreads) {
      const int i = e / D, j = e % D;
      const double v = A[i * kHbLd + j];
      fro += v * v;
      if (i != j) off += v * v;
    }
    const double off_sum = hb_block_sum(off, sred), fro_sum = hb_block_sum(fro, sred);
    if (off_sum <= 1e-30 * fro_sum) break;
    for (int s = 0; s < n - 1; ++s) {
      if (tid < half) {
        int p, q;
        rr_pair(n, s, tid, p, q);
        double c = 1.0, sn = 0.0;
        if (q < D) {
          const double apq = A[p * kHbLd + q];
          if (apq != 0.0) {
            const double app = A[p * kHbLd + p], aqq = A[q * kHbLd + q];
            const double theta = (aqq - app) / (2.0 * apq);
            const double t = (theta >= 0.0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
            c = 1.0 / sqrt(t * t + 1.0);
            sn = t * c;
            if (!(fabs(theta) < 1e300)) { c = 1.0; sn = 0.0; }   // apq negligible against the diagonal gap
          }
        } else {
          q = p;   // pair with the dummy player: identity
        }
        rp[tid] = p; rq[tid] = q; rc[tid] = c; rs[tid] = sn;
      }
      __syncthreads();
      // columns: A <- A J, V <- V J   (one thread per (row, pair))
      for (int e = tid; e < D * half; e += kHbThreads) {
        const int i = e / half, t = e - i * half;
        const int p = rp[t], q = rq[t];
        if (p != q) {
          const double c = rc[t], sn = rs[t];
          const double aip = A[i * kHbLd + p], aiq = A[i * kHbLd + q];
          A[i * kHbLd + p] = c * aip - sn * aiq;
          A[i * kHbLd + q] = sn * aip + c * aiq;
          const double vip = V[i * kHbLd + p], viq = V[i * kHbLd + q];
          V[i * kHbLd + p] = c * vip - sn * viq;
          V[i * kHbLd + q] = sn * vip + c * viq;
        }
      }
      __syncthreads();
      // rows: A <- J^T A
      for (int e = tid; e < D * half; e += kHbThreads) {
        const int j = e / half, t = e - j * half;
        const int p = rp[t], q = rq[t];
        if (p != q) {
          const double c = rc[t], sn = rs[t];
          const double apj = A[p * kHbLd + j], aqj = A[q * kHbLd + j];
          A[p * kHbLd + j] = c * apj - sn * aqj;
          A[q * kHbLd + j] = sn * apj + c * aqj;
        }
      }
      __syncthreads();
    }
  }

}

}  // namespace gcs
