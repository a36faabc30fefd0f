// gcs_assoc.cuh -- the pair cost and the scaling power of the OT association, shared by the fused association kernels
// (gcs_prims_map.cu) and the stand-alone cost / Sinkhorn entries (gcs_assoc_ops.cu).
#pragma once
#include "gcs_common.cuh"

namespace gcs {

// A_vmf(k) = log(4 pi) + log sinh k - log k with the reference's three branches (primitive_association.py:141-149)
__device__ __forceinline__ double A_vmf(double k, double eps) {
  k = fmax(k, eps);
  double ls;
  if (k > 20.0) ls = k - log(2.0);
  else if (k >= 1e-2) ls = log(sinh(k));
  else ls = log(k + (k * k * k) / 6.0);
  return log(4.0 * 3.141592653589793) + ls - log(k);
}

// cost of (measurement, map entry)   (primitive_association.py:152-197); A_k1 / A_k2 = A_vmf of the two concentrations
// (they depend on one side only: callers that see an entry many times compute them once)
__device__ __forceinline__ double pair_cost_pre(const double* mp, const double* md, double mk, double A_k1, const double* vp,
                                                const double* vd, double vk, double A_k2, double beta, double eig_min = 1e-12) {
  const double d0 = mp[0] - vp[0], d1 = mp[1] - vp[1], d2 = mp[2] - vp[2];
  const double d_pos = d0 * d0 + d1 * d1 + d2 * d2;
  const double e0 = mk * md[0] + vk * vd[0], e1 = mk * md[1] + vk * vd[1], e2 = mk * md[2] + vk * vd[2];
  const double km = 0.5 * sqrt(e0 * e0 + e1 * e1 + e2 * e2);
  const double A_km = A_vmf(fmax(km, eig_min), eig_min);
  const double bc = exp(A_km - 0.5 * (A_k1 + A_k2));
  double d_dir = fmax(0.0, 1.0 - bc);
  if (!(mk > 0.0 && vk > 0.0)) d_dir = 0.0;
  return d_pos + beta * d_dir;
}
__device__ __forceinline__ double pair_cost(const double* mp, const double* md, double mk, double A_k1, const double* vp,
                                            const double* vd, double vk, double beta, double eig_min = 1e-12) {
  return pair_cost_pre(mp, md, mk, A_k1, vp, vd, vk, A_vmf(fmax(vk, eig_min), eig_min), beta, eig_min);
}

// x^y for the Sinkhorn scalings (x >= 0, y in (0, 1)): 0^y = 0 as jnp's power gives.  exp(y log x): the scalings are
// smooth in x, the ~|y log x| ulp of this form are far inside the tolerance, and it is ~3x shorter than the correctly
// rounded pow() on the serial path of every iteration.
__device__ __forceinline__ double pow_pos(double x, double y) { return x > 0.0 ? exp(y * log(x)) : 0.0; }

}  // namespace gcs
