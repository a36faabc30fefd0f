// gcs_fusion.cu -- the step between the per-hypothesis LiDAR evidence and the hypothesis combine (SURVEY.md section 8f,
// rank 3: "batched 22-D evidence assembly, tempering beta, PSD projection, 6 x 6 eigvalsh"), for all K hypotheses of a
// scan in one launch so that none of them costs a host round trip.  One CTA per hypothesis:
//   step 9   raw evidence = IMU/odom evidence + LiDAR evidence, observability sentinels on the raw evidence,
//            closed-form power tempering beta                         fl/backend/pipeline.py:1038-1117
//            Fisher-derived excitation scaling of the prior           fl/backend/operators/excitation.py:15-64, pipeline.py:1119-1146
//   step 10  conditioning of the pose block (6 x 6 eigvalsh) and the fusion scale alpha
//                                                                      pipeline.py:1155-1193, fl/backend/operators/fusion.py:46-143
//   step 11  InfoFusionAdditive: L_post = PSD(L_prior + alpha L_ev)   fl/backend/operators/fusion.py:150-230,
//            domain_projection_psd_core                                fl/common/primitives.py:80-123
// The certificate-level inputs of the control laws (ess_total, excitation_total, nll_per_ess of the aggregated
// certificates) are host scalars in the reference too; they come in as one (K, 4) array.
#include "gcs_jacobi.cuh"

namespace gcs {

namespace {

constexpr int kFuD = 22;
constexpr int kIdxDt = 15, kIdxEx0 = 16, kIdxVel0 = 6, kIdxVel1 = 9, kPose = 6;

struct FuParams {
  const double* L_lidar; const double* h_lidar; const double* L_other; const double* h_other;
  const double* L_prior; const double* h_prior; const double* scal;
  gcs_fusion_cfg cfg;
  double* L_post; double* h_post; double* L_ev; double* h_ev; double* L_prior_s; double* h_prior_s; double* rec;
};

__device__ __forceinline__ double clip01(double x) { return fmin(fmax(x, 0.0), 1.0); }
__device__ __forceinline__ double nan_to_num(double x, double v) { return isfinite(x) ? x : v; }

__global__ void __launch_bounds__(kHbThreads) evidence_fusion_kernel(const FuParams P) {
  __shared__ double A[kHbMaxD * kHbLd], V[kHbMaxD * kHbLd], S[kHbMaxD * kHbLd], E[kHbMaxD * kHbLd];
  __shared__ JacobiScratch jsc;
  __shared__ double sred[kHbThreads];
  __shared__ double he[kFuD];
  __shared__ double sc[8];   // beta, a_dt, a_ex, alpha
  const int tid = threadIdx.x, k = blockIdx.x, D = kFuD;
  const int64_t mo = (int64_t)k * D * D, vo = (int64_t)k * D;
  const gcs_fusion_cfg& c = P.cfg;
  double* rec = P.rec + (int64_t)k * GCS_FU_NREC;

  // ---- step 9: raw evidence
  for (int e = tid; e < D * D; e += kHbThreads) {
    const double raw = (P.L_other ? P.L_other[mo + e] : 0.0) + P.L_lidar[mo + e];
    E[(e / D) * kHbLd + (e % D)] = raw;
  }
  if (tid < D) he[tid] = (P.h_other ? P.h_other[vo + tid] : 0.0) + P.h_lidar[vo + tid];
  __syncthreads();
  if (tid == 0) {
    double beta = 1.0, dt_asym = 0.0, z_to_xy = 0.0, ess_to_exc = 0.0;
    const double ess_total = P.scal ? P.scal[4 * k + GCS_FU_IN_ESS_TOTAL] : 0.0;
    const double exc_total = P.scal ? P.scal[4 * k + GCS_FU_IN_EXC_TOTAL] : 0.0;
    {
      // sentinels on the raw evidence (pipeline.py:1070-1087)
      double rp = 0.0, cp = 0.0, rv = 0.0, cv = 0.0;
      for (int j = 0; j < kPose; ++j) { rp += E[kIdxDt * kHbLd + j] * E[kIdxDt * kHbLd + j]; cp += E[j * kHbLd + kIdxDt] * E[j * kHbLd + kIdxDt]; }
      for (int j = kIdxVel0; j < kIdxVel1; ++j) { rv += E[kIdxDt * kHbLd + j] * E[kIdxDt * kHbLd + j]; cv += E[j * kHbLd + kIdxDt] * E[j * kHbLd + kIdxDt]; }
      const double dt_pose = sqrt(rp) + sqrt(cp), dt_vel = sqrt(rv) + sqrt(cv);
      dt_asym = clip01(fabs(dt_vel - dt_pose) / (dt_vel + dt_pose + c.eps_mass));
      z_to_xy = fabs(E[2 * kHbLd + 2]) / (0.5 * (fabs(E[0]) + fabs(E[1 * kHbLd + 1])) + c.eps_mass);
      ess_to_exc = ess_total / (exc_total + c.eps_mass);
    }
    if (!(c.flags & GCS_FU_SKIP_TEMPERING)) {
      // closed-form tempering (pipeline.py:1091-1102)
      const double s_z = z_to_xy / (z_to_xy + c.power_beta_z_c);
      const double s_exc = 1.0 / (1.0 + (ess_to_exc / c.power_beta_exc_c));
      const double s = clip01(dt_asym * s_z * s_exc);
      beta = c.power_beta_min + (1.0 - c.power_beta_min) * s;
      beta = fmin(fmax(beta, c.power_beta_min), 1.0);
    }
    sc[0] = beta;
    rec[GCS_FU_BETA] = beta; rec[GCS_FU_DT_ASYMMETRY] = dt_asym; rec[GCS_FU_Z_TO_XY] = z_to_xy; rec[GCS_FU_ESS_TO_EXC] = ess_to_exc;
  }
  __syncthreads();
  const double beta = sc[0];
  for (int e = tid; e < D * D; e += kHbThreads) {
    const int i = e / D, j = e % D;
    const double v = (c.flags & GCS_FU_SKIP_TEMPERING) ? E[i * kHbLd + j] : beta * E[i * kHbLd + j];
    E[i * kHbLd + j] = v;
    if (P.L_ev) P.L_ev[mo + e] = v;
  }
  __syncthreads();
  if (tid < D) {
    if (!(c.flags & GCS_FU_SKIP_TEMPERING)) he[tid] = beta * he[tid];
    if (P.h_ev) P.h_ev[vo + tid] = he[tid];
  }

  // ---- excitation scaling of the prior (excitation.py:15-64)
  if (tid == 0) {
    double a_dt = 1.0, a_ex = 1.0, s_dt = 0.0, s_ex = 0.0;
    if (!(c.flags & GCS_FU_SKIP_PRIOR_SCALING)) {
      const double e_dt = E[kIdxDt * kHbLd + kIdxDt], pi_dt = P.L_prior[mo + kIdxDt * D + kIdxDt];
      double e_ex = 0.0, pi_ex = 0.0;
      for (int j = kIdxEx0; j < D; ++j) { e_ex += E[j * kHbLd + j]; pi_ex += P.L_prior[mo + j * D + j]; }
      s_dt = e_dt / (e_dt + pi_dt + c.exc_eps);
      s_ex = e_ex / (e_ex + pi_ex + c.exc_eps);
      a_dt = 1.0 - s_dt; a_ex = 1.0 - s_ex;
    }
    sc[1] = a_dt; sc[2] = a_ex;
    rec[GCS_FU_S_DT] = s_dt; rec[GCS_FU_S_EX] = s_ex;
  }
  __syncthreads();
  {
    const double a_dt = sc[1], a_ex = sc[2];
    const bool scale = !(c.flags & GCS_FU_SKIP_PRIOR_SCALING);
    // S <- scaled prior, in the reference's order: row dt, column dt, rows ex, columns ex
    for (int e = tid; e < D * D; e += kHbThreads) {
      const int i = e / D, j = e % D;
      double v = P.L_prior[mo + e];
      if (scale) {
        if (i == kIdxDt) v = a_dt * v;
        if (j == kIdxDt) v = a_dt * v;
        if (i >= kIdxEx0) v = a_ex * v;
        if (j >= kIdxEx0) v = a_ex * v;
      }
      S[i * kHbLd + j] = v;
      if (P.L_prior_s) P.L_prior_s[mo + e] = v;
    }
  }

  // ---- step 10: pose-block conditioning (pipeline.py:1155-1177)
  for (int e = tid; e < kPose * kPose; e += kHbThreads) {
    const int i = e / kPose, j = e % kPose;
    A[i * kHbLd + j] = nan_to_num(0.5 * (E[i * kHbLd + j] + E[j * kHbLd + i]), 0.0);
    V[i * kHbLd + j] = (i == j) ? 1.0 : 0.0;
  }
  __syncthreads();
  cta_jacobi_eigh(A, V, kPose, jsc, sred);
  if (tid == 0) {
    const double eps = c.eps_psd;
    double emin = 1e300, emax = -1e300, nn = 0.0;
    for (int m = 0; m < kPose; ++m) {
      const double ev = A[m * kHbLd + m];
      nn += (ev <= eps) ? 1.0 : 0.0;
      const double v = fmax(nan_to_num(ev, eps), eps);
      emin = fmin(emin, v); emax = fmax(emax, v);
    }
    const double cond = emax / emin;
    rec[GCS_FU_POSE_EIG_MIN] = emin; rec[GCS_FU_POSE_EIG_MAX] = emax; rec[GCS_FU_POSE_COND] = cond; rec[GCS_FU_POSE_NEAR_NULL] = nn;
    // fusion scale (fusion.py:76-113)
    double alpha = c.alpha_override, quality = 0.0;
    if (!(c.flags & GCS_FU_ALPHA_GIVEN)) {
      const double ess = P.scal[4 * k + GCS_FU_IN_ESS_TOTAL], exc = P.scal[4 * k + GCS_FU_IN_EXC_TOTAL];
      const double nll = P.scal[4 * k + GCS_FU_IN_NLL_PER_ESS];
      const double dt_asym = rec[GCS_FU_DT_ASYMMETRY], z = rec[GCS_FU_Z_TO_XY];
      const double cond_q = c.c0_cond / (cond + c.c0_cond);
      const double supp_q = ess / (ess + 1.0);
      const double mis_q = exp(-nll);
      const double dt_q = clip01(dt_asym);
      const double z_q = clip01(z / (z + 1.0));
      const double exc_q = clip01(exc / (exc + 1.0));
      const double base = sqrt(cond_q * supp_q);
      quality = base * mis_q * dt_q * z_q * exc_q * clip01(sc[0]);
      const double alpha_raw = c.alpha_min + (c.alpha_max - c.alpha_min) * quality;
      alpha = fmin(fmax(alpha_raw, c.alpha_min), c.alpha_max);
    }
    sc[3] = alpha;
    rec[GCS_FU_ALPHA] = alpha; rec[GCS_FU_QUALITY] = quality;
  }
  __syncthreads();

  // ---- step 11: additive fusion + DomainProjectionPSD (fusion.py:186-195, primitives.py:80-123)
  const double alpha = sc[3];
  for (int e = tid; e < D * D; e += kHbThreads) {
    const int i = e / D, j = e % D;
    A[i * kHbLd + j] = S[i * kHbLd + j] + alpha * E[i * kHbLd + j];    // L_post_raw
  }
  if (tid < D) {
    double hp = P.h_prior[vo + tid];
    if (!(c.flags & GCS_FU_SKIP_PRIOR_SCALING)) {
      if (tid == kIdxDt) hp = sc[1] * hp;
      if (tid >= kIdxEx0) hp = sc[2] * hp;
    }
    if (P.h_prior_s) P.h_prior_s[vo + tid] = hp;
    P.h_post[vo + tid] = hp + alpha * he[tid];
  }
  __syncthreads();
  double part = 0.0, tr_prior = 0.0;
  for (int e = tid; e < D * D; e += kHbThreads) {
    const int i = e / D, j = e % D;
    const double m = 0.5 * (A[i * kHbLd + j] + A[j * kHbLd + i]);
    const double d = m - A[i * kHbLd + j];
    part += d * d;
    if (i == j) tr_prior += S[i * kHbLd + i];
    E[i * kHbLd + j] = m;      // E is free now: keeps M_sym for the projection delta
    V[i * kHbLd + j] = (i == j) ? 1.0 : 0.0;
  }
  const double sym_delta = sqrt(hb_block_sum(part, sred));
  const double trace_prior = hb_block_sum(tr_prior, sred);
  for (int e = tid; e < D * D; e += kHbThreads) A[(e / D) * kHbLd + (e % D)] = E[(e / D) * kHbLd + (e % D)];
  __syncthreads();
  cta_jacobi_eigh(A, V, D, jsc, sred);
  part = 0.0;
  double tr_post = 0.0;
  for (int e = tid; e < D * D; e += kHbThreads) {
    const int i = e / D, j = e % D;
    double acc = 0.0;
    for (int m = 0; m < D; ++m) acc += V[i * kHbLd + m] * fmax(A[m * kHbLd + m], c.eps_psd) * V[j * kHbLd + m];
    P.L_post[mo + e] = acc;
    const double d = acc - E[i * kHbLd + j];
    part += d * d;
    if (i == j) tr_post += acc;
  }
  const double pd = hb_block_sum(part, sred);
  const double trace_post = hb_block_sum(tr_post, sred);
  if (tid == 0) {
    double emin = 1e300, emax = -1e300, nn = 0.0;
    for (int m = 0; m < D; ++m) {
      const double v = fmax(A[m * kHbLd + m], c.eps_psd);
      emin = fmin(emin, v); emax = fmax(emax, v);
      nn += (v < 10.0 * c.eps_psd) ? 1.0 : 0.0;
    }
    rec[GCS_FU_PSD_PROJECTION_DELTA] = sqrt(pd); rec[GCS_FU_PSD_SYM_DELTA] = sym_delta;
    rec[GCS_FU_POST_EIG_MIN] = emin; rec[GCS_FU_POST_EIG_MAX] = emax; rec[GCS_FU_POST_COND] = emax / emin;
    rec[GCS_FU_POST_NEAR_NULL] = nn;
    rec[GCS_FU_TRACE_INCREASE] = trace_post - trace_prior;
  }
}

}  // namespace

}  // namespace gcs

using namespace gcs;

extern "C" int gcs_evidence_fusion(gcs_ctx* ctx, void* stream, const double* L_lidar, const double* h_lidar, const double* L_other,
                                   const double* h_other, const double* L_prior, const double* h_prior, const double* cert_scalars,
                                   int n_hyp, int dim, const gcs_fusion_cfg* cfg, double* L_post, double* h_post,
                                   double* L_evidence, double* h_evidence, double* L_prior_scaled, double* h_prior_scaled, double* rec) {
  if (!ctx) return GCS_EINVAL;
  GCS_CHECK_CUDA(ctx, cudaSetDevice(ctx->device));
  GCS_REQUIRE(ctx, L_lidar && h_lidar && L_prior && h_prior && cfg && L_post && h_post && rec, "gcs_evidence_fusion: NULL pointer");
  GCS_REQUIRE(ctx, (L_other == nullptr) == (h_other == nullptr), "gcs_evidence_fusion: L_other and h_other must both be set or both NULL");
  GCS_REQUIRE(ctx, dim == kFuD, "gcs_evidence_fusion: dimension %d (the state slices of the control laws are those of the 22-D chart)", dim);
  GCS_REQUIRE(ctx, n_hyp >= 1, "gcs_evidence_fusion: need at least one hypothesis (got %d)", n_hyp);
  const bool needs_scal = !(cfg->flags & GCS_FU_SKIP_TEMPERING) || !(cfg->flags & GCS_FU_ALPHA_GIVEN);
  GCS_REQUIRE(ctx, cert_scalars || !needs_scal, "gcs_evidence_fusion: cert_scalars is NULL but tempering / the fusion scale need it");
  FuParams P;
  P.L_lidar = L_lidar; P.h_lidar = h_lidar; P.L_other = L_other; P.h_other = h_other; P.L_prior = L_prior; P.h_prior = h_prior;
  P.scal = cert_scalars; P.cfg = *cfg;
  if (!cert_scalars) P.cfg.flags |= GCS_FU_SKIP_TEMPERING | GCS_FU_ALPHA_GIVEN;
  P.L_post = L_post; P.h_post = h_post; P.L_ev = L_evidence; P.h_ev = h_evidence; P.L_prior_s = L_prior_scaled; P.h_prior_s = h_prior_scaled;
  P.rec = rec;
  evidence_fusion_kernel<<<n_hyp, kHbThreads, 0, (cudaStream_t)stream>>>(P);
  GCS_LAUNCH_CHECK(ctx);
  return GCS_OK;
}
