// gcs_bins_final.cu -- per-unit epilogue of the bin family (one small CTA per scan x hypothesis):
//   raw additive sums -> ScanBinStats (InvMass, centroid, covariance, PSD projection, kappa)      [a5, a6]
//   -> MatrixFisherRotation (3x3 Jacobi SVD, so3_log, PSD projection, scatter metrics)               [a7]
//   -> PlanarTranslationEvidence (per-bin 3x3 inverse, WLS solve, planarisation)                     [a8]
//   -> 22-D embedding of the LiDAR evidence and all certificate scalars.
// Also the bin-map kernels (additive update with rigid pushforward + forgetting; derived stats)    [a9]
#include "gcs_bins.cuh"

namespace gcs {

constexpr int kFinalThreads = 128;   // 64 bin threads + 64 helpers (gcs_bins_final.cu: thread roles)

struct FinalParams {
  int n_bins, n_hyp;
  // source A: raw sums (from the per-point kernels)
  const double* raw_sums;  // (U, raw_len) or NULL
  const double* raw_max;   // (U, kNMax) or NULL
  const double* mass;      // (S, kNMass) or NULL
  int64_t cap_total;
  // source B: statistics given directly (gcs_bin_evidence)
  gcs_bin_stats in;
  // outputs
  gcs_bin_stats out;
  double* cert_bc;  // (U, GCS_BC_NCERT) or NULL
  double* cert_st;  // (U, GCS_ST_NCERT) or NULL
  // evidence
  gcs_map_bin_stats map;
  const double* poses;  // (U,6) or NULL -> no evidence
  double* evidence;     // (U, GCS_EV_NREC)
  double* L22;
  double* h22;
  double eps_psd, eps_mass;
  // stand-alone operator modes
  int do_mf, do_pt;              // which evidence blocks to compute (fused path: both)
  int use_pose_val, use_R_hat;   // take pose / R_hat from the by-value fields below
  double pose_val[6];
  double R_hat[9];
  const double* map_centroid;    // (B,3)   optional: already-derived map stats (planar_translation_evidence API)
  const double* map_Sigma_c;     // (B,3,3)
};

__device__ inline void sym6_to_mat(const double* s, Mat3& M) {
  M(0, 0) = s[0]; M(0, 1) = s[1]; M(0, 2) = s[2];
  M(1, 0) = s[1]; M(1, 1) = s[3]; M(1, 2) = s[4];
  M(2, 0) = s[2]; M(2, 1) = s[4]; M(2, 2) = s[5];
}

// compute_scatter_metrics (matrix_fisher_evidence.py:83-147) -> 17 doubles
__device__ inline void scatter_metrics17(const Mat3& S, double N_total, double eps, double* o) {
  Mat3 T;
  const double invN = 1.0 / (N_total + eps);
  for (int i = 0; i < 9; ++i) T.m[i] = S.m[i] * invN;
  double w[3];
  Mat3 V;
  eigh3(T, w, V);
  double l[3] = {fmax(w[2], 0.0), fmax(w[1], 0.0), fmax(w[0], 0.0)};
  o[0] = l[0]; o[1] = l[1]; o[2] = l[2];
  for (int r = 0; r < 3; ++r) { o[3 + 3 * r] = V(r, 2); o[3 + 3 * r + 1] = V(r, 1); o[3 + 3 * r + 2] = V(r, 0); }
  const double inv1 = 1.0 / (l[0] + eps);
  o[12] = (l[0] - l[1]) * inv1;
  o[13] = (l[1] - l[2]) * inv1;
  o[14] = l[2] * inv1;
  o[15] = 1.0 - o[14];
  const double tot = l[0] + l[1] + l[2] + eps;
  const double p1 = l[0] / tot, p2 = l[1] / tot, p3 = l[2] / tot;
  const double ent = -(p1 * log(p1 + eps) + p2 * log(p2 + eps) + p3 * log(p3 + eps));
  o[16] = exp(ent);
}

constexpr int kRedW = 32;  // doubles per bin in the reduction scratch

// Thread roles (kFinalThreads = 128): threads 0..63 own one bin each; threads 64..127 are helpers that take the work
// that does not depend on the bin threads' critical path -- the map-side derived statistics of bin (tid - 64), the
// Matrix-Fisher tail (so3_log, PSD projection, certificates), the two scatter-metric records and the map-scatter
// eigenvalues of the planarisation -- so that the serial chain of a unit is
//   per-bin PSD projection -> 3x3 SVD -> per-bin 3x3 inverse -> WLS solve + PSD projection
// instead of the sum of every eigen-decomposition in the epilogue.
__global__ void __launch_bounds__(kFinalThreads) bins_finalize_kernel(const FinalParams P) {
  __shared__ double red[kMaxBins * kRedW];
  __shared__ double tot[kRedW];
  __shared__ double sR[9];
  __shared__ double sMap[kMaxBins * 12];   // per bin: map centroid (3) + Sigma_c (9), written by the helper threads
  __shared__ double sZs;
  __shared__ double sMF[21];               // SVD factors handed from thread 0 to the Matrix-Fisher helper
  __shared__ double sL[9];                 // projected translation information matrix
  const int u = blockIdx.x, tid = threadIdx.x, B = P.n_bins;
  const bool helper = tid >= kMaxBins;
  const int b = helper ? tid - kMaxBins : tid;
  const bool on = !helper && b < B;
  const double eps = P.eps_mass;

  double N = 0.0, sdir[3] = {0, 0, 0}, pbar[3] = {0, 0, 0};
  Mat3 S, Sig;
  for (int i = 0; i < 9; ++i) { S.m[i] = 0.0; Sig.m[i] = 0.0; }

  // ---------------- helpers: map derived stats of bin b (archive/bin_atlas.py:159-198), independent of the scan
  double mN_dir = 0, mN_pos = 0, mSd[3] = {0, 0, 0};
  Mat3 mS;
  for (int i = 0; i < 9; ++i) mS.m[i] = 0.0;
  if (P.evidence && b < B) {
    if (!helper) {
      if (P.map.N_dir) mN_dir = P.map.N_dir[b];
      if (P.map.N_pos) mN_pos = P.map.N_pos[b];
      for (int k = 0; k < 3; ++k)
        if (P.map.S_dir) mSd[k] = P.map.S_dir[3 * b + k];
      for (int k = 0; k < 9; ++k)
        if (P.map.S_scatter) mS.m[k] = P.map.S_scatter[9 * b + k];
    } else if (P.do_pt) {
      double* o = sMap + b * 12;
      if (P.map_centroid && P.map_Sigma_c) {
        for (int k = 0; k < 3; ++k) o[k] = P.map_centroid[3 * b + k];
        for (int k = 0; k < 9; ++k) o[3 + k] = P.map_Sigma_c[9 * b + k];
      } else {
        const double np_ = P.map.N_pos ? P.map.N_pos[b] : 0.0;
        const double invp = 1.0 / (np_ + eps + kF64Eps);
        double cen[3];
        for (int k = 0; k < 3; ++k) cen[k] = (P.map.sum_p ? P.map.sum_p[3 * b + k] : 0.0) * invp;
        Mat3 craw;
        for (int i = 0; i < 3; ++i)
          for (int j = 0; j < 3; ++j) craw(i, j) = (P.map.sum_ppT ? P.map.sum_ppT[9 * b + 3 * i + j] : 0.0) * invp - cen[i] * cen[j];
        const Mat3 Sc = psd_project3(craw, P.eps_psd, nullptr);
        for (int k = 0; k < 3; ++k) o[k] = cen[k];
        for (int k = 0; k < 9; ++k) o[3 + k] = Sc.m[k];
      }
    }
  }

  if (P.raw_sums) {
    // ---------------- ScanBinMomentMatch epilogue (binning.py:178-209)
    double delta = 0.0, eps_ratio = 0.0;
    if (on) {
      const double* row = P.raw_sums + (int64_t)u * raw_sums_len(B) + b * kRowLen;
      N = row[0];
      sdir[0] = row[1]; sdir[1] = row[2]; sdir[2] = row[3];
      sym6_to_mat(row + 4, S);
      const double sp[3] = {row[10], row[11], row[12]};
      Mat3 Spp, Scov;
      sym6_to_mat(row + 13, Spp);
      sym6_to_mat(row + 19, Scov);
      const double den = N + eps + kF64Eps;  // inv_mass_core: fl/common/primitives.py:195-212
      const double inv = 1.0 / den;
      eps_ratio = eps / den;
      for (int k = 0; k < 3; ++k) pbar[k] = sp[k] * inv;
      Mat3 raw;
      for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) raw(i, j) = (Spp(i, j) * inv - pbar[i] * pbar[j]) + Scov(i, j) * inv;
      double c6[6];
      Sig = psd_project3(raw, P.eps_psd, c6);
      delta = c6[0];
      const double Rbar = sqrt(sdir[0] * sdir[0] + sdir[1] * sdir[1] + sdir[2] * sdir[2]) * inv;
      const double kap = kappa_from_resultant(Rbar, kEpsR, 3.0, kKappaR0, kKappaTau);
      const int64_t ub = (int64_t)u * B + b;
      if (P.out.N) P.out.N[ub] = N;
      if (P.out.kappa) P.out.kappa[ub] = kap;
      for (int k = 0; k < 3; ++k) {
        if (P.out.s_dir) P.out.s_dir[ub * 3 + k] = sdir[k];
        if (P.out.p_bar) P.out.p_bar[ub * 3 + k] = pbar[k];
        if (P.out.sum_p) P.out.sum_p[ub * 3 + k] = sp[k];
      }
      for (int k = 0; k < 9; ++k) {
        if (P.out.S_scatter) P.out.S_scatter[ub * 9 + k] = S.m[k];
        if (P.out.Sigma_p) P.out.Sigma_p[ub * 9 + k] = Sig.m[k];
        if (P.out.sum_ppT) P.out.sum_ppT[ub * 9 + k] = Spp.m[k];
      }
    }
    if (!helper) {
      red[b * kRedW + 0] = on ? N : 0.0;
      red[b * kRedW + 1] = on ? N * N : 0.0;
      red[b * kRedW + 2] = on ? N / (N + eps) : 0.0;
      red[b * kRedW + 3] = on ? delta : 0.0;
      red[b * kRedW + 4] = on ? eps_ratio : 0.0;
    }
    __syncthreads();
    if (tid == kMaxBins) {   // first helper: certificate scalars of the statistics
      double sN = 0, sN2 = 0, sFr = 0, sD = 0, mR = 0;
      for (int k = 0; k < B; ++k) {
        sN += red[k * kRedW]; sN2 += red[k * kRedW + 1]; sFr += red[k * kRedW + 2]; sD += red[k * kRedW + 3];
        mR = fmax(mR, red[k * kRedW + 4]);
      }
      const double ess = sN * sN / (sN2 + eps);
      const double sup = sFr / (double)B;
      if (P.cert_st) {
        double* c = P.cert_st + (int64_t)u * GCS_ST_NCERT;
        c[GCS_ST_ESS] = ess; c[GCS_ST_SUPPORT_FRAC] = sup; c[GCS_ST_PSD_DELTA] = sD; c[GCS_ST_MASS_EPS_RATIO] = mR;
      }
      if (P.cert_bc) {
        double* c = P.cert_bc + (int64_t)u * GCS_BC_NCERT;
        const double* ex = P.raw_sums + (int64_t)u * raw_sums_len(B) + B * kRowLen;
        const int s = u / P.n_hyp;
        const double M = P.mass[s * kNMass + kMassAll], Ms = P.mass[s * kNMass + kMassSel];
        const double Qs = P.mass[s * kNMass + kMassSelSq];
        const double scale = M / (Ms + eps);
        const double g = scale / (M + eps);
        c[GCS_BC_RS_MASS_IN] = M;
        c[GCS_BC_RS_ESS] = 1.0 / (g * g * Qs + (double)P.cap_total * eps);
        c[GCS_BC_RS_MASS_SCALE] = scale;
        c[GCS_BC_DK_SUM_W_OUT] = ex[kExSumWdk];
        c[GCS_BC_DK_SUM_W_IN] = ex[kExSumWrs];
        // -sum r log(r+eps) = -sum r log r - B*eps per row to first order (see soft_assign_cert_kernel)
        c[GCS_BC_SA_ENTROPY_SUM] = (ex[kExEntLog] - ex[kExEntDot]) - ex[kExCount] * (double)B * eps;
        c[GCS_BC_SA_MAX_RESP] = P.raw_max[(int64_t)u * kNMax + kMxResp];
        c[GCS_BC_ST_ESS] = ess; c[GCS_BC_ST_SUPPORT_FRAC] = sup; c[GCS_BC_ST_PSD_DELTA] = sD;
        c[GCS_BC_ST_MASS_EPS_RATIO] = mR;
        for (int k = GCS_BC_ST_MASS_EPS_RATIO + 1; k < GCS_BC_NCERT; ++k) c[k] = 0.0;
      }
    }
    __syncthreads();   // red[] is rewritten below
  } else if (on) {
    const int64_t ub = (int64_t)u * B + b;
    N = P.in.N[ub];
    for (int k = 0; k < 3; ++k) {
      if (P.in.s_dir) sdir[k] = P.in.s_dir[ub * 3 + k];
      if (P.in.p_bar) pbar[k] = P.in.p_bar[ub * 3 + k];
    }
    for (int k = 0; k < 9; ++k) {
      if (P.in.S_scatter) S.m[k] = P.in.S_scatter[ub * 9 + k];
      if (P.in.Sigma_p) Sig.m[k] = P.in.Sigma_p[ub * 9 + k];
    }
  }
  if (!P.evidence) return;

  // ---------------- MatrixFisherRotation (matrix_fisher_evidence.py:155-256): per-bin terms, then column sums
  if (!helper) {
    const double w_b = sqrt(N * mN_dir + eps);
    const double sn = sqrt(sdir[0] * sdir[0] + sdir[1] * sdir[1] + sdir[2] * sdir[2]);
    const double mn = sqrt(mSd[0] * mSd[0] + mSd[1] * mSd[1] + mSd[2] * mSd[2]);
    const double us[3] = {sdir[0] / (sn + eps), sdir[1] / (sn + eps), sdir[2] / (sn + eps)};
    const double um[3] = {mSd[0] / (mn + eps), mSd[1] / (mn + eps), mSd[2] / (mn + eps)};
    const double Rs = sn * (1.0 / (N + eps)), Rm = mn * (1.0 / (mN_dir + eps));
    const double wf = w_b * (Rs * Rm);
    double* r = red + b * kRedW;
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) r[3 * i + j] = on ? wf * um[i] * us[j] : 0.0;
    for (int k = 0; k < 9; ++k) { r[9 + k] = on ? S.m[k] : 0.0; r[18 + k] = on ? mS.m[k] : 0.0; }
    r[27] = on ? N : 0.0; r[28] = on ? mN_dir : 0.0; r[29] = on ? wf : 0.0;
  }
  __syncthreads();
  if (tid < 30) {
    double a = 0.0;
    for (int k = 0; k < B; ++k) a += red[k * kRedW + tid];
    tot[tid] = a;
  }
  __syncthreads();
  double* ev = P.evidence + (int64_t)u * GCS_EV_NREC;
  double pose[6];
  for (int k = 0; k < 6; ++k) pose[k] = P.use_pose_val ? P.pose_val[k] : P.poses[(int64_t)u * 6 + k];

  // critical path: the SVD alone (thread 0); its consumers wait at the next barrier, the rest of the Matrix-Fisher
  // record is finished by helper threads while the bin threads go on with the translation evidence
  Mat3 mfU, mfV;
  double mfs[3] = {0, 0, 0};
  if (!P.do_mf) {
    if (tid < 9) sR[tid] = P.R_hat[tid];
    if (tid == 0) for (int k = 0; k < GCS_EV_T_WLS; ++k) ev[k] = 0.0;
  } else if (tid == 0) {
    Mat3 H;
    for (int k = 0; k < 9; ++k) H.m[k] = tot[k];
    svd3(H, mfU, mfs, mfV);
    Mat3 Vt = mat3_T(mfV);
    const double det = mat3_det(mat3_mul(mfU, Vt));
    const double sg = det > 0.0 ? 1.0 : (det < 0.0 ? -1.0 : 0.0);  // jnp.sign
    for (int r = 0; r < 3; ++r) mfU(r, 2) *= sg;
    const Mat3 Rmf = mat3_mul(mfU, Vt);
    for (int k = 0; k < 9; ++k) { ev[GCS_EV_R_MF + k] = Rmf.m[k]; sR[k] = P.use_R_hat ? P.R_hat[k] : Rmf.m[k]; }
    // hand the factors to the helper that finishes the record
    for (int k = 0; k < 9; ++k) { sMF[k] = Rmf.m[k]; sMF[9 + k] = mfV.m[k]; }
    for (int k = 0; k < 3; ++k) sMF[18 + k] = mfs[k];
  } else if (tid == kMaxBins && P.do_mf) {            // warp 2: the four serial tasks of this phase sit in four warps
    Mat3 St;
    for (int k = 0; k < 9; ++k) St.m[k] = tot[9 + k];
    scatter_metrics17(St, tot[27], eps, ev + GCS_EV_SCAN_METRICS);
  } else if (tid == kMaxBins + 32 && P.do_mf) {       // warp 3
    Mat3 Mt;
    for (int k = 0; k < 9; ++k) Mt.m[k] = tot[18 + k];
    scatter_metrics17(Mt, tot[28], eps, ev + GCS_EV_MAP_METRICS);
  }
  if (tid == 32 && P.do_pt) {                          // warp 1 (its bin threads idle until the barrier below)
    // self-adaptive z precision from the total map scatter (:579-592): needs only the column sums
    Mat3 Tm;
    const double Nd = tot[28] + eps;
    for (int k = 0; k < 9; ++k) Tm.m[k] = tot[18 + k] / Nd;
    double w[3];
    Mat3 Vd;
    eigh3(Tm, w, Vd);
    const double l1 = fmax(w[2], eps), l3 = fmax(w[0], 0.0);
    sZs = l3 / l1;
  }
  const double Neff_mf = tot[29];
  __syncthreads();   // sR, the SVD factors and sZs are published; tot[0..29] may be overwritten from here on

  // From here the two halves of the CTA run independently until the final barrier: warps 0-1 (bin threads) carry the
  // translation evidence and synchronise among themselves on named barrier 1; warp 2's first thread finishes the
  // Matrix-Fisher record.
  if (helper) {
    if (tid == kMaxBins && P.do_mf) {
      Mat3 Rmf, V;
      double sv[3];
      for (int k = 0; k < 9; ++k) { Rmf.m[k] = sMF[k]; V.m[k] = sMF[9 + k]; }
      for (int k = 0; k < 3; ++k) sv[k] = sMF[18 + k];
      const double ld[3] = {sv[1] + sv[2], sv[0] + sv[2], sv[0] + sv[1]};
      Mat3 Lraw;
      for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
          double a = 0.0;
          for (int k = 0; k < 3; ++k) a += V(i, k) * ld[k] * V(j, k);
          Lraw(i, j) = a;
        }
      Mat3 Rp = so3_exp(pose + 3);
      Mat3 Rerr = mat3_mul(mat3_T(Rp), Rmf);
      double dr[3];
      so3_log(Rerr, dr);
      double c6[6];
      Mat3 L = psd_project3(Lraw, P.eps_psd, c6);
      double hr[3];
      mat3_vec(L, dr, hr);
      const double nll = 0.5 * (dr[0] * hr[0] + dr[1] * hr[1] + dr[2] * hr[2]);
      const double Neff = Neff_mf;
      for (int k = 0; k < 9; ++k) ev[GCS_EV_L_ROT + k] = L.m[k];
      for (int k = 0; k < 3; ++k) { ev[GCS_EV_H_ROT + k] = hr[k]; ev[GCS_EV_DELTA_ROT + k] = dr[k]; ev[GCS_EV_SVD_S + k] = sv[k]; }
      const double smin = fmin(sv[0], fmin(sv[1], sv[2])), smax = fmax(sv[0], fmax(sv[1], sv[2]));
      ev[GCS_EV_MF_EIG_MIN] = smin; ev[GCS_EV_MF_EIG_MAX] = smax; ev[GCS_EV_MF_COND] = smax / (smin + eps);
      ev[GCS_EV_MF_NEAR_NULL] = (double)((sv[0] < eps) + (sv[1] < eps) + (sv[2] < eps));
      ev[GCS_EV_MF_NLL_PER_ESS] = nll / (Neff + eps);
      ev[GCS_EV_MF_DIR_SCORE] = sv[0] + sv[1] + sv[2];
      ev[GCS_EV_MF_PSD_DELTA] = c6[0];
      ev[GCS_EV_MF_MASS_EPS] = eps / (Neff + eps);
      ev[GCS_EV_MF_ROT_NLL] = nll;
      ev[GCS_EV_MF_N_EFF] = Neff;
    }
  } else if (!P.do_pt) {
    if (tid == 0) for (int k = GCS_EV_T_WLS; k < GCS_EV_MF_EIG_MIN; ++k) ev[k] = 0.0;
    if (tid == 0) for (int k = GCS_EV_PT_EIG_MIN; k < GCS_EV_NREC; ++k) ev[k] = 0.0;
  } else {
    // ---------------- PlanarTranslationEvidence (matrix_fisher_evidence.py:413-499, :502-671)
    {
      Mat3 R;
      for (int k = 0; k < 9; ++k) R.m[k] = sR[k];
      const double* mo = sMap + b * 12;
      double cen[3] = {0, 0, 0};
      Mat3 Sc;
      for (int k = 0; k < 9; ++k) Sc.m[k] = 0.0;
      if (on) {
        for (int k = 0; k < 3; ++k) cen[k] = mo[k];
        for (int k = 0; k < 9; ++k) Sc.m[k] = mo[3 + k];
      }
      double pr[3];
      mat3_vec(R, pbar, pr);
      const double tb[3] = {cen[0] - pr[0], cen[1] - pr[1], cen[2] - pr[2]};
      Mat3 RS = mat3_mul(mat3_mul(R, Sig), mat3_T(R));
      Mat3 Sg;
      for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) Sg(i, j) = (Sc(i, j) + RS(i, j)) + ((i == j) ? eps : 0.0);
      const double wb = sqrt(N * mN_pos + eps);
      Mat3 Wi = mat3_inv(Sg);
      double* r = red + b * kRedW;
      for (int k = 0; k < 9; ++k) { Wi.m[k] *= wb; r[k] = on ? Wi.m[k] : 0.0; }
      double hb[3];
      mat3_vec(Wi, tb, hb);
      for (int k = 0; k < 3; ++k) r[9 + k] = on ? hb[k] : 0.0;
      r[12] = on ? wb : 0.0;
    }
    asm volatile("bar.sync 1, 64;" ::: "memory");
    if (tid < 13) {
      double a = 0.0;
      for (int k = 0; k < B; ++k) a += red[k * kRedW + tid];
      tot[tid] = a;
    }
    asm volatile("bar.sync 1, 64;" ::: "memory");
    if (tid == 0) {
      const double zs = sZs;
      Mat3 Lf;
      for (int k = 0; k < 9; ++k) Lf.m[k] = tot[k];
      const double hf[3] = {tot[9], tot[10], tot[11]};
      Mat3 Lreg = Lf;
      Lreg(0, 0) += eps; Lreg(1, 1) += eps; Lreg(2, 2) += eps;
      double tw[3];
      mat3_solve(Lreg, hf, tw);
      const double mk[3] = {1.0, 1.0, zs};
      Mat3 Lraw;
      for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) Lraw(i, j) = Lf(i, j) * mk[i] * mk[j];
      const double Neff = tot[12];
      const double dt[3] = {tw[0] - pose[0], tw[1] - pose[1], tw[2] - pose[2]};
      double c6[6];
      Mat3 L = psd_project3(Lraw, P.eps_psd, c6);
      double ht[3];
      mat3_vec(L, dt, ht);
      const double nll = 0.5 * (dt[0] * ht[0] + dt[1] * ht[1] + dt[2] * ht[2]);
      for (int k = 0; k < 3; ++k) { ev[GCS_EV_T_WLS + k] = tw[k]; ev[GCS_EV_H_TRANS + k] = ht[k]; ev[GCS_EV_DELTA_TRANS + k] = dt[k]; }
      for (int k = 0; k < 9; ++k) { ev[GCS_EV_L_TRANS + k] = L.m[k]; sL[k] = L.m[k]; }
      ev[GCS_EV_XY_INFO] = 0.5 * (L(0, 0) + L(1, 1));
      ev[GCS_EV_Z_INFO] = L(2, 2);
      ev[GCS_EV_Z_SCALE] = zs;
      ev[GCS_EV_PT_NLL_PER_ESS] = nll / (Neff + eps);
      ev[GCS_EV_PT_PSD_DELTA] = c6[0];
      ev[GCS_EV_PT_MASS_EPS] = eps / (Neff + eps);
      ev[GCS_EV_PT_TRANS_NLL] = nll;
      ev[GCS_EV_PT_N_EFF] = Neff;
      for (int k = GCS_EV_PT_N_EFF + 1; k < GCS_EV_NREC; ++k) ev[k] = 0.0;
    }
  }
  __syncthreads();   // both halves done: every field of the evidence record except the conditioning scalars is written
  if (!P.do_pt) return;
  if (tid == kMaxBins) {
    // conditioning of the projected information matrix (eigvalsh of L, :640-655): nothing below depends on it
    Mat3 L;
    for (int k = 0; k < 9; ++k) L.m[k] = sL[k];
    double le[3];
    Mat3 Vl;
    eigh3(L, le, Vl);
    ev[GCS_EV_PT_EIG_MIN] = le[0]; ev[GCS_EV_PT_EIG_MAX] = le[2]; ev[GCS_EV_PT_COND] = le[2] / (le[0] + eps);
    ev[GCS_EV_PT_NEAR_NULL] = (double)((le[0] < eps) + (le[1] < eps) + (le[2] < eps));
  }
  // ---------------- build_combined_lidar_evidence_22d (:729-756)
  if (P.L22) {
    double* L22 = P.L22 + (int64_t)u * 22 * 22;
    for (int idx = tid; idx < 22 * 22; idx += kFinalThreads) {
      const int r = idx / 22, c = idx - r * 22;
      double v = 0.0;
      if (r < 3 && c < 3) v = sL[3 * r + c];
      else if (r >= 3 && r < 6 && c >= 3 && c < 6) v = ev[GCS_EV_L_ROT + 3 * (r - 3) + (c - 3)];
      L22[idx] = v;
    }
  }
  if (P.h22 && tid < 22) {
    double v = 0.0;
    if (tid < 3) v = ev[GCS_EV_H_TRANS + tid];
    else if (tid < 6) v = ev[GCS_EV_H_ROT + (tid - 3)];
    P.h22[(int64_t)u * 22 + tid] = v;
  }
}

// ---- bin map (a9) ------------------------------------------------------------------------------------
__global__ void map_bin_update_kernel(gcs_map_bin_stats map, const double* __restrict__ sN, const double* __restrict__ ssd,
                                      const double* __restrict__ sS, const double* __restrict__ ssp,
                                      const double* __restrict__ sspp, int n_bins, double t0, double t1, double t2,
                                      double r0, double r1, double r2, double gamma) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= n_bins) return;
  const double rv[3] = {r0, r1, r2};
  const double t[3] = {t0, t1, t2};
  Mat3 R = so3_exp(rv);
  Mat3 Rt = mat3_T(R);
  const double N = sN[b];
  double sd[3] = {ssd[3 * b], ssd[3 * b + 1], ssd[3 * b + 2]}, sp[3] = {ssp[3 * b], ssp[3 * b + 1], ssp[3 * b + 2]};
  Mat3 S, Spp;
  for (int k = 0; k < 9; ++k) { S.m[k] = sS[9 * b + k]; Spp.m[k] = sspp[9 * b + k]; }
  double Rsd[3], Rsp[3];
  mat3_vec(R, sd, Rsd);
  mat3_vec(R, sp, Rsp);
  Mat3 RS = mat3_mul(mat3_mul(R, S), Rt);
  Mat3 RP = mat3_mul(mat3_mul(R, Spp), Rt);
  map.N_dir[b] = gamma * (map.N_dir[b] + N);
  map.N_pos[b] = gamma * (map.N_pos[b] + N);
  for (int i = 0; i < 3; ++i) {
    map.S_dir[3 * b + i] = gamma * (map.S_dir[3 * b + i] + Rsd[i]);
    map.sum_p[3 * b + i] = gamma * (map.sum_p[3 * b + i] + (Rsp[i] + N * t[i]));
    for (int j = 0; j < 3; ++j) {
      map.S_scatter[9 * b + 3 * i + j] = gamma * (map.S_scatter[9 * b + 3 * i + j] + RS(i, j));
      const double inc = RP(i, j) + Rsp[i] * t[j] + t[i] * Rsp[j] + N * (t[i] * t[j]);
      map.sum_ppT[9 * b + 3 * i + j] = gamma * (map.sum_ppT[9 * b + 3 * i + j] + inc);
    }
  }
}

__global__ void map_bin_derived_kernel(gcs_map_bin_stats map, int n_bins, double eps_mass, double eps_psd,
                                       double* __restrict__ mu_dir, double* __restrict__ kappa,
                                       double* __restrict__ centroid, double* __restrict__ Sigma_c) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= n_bins) return;
  const double s[3] = {map.S_dir[3 * b], map.S_dir[3 * b + 1], map.S_dir[3 * b + 2]};
  const double nrm = sqrt(s[0] * s[0] + s[1] * s[1] + s[2] * s[2]);
  for (int k = 0; k < 3; ++k) mu_dir[3 * b + k] = s[k] / (nrm + eps_mass);
  const double invd = 1.0 / (map.N_dir[b] + eps_mass + kF64Eps);
  kappa[b] = kappa_from_resultant(nrm * invd, kEpsR, 3.0, kKappaR0, kKappaTau);
  const double invp = 1.0 / (map.N_pos[b] + eps_mass + kF64Eps);
  double c[3];
  for (int k = 0; k < 3; ++k) { c[k] = map.sum_p[3 * b + k] * invp; centroid[3 * b + k] = c[k]; }
  Mat3 raw;
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) raw(i, j) = map.sum_ppT[9 * b + 3 * i + j] * invp - c[i] * c[j];
  Mat3 Sc = psd_project3(raw, eps_psd, nullptr);
  for (int k = 0; k < 9; ++k) Sigma_c[9 * b + k] = Sc.m[k];
}

}  // namespace gcs

using namespace gcs;

static void zero_stats(gcs_bin_stats* s) { memset(s, 0, sizeof(*s)); }

int gcs_bins_finalize_impl(gcs_ctx* ctx, cudaStream_t st, const gcs_bins_args* a, int64_t cap_total, const double* mass,
                           const double* raw_sums, const double* raw_max) {
  FinalParams P;
  memset(&P, 0, sizeof(P));
  P.n_bins = a->n_bins; P.n_hyp = a->n_hyp;
  P.raw_sums = raw_sums; P.raw_max = raw_max; P.mass = mass; P.cap_total = cap_total;
  zero_stats(&P.in);
  P.out = a->stats;
  P.cert_bc = a->cert; P.cert_st = nullptr;
  P.eps_psd = a->eps_psd; P.eps_mass = a->eps_mass;
  if (a->evidence) {
    GCS_REQUIRE(ctx, a->map && a->poses, "gcs_bins: evidence requested but map/poses is NULL");
    GCS_REQUIRE(ctx, a->map->S_dir && a->map->S_scatter && a->map->N_dir && a->map->N_pos && a->map->sum_p && a->map->sum_ppT,
                "gcs_bins: map statistics pointer is NULL");
    P.map = *a->map; P.poses = a->poses; P.evidence = a->evidence; P.L22 = a->L22; P.h22 = a->h22;
    P.do_mf = 1; P.do_pt = 1;
  }
  bins_finalize_kernel<<<a->n_scans * a->n_hyp, kFinalThreads, 0, st>>>(P);
  GCS_LAUNCH_CHECK(ctx);
  return GCS_OK;
}

int gcs_stats_from_raw_impl(gcs_ctx* ctx, cudaStream_t st, const double* raw_sums, int n_units, int n_bins, double eps_psd,
                            double eps_mass, const gcs_bin_stats* out, double* cert_st) {
  FinalParams P;
  memset(&P, 0, sizeof(P));
  P.n_bins = n_bins; P.n_hyp = 1;
  P.raw_sums = raw_sums;
  P.out = *out;
  P.cert_st = cert_st;
  P.eps_psd = eps_psd; P.eps_mass = eps_mass;
  bins_finalize_kernel<<<n_units, kFinalThreads, 0, st>>>(P);
  GCS_LAUNCH_CHECK(ctx);
  return GCS_OK;
}

extern "C" {

int gcs_bin_evidence(gcs_ctx* ctx, void* stream, const gcs_bin_stats* scan, int n_units, int n_bins,
                     const gcs_map_bin_stats* map, const double* poses, double eps_psd, double eps_mass,
                     double* out_evidence, double* out_L22, double* out_h22) {
  if (!ctx) return GCS_EINVAL;
  GCS_CHECK_CUDA(ctx, cudaSetDevice(ctx->device));
  GCS_REQUIRE(ctx, scan && map && poses && out_evidence && n_units >= 1, "bin_evidence: bad args");
  GCS_REQUIRE(ctx, n_bins >= 1 && n_bins <= kMaxBins, "bin_evidence: n_bins=%d not in [1,%d]", n_bins, kMaxBins);
  GCS_REQUIRE(ctx, scan->N && scan->s_dir && scan->S_scatter && scan->p_bar && scan->Sigma_p, "bin_evidence: scan stats NULL");
  GCS_REQUIRE(ctx, map->S_dir && map->S_scatter && map->N_dir && map->N_pos && map->sum_p && map->sum_ppT,
              "bin_evidence: map stats NULL");
  FinalParams P;
  memset(&P, 0, sizeof(P));
  P.n_bins = n_bins; P.n_hyp = 1;
  P.in = *scan;
  P.map = *map; P.poses = poses; P.evidence = out_evidence; P.L22 = out_L22; P.h22 = out_h22;
  P.do_mf = 1; P.do_pt = 1;
  P.eps_psd = eps_psd; P.eps_mass = eps_mass;
  bins_finalize_kernel<<<n_units, kFinalThreads, 0, (cudaStream_t)stream>>>(P);
  GCS_LAUNCH_CHECK(ctx);
  return GCS_OK;
}

int gcs_matrix_fisher_rotation(gcs_ctx* ctx, void* stream, const double* scan_s_dir, const double* scan_S_scatter,
                               const double* scan_N, const double* map_S_dir, const double* map_S_scatter,
                               const double* map_N_dir, int n_bins, const double* pose6, double eps_psd, double eps_mass,
                               double* out_evidence) {
  if (!ctx) return GCS_EINVAL;
  GCS_CHECK_CUDA(ctx, cudaSetDevice(ctx->device));
  GCS_REQUIRE(ctx, scan_s_dir && scan_S_scatter && scan_N && map_S_dir && map_S_scatter && map_N_dir && pose6 && out_evidence,
              "matrix_fisher_rotation: NULL pointer");
  GCS_REQUIRE(ctx, n_bins >= 1 && n_bins <= kMaxBins, "matrix_fisher_rotation: n_bins=%d not in [1,%d]", n_bins, kMaxBins);
  FinalParams P;
  memset(&P, 0, sizeof(P));
  P.n_bins = n_bins; P.n_hyp = 1;
  P.in.N = (double*)scan_N; P.in.s_dir = (double*)scan_s_dir; P.in.S_scatter = (double*)scan_S_scatter;
  P.map.S_dir = (double*)map_S_dir; P.map.S_scatter = (double*)map_S_scatter; P.map.N_dir = (double*)map_N_dir;
  P.evidence = out_evidence;
  P.do_mf = 1; P.do_pt = 0; P.use_pose_val = 1;
  for (int k = 0; k < 6; ++k) P.pose_val[k] = pose6[k];
  P.eps_psd = eps_psd; P.eps_mass = eps_mass;
  bins_finalize_kernel<<<1, kFinalThreads, 0, (cudaStream_t)stream>>>(P);
  GCS_LAUNCH_CHECK(ctx);
  return GCS_OK;
}

int gcs_planar_translation(gcs_ctx* ctx, void* stream, const double* scan_p_bar, const double* scan_Sigma_p,
                           const double* scan_N, const double* map_centroid, const double* map_Sigma_c,
                           const double* map_N_pos, const double* map_S_scatter, const double* map_N_dir, int n_bins,
                           const double* R_hat, const double* t_pred, double eps_psd, double eps_mass,
                           double* out_evidence) {
  if (!ctx) return GCS_EINVAL;
  GCS_CHECK_CUDA(ctx, cudaSetDevice(ctx->device));
  GCS_REQUIRE(ctx, scan_p_bar && scan_Sigma_p && scan_N && map_centroid && map_Sigma_c && map_N_pos && map_S_scatter &&
                       map_N_dir && R_hat && t_pred && out_evidence, "planar_translation: NULL pointer");
  GCS_REQUIRE(ctx, n_bins >= 1 && n_bins <= kMaxBins, "planar_translation: n_bins=%d not in [1,%d]", n_bins, kMaxBins);
  FinalParams P;
  memset(&P, 0, sizeof(P));
  P.n_bins = n_bins; P.n_hyp = 1;
  P.in.N = (double*)scan_N; P.in.p_bar = (double*)scan_p_bar; P.in.Sigma_p = (double*)scan_Sigma_p;
  P.map.N_pos = (double*)map_N_pos; P.map.S_scatter = (double*)map_S_scatter; P.map.N_dir = (double*)map_N_dir;
  P.map_centroid = map_centroid; P.map_Sigma_c = map_Sigma_c;
  P.evidence = out_evidence;
  P.do_mf = 0; P.do_pt = 1; P.use_pose_val = 1; P.use_R_hat = 1;
  for (int k = 0; k < 3; ++k) { P.pose_val[k] = t_pred[k]; P.pose_val[3 + k] = 0.0; }
  for (int k = 0; k < 9; ++k) P.R_hat[k] = R_hat[k];
  P.eps_psd = eps_psd; P.eps_mass = eps_mass;
  bins_finalize_kernel<<<1, kFinalThreads, 0, (cudaStream_t)stream>>>(P);
  GCS_LAUNCH_CHECK(ctx);
  return GCS_OK;
}

int gcs_map_bin_update(gcs_ctx* ctx, void* stream, const gcs_map_bin_stats* map, const double* scan_N,
                       const double* scan_s_dir, const double* scan_S_scatter, const double* scan_sum_p,
                       const double* scan_sum_ppT, int n_bins, const double* pose6, int planar_z, double forgetting) {
  if (!ctx) return GCS_EINVAL;
  GCS_CHECK_CUDA(ctx, cudaSetDevice(ctx->device));
  GCS_REQUIRE(ctx, map && scan_N && scan_s_dir && scan_S_scatter && scan_sum_p && scan_sum_ppT && pose6 && n_bins >= 1,
              "map_bin_update: bad args");
  map_bin_update_kernel<<<(n_bins + 63) / 64, 64, 0, (cudaStream_t)stream>>>(
      *map, scan_N, scan_s_dir, scan_S_scatter, scan_sum_p, scan_sum_ppT, n_bins, pose6[0], pose6[1],
      planar_z ? 0.0 : pose6[2], pose6[3], pose6[4], pose6[5], forgetting);
  GCS_LAUNCH_CHECK(ctx);
  return GCS_OK;
}

int gcs_map_bin_derived(gcs_ctx* ctx, void* stream, const gcs_map_bin_stats* map, int n_bins, double eps_mass,
                        double eps_psd, double* mu_dir, double* kappa, double* centroid, double* Sigma_c) {
  if (!ctx) return GCS_EINVAL;
  GCS_CHECK_CUDA(ctx, cudaSetDevice(ctx->device));
  GCS_REQUIRE(ctx, map && mu_dir && kappa && centroid && Sigma_c && n_bins >= 1, "map_bin_derived: bad args");
  map_bin_derived_kernel<<<(n_bins + 63) / 64, 64, 0, (cudaStream_t)stream>>>(*map, n_bins, eps_mass, eps_psd, mu_dir, kappa,
                                                                              centroid, Sigma_c);
  GCS_LAUNCH_CHECK(ctx);
  return GCS_OK;
}

}  // extern "C"
