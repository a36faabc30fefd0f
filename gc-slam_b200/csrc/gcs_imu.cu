// gcs_imu.cu -- the prologue that produces deskew's only non-point input: IMU window
// weights + fixed-cost preintegration over the scan window -> relative pose -> constant body twist xi_body.
//   smooth_window_weights               fl/backend/operators/imu_preintegration.py:19-43
//   preintegrate_imu_relative_pose_jax  fl/backend/operators/imu_preintegration.py:46-146   (lax.scan over M samples)
//   se3_log / _se3_V_inv                fl/common/geometry/se3_jax.py:178-256
//   glue (within-scan window, se3_log, rotation-only scale)   fl/backend/pipeline.py:436-483
//
// The reference walks the M (= 512) samples sequentially.  Here one CTA serves one hypothesis: every thread owns a
// contiguous run of samples, the orientation at the start of each run comes from an ordered prefix product of the
// per-run rotation increments (matrix product is associative), and velocity / position / mean sums are combined
// over the runs in index order by one thread per scalar -- a fixed order, so results are bit-identical run to run
// and within ~1e-15 of the sequential recurrence.  xi_body is written contiguously so that it can be handed to the
// bin path (gcs_bins_args.xi) without leaving the device.
#include "gcs_common.cuh"

namespace gcs {

namespace {

constexpr int kImuThreads = 128;
constexpr int kNPart = 17;   // per-run partials: dv(3) T pl(3) s_ab(3) s_awn(3) s_aw(3) ess

struct ImuParams {
  const double* stamps; const double* gyro; const double* accel;
  int64_t M;
  const double* params;   // (H, GCS_IMU_NPARAM)
  const double* weights_in;   // (H, M) or NULL: membership weights supplied by the caller instead of the window
  double* out; double* xi_out; double* weights_out;
};

__device__ __forceinline__ Mat3 mat3_load(const double* p) {
  Mat3 m;
#pragma unroll
  for (int k = 0; k < 9; ++k) m.m[k] = p[k];
  return m;
}
__device__ __forceinline__ void mat3_store(double* p, const Mat3& m) {
#pragma unroll
  for (int k = 0; k < 9; ++k) p[k] = m.m[k];
}

// _se3_V_inv (se3_jax.py:178-218)
__device__ inline Mat3 se3_V_inv(const double* phi) {
  const double theta_sq = phi[0] * phi[0] + phi[1] * phi[1] + phi[2] * phi[2];
  const double theta = sqrt(theta_sq);
  const bool small = theta < kSmallAngle;
  const double st = small ? 1.0 : theta;
  const double stsq = (theta_sq < kSmallAngle * kSmallAngle) ? 1.0 : theta_sq;
  double sn, cs;
  sincos(st, &sn, &cs);
  const double denom = 2.0 * st * sn + 1e-12;
  const double D = small ? 1.0 / 12.0 + theta_sq / 720.0 : (1.0 / stsq) - (1.0 + cs) / denom;
  Mat3 K, K2, V;
  skew_and_square(phi, K, K2);
#pragma unroll
  for (int i = 0; i < 9; ++i) V.m[i] = ((i % 4 == 0) ? 1.0 : 0.0) - 0.5 * K.m[i] + D * K2.m[i];
  return V;
}

struct ImuSample { double dt_eff, w; double a_body[3]; Mat3 dR; };

__device__ __forceinline__ void imu_sample(const ImuParams& P, const double* hp, int h, int64_t k, ImuSample& s) {
  const double sig = fmax(hp[12], 1e-6);
  const double t = P.stamps[k];
  const double a = (t - hp[13]) / sig, b = (hp[14] - t) / sig;
  const double w_raw = (1.0 / (1.0 + exp(-a))) * (1.0 / (1.0 + exp(-b)));
  s.w = P.weights_in ? P.weights_in[(int64_t)h * P.M + k] : w_raw * (1.0 - kWeightFloor) + kWeightFloor;
  double dt = (k + 1 < P.M) ? P.stamps[k + 1] - t : 0.0;
  dt = fmax(dt, 0.0);
  s.dt_eff = s.w * dt;
  if (P.gyro) {
    const double om[3] = {(P.gyro[3 * k] - hp[3]) * s.dt_eff, (P.gyro[3 * k + 1] - hp[4]) * s.dt_eff,
                          (P.gyro[3 * k + 2] - hp[5]) * s.dt_eff};
    s.dR = so3_exp(om);
#pragma unroll
    for (int c = 0; c < 3; ++c) s.a_body[c] = P.accel[3 * k + c] - hp[6 + c];
  }
}

__global__ void __launch_bounds__(kImuThreads) imu_twist_kernel(const ImuParams P) {
  __shared__ double sc[2][kImuThreads][9];
  __shared__ double part[kImuThreads][kNPart];
  __shared__ double fin[kNPart];   // static shared memory: 18 KB scan buffers + 17 KB partials
  const int h = blockIdx.x, i = threadIdx.x;
  const double* hp = P.params + (int64_t)h * GCS_IMU_NPARAM;
  const int64_t run = (P.M + kImuThreads - 1) / kImuThreads;
  const int64_t k0 = (int64_t)i * run, k1 = (k0 + run < P.M) ? k0 + run : P.M;

  // ---- pass 1: weights, rotation increments, per-run ordered product
  Mat3 Q = mat3_identity();
  for (int64_t k = k0; k < k1; ++k) {
    ImuSample s;
    imu_sample(P, hp, h, k, s);
    if (P.weights_out) P.weights_out[(int64_t)h * P.M + k] = s.w;
    if (P.gyro) Q = mat3_mul(Q, s.dR);
  }
  if (!P.gyro) return;   // weights only
  mat3_store(sc[0][i], Q);
  __syncthreads();
  int b = 0;
  for (int off = 1; off < kImuThreads; off <<= 1, b ^= 1) {
    Mat3 m = mat3_load(sc[b][i]);
    if (i >= off) m = mat3_mul(mat3_load(sc[b][i - off]), m);   // earlier runs on the left
    mat3_store(sc[b ^ 1][i], m);
    __syncthreads();
  }
  const Mat3 R0 = so3_exp(hp);
  Mat3 R = i > 0 ? mat3_mul(R0, mat3_load(sc[b][i - 1])) : R0;

  // ---- pass 2: accelerations in the world frame, per-run velocity / position / mean partials
  double dv[3] = {0, 0, 0}, pl[3] = {0, 0, 0}, sab[3] = {0, 0, 0}, sawn[3] = {0, 0, 0}, saw[3] = {0, 0, 0};
  double T = 0.0, ess = 0.0;
  for (int64_t k = k0; k < k1; ++k) {
    ImuSample s;
    imu_sample(P, hp, h, k, s);
    double awn[3];
    mat3_vec(R, s.a_body, awn);
    ess += s.w;
    T += s.dt_eff;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const double aw = awn[c] + hp[9 + c];
      sab[c] += s.a_body[c] * s.dt_eff;
      sawn[c] += awn[c] * s.dt_eff;
      saw[c] += aw * s.dt_eff;
      pl[c] = pl[c] + dv[c] * s.dt_eff + 0.5 * aw * (s.dt_eff * s.dt_eff);
      dv[c] += aw * s.dt_eff;
    }
    R = mat3_mul(R, s.dR);
  }
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    part[i][c] = dv[c]; part[i][4 + c] = pl[c]; part[i][7 + c] = sab[c]; part[i][10 + c] = sawn[c]; part[i][13 + c] = saw[c];
  }
  part[i][3] = T; part[i][16] = ess;
  __syncthreads();

  // ---- ordered combine: thread q owns scalar q (position threads carry the running velocity)
  if (i < 17) {
    double acc = 0.0, v = 0.0;
    if (i >= 4 && i < 7) {
      for (int r = 0; r < kImuThreads; ++r) {
        acc = acc + v * part[r][3] + part[r][i];
        v += part[r][i - 4];
      }
    } else {
      for (int r = 0; r < kImuThreads; ++r) acc += part[r][i];
    }
    fin[i] = acc;
  }
  __syncthreads();
  if (i != 0) return;

  // ---- relative pose in the start body frame, se3_log, rotation-only scale
  const Mat3 R_end = mat3_mul(R0, mat3_load(sc[b][kImuThreads - 1]));
  const Mat3 R0T = mat3_T(R0);
  const Mat3 dR = mat3_mul(R0T, R_end);
  double rotvec[3], p_body[3], v_body[3];
  so3_log(dR, rotvec);
  mat3_vec(R0T, &fin[4], p_body);
  mat3_vec(R0T, &fin[0], v_body);
  double* o = P.out + (int64_t)h * GCS_IMU_NOUT;
  const double denom = fmax(fin[3], 1e-12);
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    o[GCS_IMU_DELTA_POSE + c] = p_body[c]; o[GCS_IMU_DELTA_POSE + 3 + c] = rotvec[c];
    o[GCS_IMU_DELTA_P + c] = p_body[c]; o[GCS_IMU_DELTA_V + c] = v_body[c];
    o[GCS_IMU_A_BODY_MEAN + c] = fin[7 + c] / denom;
    o[GCS_IMU_A_WORLD_NOG_MEAN + c] = fin[10 + c] / denom;
    o[GCS_IMU_A_WORLD_MEAN + c] = fin[13 + c] / denom;
  }
#pragma unroll
  for (int c = 0; c < 9; ++c) o[GCS_IMU_DELTA_R + c] = dR.m[c];
  o[GCS_IMU_ESS] = fin[16];
  o[GCS_IMU_DT_EFF_SUM] = fin[3];
  // se3_log(delta_pose): phi = Log(Exp(rotvec)), rho = V(phi)^-1 t
  double phi[3], rho[3];
  so3_log(so3_exp(rotvec), phi);
  mat3_vec(se3_V_inv(phi), p_body, rho);
  const double ts = hp[15];
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    o[GCS_IMU_XI_BODY + c] = rho[c] * ts; o[GCS_IMU_XI_BODY + 3 + c] = phi[c];
    if (P.xi_out) { P.xi_out[(int64_t)h * 6 + c] = rho[c] * ts; P.xi_out[(int64_t)h * 6 + 3 + c] = phi[c]; }
  }
  o[38] = 0.0; o[39] = 0.0;
}

}  // namespace

}  // namespace gcs

using namespace gcs;

extern "C" int gcs_imu_scan_twist(gcs_ctx* ctx, void* stream, const double* stamps, const double* gyro, const double* accel,
                                  int64_t n_samples, const double* params, const double* weights_in, int n_hyp, double* out,
                                  double* xi_out, double* weights_out) {
  if (!ctx) return GCS_EINVAL;
  GCS_CHECK_CUDA(ctx, cudaSetDevice(ctx->device));
  GCS_REQUIRE(ctx, stamps && params, "gcs_imu_scan_twist: stamps / params are NULL");
  GCS_REQUIRE(ctx, n_samples >= 1 && n_hyp >= 1, "gcs_imu_scan_twist: need n_samples >= 1 and n_hyp >= 1 (got %lld, %d)",
              (long long)n_samples, n_hyp);
  GCS_REQUIRE(ctx, (gyro == nullptr) == (accel == nullptr), "gcs_imu_scan_twist: gyro and accel must both be set or both NULL");
  GCS_REQUIRE(ctx, gyro ? out != nullptr : weights_out != nullptr,
              "gcs_imu_scan_twist: out is NULL (or, for the weights-only form, weights_out is NULL)");
  ImuParams P;
  P.stamps = stamps; P.gyro = gyro; P.accel = accel; P.M = n_samples; P.params = params;
  P.weights_in = weights_in; P.out = out; P.xi_out = xi_out; P.weights_out = weights_out;
  imu_twist_kernel<<<n_hyp, kImuThreads, 0, (cudaStream_t)stream>>>(P);
  GCS_LAUNCH_CHECK(ctx);
  return GCS_OK;
}
