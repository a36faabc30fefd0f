// gcs_ingest.cu -- the step in front of the LiDAR evidence path: PointCloud2 wire bytes
// (VLP-16 layout) -> SoA (points, timestamps, weights, ring, tag) in the base frame, on the device.
//   parse_pointcloud2_vlp16   fl/backend/backend_node.py:377-468
//   p_base = R p_lidar + t    fl/backend/backend_node.py:1682-1684
// The host uploads 22 bytes per point (the message payload) instead of the 42 bytes per point of five float64 / uint8
// arrays, and the NumPy structured-array decode, the two sigmoids and the 3x3 transform leave the host.
//
// pc2_parse_kernel: a CTA stages the payload of 256 points into shared memory with aligned 16-byte loads (any
// point_step, any field offsets), then one thread decodes one point.  The reference's data-dependent unit switch
// ("if any per-point time > 1e6 the field is in nanoseconds", :440-443) is an integer flag per message, applied by
// pc2_finish_kernel only when it is set.
#include "gcs_common.cuh"

namespace gcs {

namespace {

constexpr int kParseThreads = 256;
constexpr int kMaxPointStep = 256;

struct Pc2Params {
  const uint8_t* data;
  int64_t n_points;          // per message
  gcs_pc2_layout lay;
  const double* header_stamp;
  double R[9], tb[3];
  int has_tf;
  double sentinel, sigma, min_r, max_r, wfloor;
  double* pts; double* t; double* w; uint8_t* ring; uint8_t* tag;
  int* flags;                // (n_msgs, 2): [any t_raw > 1e6, number of non-finite coordinates]
};

// little-endian scalar of a sensor_msgs/PointField datatype, as float64 (what np.asarray(..., dtype=float64) yields)
__device__ __forceinline__ double load_field(const uint8_t* p, int type) {
  switch (type) {
    case 1: return (double)(int8_t)p[0];
    case 2: return (double)p[0];
    case 3: return (double)(int16_t)((uint16_t)p[0] | ((uint16_t)p[1] << 8));
    case 4: return (double)((uint16_t)p[0] | ((uint16_t)p[1] << 8));
    case 5: return (double)(int32_t)((uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24));
    case 6: return (double)((uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24));
    case 7: return (double)__uint_as_float((uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24));
    default: {
      uint64_t v = 0;
#pragma unroll
      for (int k = 0; k < 8; ++k) v |= (uint64_t)p[k] << (8 * k);
      return __longlong_as_double((long long)v);
    }
  }
}
// np.asarray(arr["ring"], dtype=np.uint8): integer value modulo 256
__device__ __forceinline__ uint8_t load_ring(const uint8_t* p) { return p[0]; }   // little-endian integer: low byte

__device__ __forceinline__ double nan_to_num(double v, double s, int& n_bad) {
  if (v != v) { ++n_bad; return s; }
  if (isinf(v)) { ++n_bad; return v > 0.0 ? s : -s; }
  return v;
}

__global__ void __launch_bounds__(kParseThreads) pc2_parse_kernel(const Pc2Params P) {
  extern __shared__ __align__(16) uint8_t stage[];
  const int m = blockIdx.y;
  const int step = P.lay.point_step;
  const int64_t p0 = (int64_t)blockIdx.x * kParseThreads;
  const int64_t n_here = (P.n_points - p0) < kParseThreads ? (P.n_points - p0) : kParseThreads;
  const int64_t byte0 = ((int64_t)m * P.n_points + p0) * step;
  const int64_t byte1 = byte0 + n_here * step;
  const int64_t a0 = byte0 & ~(int64_t)15;
  const int64_t a1 = (byte1 + 15) & ~(int64_t)15;       // the caller pads the buffer to a multiple of 16 bytes
  const int shift = (int)(byte0 - a0);
  for (int64_t k = threadIdx.x; k < (a1 - a0) / 16; k += kParseThreads)
    reinterpret_cast<uint4*>(stage)[k] = __ldg(reinterpret_cast<const uint4*>(P.data + a0) + k);
  __syncthreads();
  int n_bad = 0;
  bool big_t = false;
  if (threadIdx.x < n_here) {
    const uint8_t* rec = stage + shift + threadIdx.x * step;
    const double x = nan_to_num(load_field(rec + P.lay.off_x, P.lay.type_x), P.sentinel, n_bad);
    const double y = nan_to_num(load_field(rec + P.lay.off_y, P.lay.type_y), P.sentinel, n_bad);
    const double z = nan_to_num(load_field(rec + P.lay.off_z, P.lay.type_z), P.sentinel, n_bad);
    const uint8_t rg = load_ring(rec + P.lay.off_ring);
    double tt;
    if (P.lay.off_time >= 0) {
      tt = load_field(rec + P.lay.off_time, P.lay.type_time);
      big_t = tt > 1.0e6;
    } else {
      tt = P.header_stamp[m];
    }
    // range sigmoid window (backend_node.py:449-460)
    const double dist = sqrt(x * x + y * y + z * z);
    const double a = (dist - P.min_r) / P.sigma, b = (P.max_r - dist) / P.sigma;
    const double w_raw = (1.0 / (1.0 + exp(-a))) * (1.0 / (1.0 + exp(-b)));
    const double wv = w_raw * (1.0 - P.wfloor) + P.wfloor;
    double q0 = x, q1 = y, q2 = z;
    if (P.has_tf) {
      q0 = (P.R[0] * x + P.R[1] * y + P.R[2] * z) + P.tb[0];
      q1 = (P.R[3] * x + P.R[4] * y + P.R[5] * z) + P.tb[1];
      q2 = (P.R[6] * x + P.R[7] * y + P.R[8] * z) + P.tb[2];
    }
    const int64_t o = (int64_t)m * P.n_points + p0 + threadIdx.x;
    P.pts[3 * o] = q0; P.pts[3 * o + 1] = q1; P.pts[3 * o + 2] = q2;
    P.t[o] = tt; P.w[o] = wv; P.ring[o] = rg; P.tag[o] = 0;
  }
  // integer flags: order-independent, so atomics keep the result deterministic
  const unsigned any_big = __ballot_sync(0xffffffffu, big_t);
  n_bad = __reduce_add_sync(0xffffffffu, n_bad);
  if ((threadIdx.x & 31) == 0) {
    if (any_big) atomicOr(&P.flags[2 * m], 1);
    if (n_bad) atomicAdd(&P.flags[2 * m + 1], n_bad);
  }
}

// one CTA per message: nanoseconds -> seconds when the flag is set, and the certificate record
__global__ void __launch_bounds__(1024) pc2_finish_kernel(const int* __restrict__ flags, int64_t n_points,
                                                          double* __restrict__ t, double* __restrict__ cert) {
  const int m = blockIdx.x;
  const int ns = flags[2 * m];
  if (threadIdx.x == 0) {
    cert[m * GCS_PC_NCERT + GCS_PC_N_NONFINITE] = (double)flags[2 * m + 1];
    cert[m * GCS_PC_NCERT + GCS_PC_TIME_RESCALED] = (double)ns;
    cert[m * GCS_PC_NCERT + 2] = 0.0; cert[m * GCS_PC_NCERT + 3] = 0.0;
  }
  if (!ns) return;
  double* tm = t + (int64_t)m * n_points;
  for (int64_t i = threadIdx.x; i < n_points; i += blockDim.x) tm[i] *= 1e-9;
}

bool type_ok(int t) { return t >= 1 && t <= 8; }
int type_size(int t) { return (t <= 2) ? 1 : (t <= 4) ? 2 : (t <= 7) ? 4 : 8; }

}  // namespace
}  // namespace gcs

extern "C" int gcs_parse_pointcloud2_vlp16(gcs_ctx* ctx, void* stream, const uint8_t* data, int n_msgs, int64_t n_points,
                                           const gcs_pc2_layout* lay, const double* header_stamp,
                                           const double* R_base_lidar, const double* t_base_lidar, double* pts, double* t,
                                           double* w, uint8_t* ring, uint8_t* tag, double* cert) {
  if (!ctx) return GCS_EINVAL;
  GCS_CHECK_CUDA(ctx, cudaSetDevice(ctx->device));
  GCS_REQUIRE(ctx, lay != nullptr, "gcs_parse_pointcloud2_vlp16: layout is NULL");
  // the decoder reads the payload in aligned 16-byte words
  GCS_REQUIRE(ctx, n_points == 0 || (data != nullptr && ((uintptr_t)data & 15) == 0),
              "gcs_parse_pointcloud2_vlp16: payload pointer must be 16-byte aligned (clone a view with an odd storage offset)");
  GCS_REQUIRE(ctx, n_msgs >= 1 && n_points >= 0, "gcs_parse_pointcloud2_vlp16: n_msgs=%d n_points=%lld", n_msgs, (long long)n_points);
  // backend_node.py:398-403: x, y, z, ring are required
  GCS_REQUIRE(ctx, lay->off_x >= 0 && lay->off_y >= 0 && lay->off_z >= 0 && lay->off_ring >= 0,
              "PointCloud2 (VLP-16 layout) missing required fields (x, y, z, ring)");
  GCS_REQUIRE(ctx, lay->point_step >= 1 && lay->point_step <= gcs::kMaxPointStep, "gcs_parse_pointcloud2_vlp16: point_step=%d not in [1,%d]",
              lay->point_step, gcs::kMaxPointStep);
  const int offs[5] = {lay->off_x, lay->off_y, lay->off_z, lay->off_ring, lay->off_time};
  const int types[5] = {lay->type_x, lay->type_y, lay->type_z, lay->type_ring, lay->type_time};
  for (int k = 0; k < 5; ++k) {
    if (k == 4 && offs[k] < 0) continue;
    GCS_REQUIRE(ctx, gcs::type_ok(types[k]), "gcs_parse_pointcloud2_vlp16: unsupported PointField datatype %d", types[k]);
    GCS_REQUIRE(ctx, offs[k] + gcs::type_size(types[k]) <= lay->point_step, "gcs_parse_pointcloud2_vlp16: field at offset %d exceeds point_step %d",
                offs[k], lay->point_step);
  }
  GCS_REQUIRE(ctx, lay->off_time >= 0 || header_stamp != nullptr, "gcs_parse_pointcloud2_vlp16: no time field and no header stamp");
  GCS_REQUIRE(ctx, (R_base_lidar == nullptr) == (t_base_lidar == nullptr), "gcs_parse_pointcloud2_vlp16: give both R_base_lidar and t_base_lidar or neither");
  GCS_REQUIRE(ctx, cert && (n_points == 0 || (data && pts && t && w && ring && tag)), "gcs_parse_pointcloud2_vlp16: a required device pointer is NULL");
  cudaStream_t st = (cudaStream_t)stream;
  int rc = gcs_ws_reserve(ctx, (uint64_t)n_msgs * 2 * sizeof(int) + 64);
  if (rc != GCS_OK) return rc;
  int* flags = (int*)ctx->ws;
  GCS_CHECK_CUDA(ctx, cudaMemsetAsync(flags, 0, (size_t)n_msgs * 2 * sizeof(int), st));
  if (n_points > 0) {
    gcs::Pc2Params P;
    P.data = data; P.n_points = n_points; P.lay = *lay; P.header_stamp = header_stamp;
    P.has_tf = R_base_lidar != nullptr;
    for (int k = 0; k < 9; ++k) P.R[k] = P.has_tf ? R_base_lidar[k] : (k % 4 == 0 ? 1.0 : 0.0);
    for (int k = 0; k < 3; ++k) P.tb[k] = P.has_tf ? t_base_lidar[k] : 0.0;
    P.sentinel = 1e6; P.sigma = 0.25; P.min_r = 0.5; P.max_r = 50.0; P.wfloor = 1e-12;   // common/constants.py:256-262
    P.pts = pts; P.t = t; P.w = w; P.ring = ring; P.tag = tag; P.flags = flags;
    const int smem = gcs::kParseThreads * lay->point_step + 32;
    GCS_CHECK_CUDA(ctx, gcs_smem_attr_once((const void*)gcs::pc2_parse_kernel, gcs::kParseThreads * gcs::kMaxPointStep + 32));
    dim3 grid((unsigned)((n_points + gcs::kParseThreads - 1) / gcs::kParseThreads), (unsigned)n_msgs);
    gcs::pc2_parse_kernel<<<grid, gcs::kParseThreads, smem, st>>>(P);
    GCS_LAUNCH_CHECK(ctx);
  }
  gcs::pc2_finish_kernel<<<n_msgs, 1024, 0, st>>>(flags, n_points, t, cert);
  GCS_LAUNCH_CHECK(ctx);
  return GCS_OK;
}
