// gcs_context.cu -- library context, error reporting, workspace.
#include <stdarg.h>
#include <stdlib.h>

#include "gcs_common.cuh"

static char g_create_err[512] = "";

int gcs_set_error(gcs_ctx* ctx, int code, const char* fmt, ...) {
  char* dst = ctx ? ctx->err : g_create_err;
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(dst, 511, fmt, ap);
  va_end(ap);
  return code;
}

int gcs_ws_reserve(gcs_ctx* ctx, uint64_t bytes) {
  if (bytes <= ctx->ws_bytes) return GCS_OK;
  // grow-only; rounded up so that steady-state calls never allocate
  uint64_t want = ((bytes + (1ull << 20) - 1) >> 20) << 20;
  void* p = nullptr;
  cudaError_t e = cudaMalloc(&p, want);
  if (e != cudaSuccess)
    return gcs_set_error(ctx, GCS_ENOMEM, "workspace cudaMalloc(%llu) failed: %s", (unsigned long long)want,
                         cudaGetErrorString(e));
  if (ctx->ws) cudaFree(ctx->ws);
  ctx->ws = p;
  ctx->ws_bytes = want;
  return GCS_OK;
}

extern "C" {

int gcs_version(void) { return GCS_VERSION_MAJOR * 10000 + GCS_VERSION_MINOR * 100 + GCS_VERSION_PATCH; }

const char* gcs_version_string(void) { return "gcs_sm100a 0.1.0"; }

int gcs_create(gcs_ctx** out, int device) {
  if (!out) return gcs_set_error(nullptr, GCS_EINVAL, "gcs_create: out is NULL");
  *out = nullptr;
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n <= 0)
    return gcs_set_error(nullptr, GCS_ECUDA, "gcs_create: no CUDA device (%s); this library has no CPU fallback",
                         cudaGetErrorString(e));
  if (device < 0 || device >= n) return gcs_set_error(nullptr, GCS_EINVAL, "gcs_create: device %d of %d", device, n);
  e = cudaSetDevice(device);
  if (e != cudaSuccess) return gcs_set_error(nullptr, GCS_ECUDA, "cudaSetDevice: %s", cudaGetErrorString(e));
  cudaDeviceProp prop;
  e = cudaGetDeviceProperties(&prop, device);
  if (e != cudaSuccess) return gcs_set_error(nullptr, GCS_ECUDA, "cudaGetDeviceProperties: %s", cudaGetErrorString(e));
  if (prop.major != 10)
    return gcs_set_error(nullptr, GCS_ECUDA, "gcs_create: device is sm_%d%d; this build is sm_100a only", prop.major,
                         prop.minor);
  gcs_ctx* c = (gcs_ctx*)calloc(1, sizeof(gcs_ctx));
  if (!c) return gcs_set_error(nullptr, GCS_ENOMEM, "gcs_create: calloc");
  c->device = device;
  c->sm_count = prop.multiProcessorCount;
  *out = c;
  return GCS_OK;
}

int gcs_timing_enable(gcs_ctx* ctx, int on) {
  if (!ctx) return GCS_EINVAL;
  GCS_CHECK_CUDA(ctx, cudaSetDevice(ctx->device));
  if (on && !ctx->timing_ev[0])
    for (int i = 0; i < 2 * 256; ++i) GCS_CHECK_CUDA(ctx, cudaEventCreate(&ctx->timing_ev[i]));
  ctx->timing_on = on < 0 ? 0 : on;
  ctx->timing_n = 0;
  return GCS_OK;
}

int gcs_timing_collect(gcs_ctx* ctx, double* total_ms, int* count) {
  if (!ctx || !total_ms || !count) return GCS_EINVAL;
  GCS_CHECK_CUDA(ctx, cudaSetDevice(ctx->device));
  double tot = 0.0;
  for (int i = 0; i < ctx->timing_n; ++i) {
    GCS_CHECK_CUDA(ctx, cudaEventSynchronize(ctx->timing_ev[2 * i + 1]));
    float ms = 0.f;
    GCS_CHECK_CUDA(ctx, cudaEventElapsedTime(&ms, ctx->timing_ev[2 * i], ctx->timing_ev[2 * i + 1]));
    tot += ms;
  }
  *total_ms = tot;
  *count = ctx->timing_n;
  ctx->timing_n = 0;
  return GCS_OK;
}

int gcs_destroy(gcs_ctx* ctx) {
  if (!ctx) return GCS_OK;
  cudaSetDevice(ctx->device);
  if (ctx->timing_ev[0])
    for (int i = 0; i < 2 * 256; ++i) cudaEventDestroy(ctx->timing_ev[i]);
  if (ctx->ws) cudaFree(ctx->ws);
  free(ctx);
  return GCS_OK;
}

const char* gcs_last_error(gcs_ctx* ctx) { return ctx ? ctx->err : g_create_err; }

int gcs_reserve_workspace(gcs_ctx* ctx, uint64_t bytes) {
  if (!ctx) return GCS_EINVAL;
  GCS_CHECK_CUDA(ctx, cudaSetDevice(ctx->device));
  return gcs_ws_reserve(ctx, bytes);
}

int gcs_device_sm_count(gcs_ctx* ctx) { return ctx ? ctx->sm_count : 0; }

uint64_t gcs_kernel_launches(gcs_ctx* ctx) { return ctx ? ctx->launches : 0; }

}  // extern "C"
