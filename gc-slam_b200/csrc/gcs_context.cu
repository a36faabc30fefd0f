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

#include <mutex>

cudaError_t gcs_smem_attr_once(const void* kern, int bytes) {
  struct Entry { const void* k; int bytes[64]; };   // bytes[d]: largest size set for this kernel on device d (0: never)
  static Entry tab[128];
  static int n = 0;
  static std::mutex mu;
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  std::lock_guard<std::mutex> lock(mu);
  Entry* ent = nullptr;
  for (int i = 0; i < n; ++i)
    if (tab[i].k == kern) { ent = &tab[i]; break; }
  if (ent && dev < 64 && ent->bytes[dev] >= bytes) return cudaSuccess;
  e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  if (e != cudaSuccess) return e;
  if (!ent && n < 128) { ent = &tab[n++]; ent->k = kern; memset(ent->bytes, 0, sizeof(ent->bytes)); }
  if (ent && dev < 64) ent->bytes[dev] = bytes;
  return cudaSuccess;
}

// Grow-only workspace.  Growth doubles (at least) so that it happens a handful of times per process; an outgrown block
// is RETIRED, not freed: cudaFree synchronises the device and kernels enqueued by earlier calls may still be reading it.
// Retired blocks are released by gcs_reserve_workspace (a call made outside the steady state) and gcs_destroy.
int gcs_ws_reserve(gcs_ctx* ctx, uint64_t bytes) {
  if (bytes <= ctx->ws_bytes) return GCS_OK;
  if (ctx->ws_frozen)
    return gcs_set_error(ctx, GCS_ENOMEM, "workspace frozen at %llu bytes, this call needs %llu: size it with "
                         "gcs_reserve_workspace before gcs_workspace_freeze", (unsigned long long)ctx->ws_bytes,
                         (unsigned long long)bytes);
  uint64_t want = ((bytes + (1ull << 20) - 1) >> 20) << 20;
  if (want < 2 * ctx->ws_bytes) want = 2 * ctx->ws_bytes;
  void* p = nullptr;
  cudaError_t e = cudaMalloc(&p, want);
  if (e != cudaSuccess)
    return gcs_set_error(ctx, GCS_ENOMEM, "workspace cudaMalloc(%llu) failed: %s", (unsigned long long)want,
                         cudaGetErrorString(e));
  if (ctx->ws) {
    if (ctx->n_retired < 32) ctx->ws_retired[ctx->n_retired++] = ctx->ws;
    else { cudaDeviceSynchronize(); cudaFree(ctx->ws); }
  }
  ctx->ws = p;
  ctx->ws_bytes = want;
  return GCS_OK;
}

int gcs_side_reserve(gcs_ctx* ctx, uint64_t bytes) {
  if (!ctx->side_stream) {
    cudaError_t e = cudaStreamCreateWithFlags(&ctx->side_stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ctx->ev_fork, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ctx->ev_join, cudaEventDisableTiming);
    if (e != cudaSuccess) return gcs_set_error(ctx, GCS_ECUDA, "side stream: %s", cudaGetErrorString(e));
  }
  if (bytes <= ctx->ws_side_bytes) return GCS_OK;
  if (ctx->ws_frozen)
    return gcs_set_error(ctx, GCS_ENOMEM, "workspace frozen: the side workspace needs %llu bytes", (unsigned long long)bytes);
  // growth (once per view shape): nothing enqueued earlier may still use the old block
  cudaDeviceSynchronize();
  if (ctx->ws_side) cudaFree(ctx->ws_side);
  ctx->ws_side = nullptr; ctx->ws_side_bytes = 0;
  const uint64_t want = ((bytes + (1ull << 20) - 1) >> 20) << 20;
  cudaError_t e = cudaMalloc(&ctx->ws_side, want);
  if (e != cudaSuccess)
    return gcs_set_error(ctx, GCS_ENOMEM, "side workspace cudaMalloc(%llu) failed: %s", (unsigned long long)want, cudaGetErrorString(e));
  ctx->ws_side_bytes = want;
  return GCS_OK;
}

int gcs_maint_stream(gcs_ctx* ctx, cudaStream_t caller, uint64_t ws_bytes, cudaStream_t* st, char** ws) {
  if (!ctx->route_side) {
    const int rc = gcs_ws_reserve(ctx, ws_bytes);
    if (rc) return rc;
    *st = caller; *ws = (char*)ctx->ws;
    return GCS_OK;
  }
  const int rc = gcs_side_reserve(ctx, ws_bytes);
  if (rc) return rc;
  GCS_CHECK_CUDA(ctx, cudaEventRecord(ctx->ev_fork, caller));
  GCS_CHECK_CUDA(ctx, cudaStreamWaitEvent(ctx->side_stream, ctx->ev_fork, 0));
  *st = ctx->side_stream; *ws = (char*)ctx->ws_side;
  return GCS_OK;
}

static void gcs_ws_release_retired(gcs_ctx* ctx) {
  if (ctx->n_retired == 0) return;
  cudaDeviceSynchronize();
  for (int i = 0; i < ctx->n_retired; ++i) cudaFree(ctx->ws_retired[i]);
  ctx->n_retired = 0;
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
  // default workspace: enough for a single scan of the reference budgets through either family, so that the first calls
  // do not allocate; batches size themselves once (or up front through gcs_reserve_workspace)
  if (gcs_ws_reserve(c, 32ull << 20) != GCS_OK) { free(c); return gcs_set_error(nullptr, GCS_ENOMEM, "gcs_create: workspace"); }
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
  gcs_ws_release_retired(ctx);
  if (ctx->ws) cudaFree(ctx->ws);
  if (ctx->side_stream) {
    cudaStreamSynchronize(ctx->side_stream);
    cudaStreamDestroy(ctx->side_stream);
    cudaEventDestroy(ctx->ev_fork);
    cudaEventDestroy(ctx->ev_join);
  }
  if (ctx->ws_side) cudaFree(ctx->ws_side);
  free(ctx);
  return GCS_OK;
}

const char* gcs_last_error(gcs_ctx* ctx) { return ctx ? ctx->err : g_create_err; }

int gcs_reserve_workspace(gcs_ctx* ctx, uint64_t bytes) {
  if (!ctx) return GCS_EINVAL;
  GCS_CHECK_CUDA(ctx, cudaSetDevice(ctx->device));
  const int frozen = ctx->ws_frozen;
  ctx->ws_frozen = 0;
  const int rc = gcs_ws_reserve(ctx, bytes);
  ctx->ws_frozen = frozen;
  gcs_ws_release_retired(ctx);   // an explicit sizing call is made outside the steady state: safe to synchronise
  return rc;
}

int gcs_side_route(gcs_ctx* ctx, int on) {
  if (!ctx) return GCS_EINVAL;
  ctx->route_side = on ? 1 : 0;
  return GCS_OK;
}

int gcs_side_join(gcs_ctx* ctx, void* stream) {
  if (!ctx) return GCS_EINVAL;
  if (!ctx->side_stream) return GCS_OK;
  GCS_CHECK_CUDA(ctx, cudaSetDevice(ctx->device));
  GCS_CHECK_CUDA(ctx, cudaEventRecord(ctx->ev_join, ctx->side_stream));
  GCS_CHECK_CUDA(ctx, cudaStreamWaitEvent((cudaStream_t)stream, ctx->ev_join, 0));
  return GCS_OK;
}

int gcs_workspace_freeze(gcs_ctx* ctx, int frozen) {
  if (!ctx) return GCS_EINVAL;
  ctx->ws_frozen = frozen ? 1 : 0;
  return GCS_OK;
}

uint64_t gcs_workspace_bytes(gcs_ctx* ctx) { return ctx ? ctx->ws_bytes : 0; }

int gcs_device_sm_count(gcs_ctx* ctx) { return ctx ? ctx->sm_count : 0; }

uint64_t gcs_kernel_launches(gcs_ctx* ctx) { return ctx ? ctx->launches : 0; }

}  // extern "C"
