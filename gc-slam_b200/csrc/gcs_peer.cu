// gcs_peer.cu -- peer-memory exchange of the per-bin statistics of a point-sharded cloud.
//
// The payload of the exchange is tiny (32 B of resample masses, 9.7 KB of raw sums per unit): a library collective costs
// its launch latency (20-45 us measured for all-gather + reduction), not bandwidth.  Here every rank owns a receive window
// in its own HBM (cudaMalloc, exported with cudaIpcGetMemHandle and mapped by every peer: NVLink / NVSwitch peer access)
// and ONE kernel per rank does the whole exchange:
//     push    my packed block [additive | maxima] into slot `my rank` of EVERY rank's window (remote 8-byte stores)
//     signal  __threadfence_system, then a release store of the epoch into my flag in every rank's window
//     wait    acquire-poll my own window's flags until every rank's epoch has arrived
//     reduce  add / maximise the `world` blocks of my window in rank order into the caller's buffer
// Same data, same order on every rank: bit-identical statistics everywhere.  Two windows alternate with the epoch's
// parity: a rank can only be one exchange ahead of a peer (it needs the peer's flag of exchange k+1 to get past it, and
// the peer sets that after it has finished reading exchange k), so a window is never overwritten while it is being read.
// The poll is bounded: after ~2 s without a peer's flag the kernel records a time-out in the window's status word and
// returns (the host reports it as GCS_ECOMM on the next call) instead of hanging the device.
#include "gcs_common.cuh"

struct gcs_peer_xchg {
  int rank, world;
  uint64_t bytes_per_rank;   // capacity of one slot
  uint64_t window_bytes;     // one window: world slots + world flags + status
  unsigned char* local;      // my allocation: 2 windows
  unsigned char* peer[64];   // every rank's allocation as mapped here (peer[rank] == local)
  unsigned epoch;
  int connected;
};

namespace gcs {

struct PeerPtrs { unsigned char* p[64]; };

__device__ __forceinline__ void st_release_sys(unsigned* p, unsigned v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// window layout: [world slots of slot_bytes][world flags (u32), padded to 256 B][status u32]
__global__ void __launch_bounds__(512) peer_exchange_kernel(PeerPtrs P, int rank, int world, uint64_t window_off, uint64_t slot_bytes,
                                                           unsigned epoch, double* __restrict__ pack, int64_t n_sum, int64_t n_max) {
  const int tid = threadIdx.x;
  const int64_t n = n_sum + n_max;
  const uint64_t flags_off = window_off + (uint64_t)world * slot_bytes;
  // 1. push
  for (int r = 0; r < world; ++r) {
    double* dst = reinterpret_cast<double*>(P.p[r] + window_off + (uint64_t)rank * slot_bytes);
    for (int64_t i = tid; i < n; i += blockDim.x) dst[i] = pack[i];
  }
  __threadfence_system();
  __syncthreads();
  // 2. signal  3. wait
  __shared__ int s_timeout;
  if (tid == 0) s_timeout = 0;
  __syncthreads();
  if (tid < world) {
    st_release_sys(reinterpret_cast<unsigned*>(P.p[tid] + flags_off) + rank, epoch);
    const unsigned* mine = reinterpret_cast<const unsigned*>(P.p[rank] + flags_off) + tid;
    long long t0 = clock64();
    while (ld_acquire_sys(mine) != epoch) {
      if (clock64() - t0 > 4000000000ll) { s_timeout = 1; break; }   // ~2 s at 1.9 GHz
      __nanosleep(100);
    }
  }
  __syncthreads();
  if (s_timeout) {
    if (tid == 0) *reinterpret_cast<unsigned*>(P.p[rank] + flags_off + 256) = epoch;   // status: the exchange that timed out
    return;
  }
  // 4. reduce my window in rank order
  const unsigned char* win = P.p[rank] + window_off;
  for (int64_t i = tid; i < n; i += blockDim.x) {
    double a = reinterpret_cast<const double*>(win)[i];
    if (i < n_sum)
      for (int r = 1; r < world; ++r) a += reinterpret_cast<const double*>(win + (uint64_t)r * slot_bytes)[i];
    else
      for (int r = 1; r < world; ++r) a = fmax(a, reinterpret_cast<const double*>(win + (uint64_t)r * slot_bytes)[i]);
    pack[i] = a;
  }
}

}  // namespace gcs

extern "C" {

int gcs_peer_xchg_create(gcs_ctx* ctx, int32_t rank, int32_t world, uint64_t bytes_per_rank, gcs_peer_xchg** out,
                         void* ipc_handle_out) {
  if (!ctx) return GCS_EINVAL;
  GCS_CHECK_CUDA(ctx, cudaSetDevice(ctx->device));
  GCS_REQUIRE(ctx, out && ipc_handle_out && world >= 1 && world <= 64 && rank >= 0 && rank < world && bytes_per_rank >= 8,
              "peer_xchg_create: bad args (world %d, rank %d)", world, rank);
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  gcs_peer_xchg* x = (gcs_peer_xchg*)calloc(1, sizeof(gcs_peer_xchg));
  if (!x) return gcs_set_error(ctx, GCS_ENOMEM, "peer_xchg_create: calloc");
  x->rank = rank; x->world = world;
  x->bytes_per_rank = (bytes_per_rank + 255) & ~(uint64_t)255;
  x->window_bytes = (uint64_t)world * x->bytes_per_rank + 512;
  cudaError_t e = cudaMalloc((void**)&x->local, 2 * x->window_bytes);
  if (e != cudaSuccess) { free(x); return gcs_set_error(ctx, GCS_ENOMEM, "peer_xchg_create: cudaMalloc: %s", cudaGetErrorString(e)); }
  cudaMemset(x->local, 0, 2 * x->window_bytes);
  e = cudaIpcGetMemHandle((cudaIpcMemHandle_t*)ipc_handle_out, x->local);
  if (e != cudaSuccess) {
    cudaFree(x->local); free(x);
    return gcs_set_error(ctx, GCS_ECOMM, "peer_xchg_create: cudaIpcGetMemHandle: %s", cudaGetErrorString(e));
  }
  x->peer[rank] = x->local;
  *out = x;
  return GCS_OK;
}

int gcs_peer_xchg_connect(gcs_ctx* ctx, gcs_peer_xchg* x, const void* all_handles) {
  if (!ctx || !x || !all_handles) return GCS_EINVAL;
  GCS_CHECK_CUDA(ctx, cudaSetDevice(ctx->device));
  const cudaIpcMemHandle_t* h = (const cudaIpcMemHandle_t*)all_handles;
  for (int r = 0; r < x->world; ++r) {
    if (r == x->rank) continue;
    void* p = nullptr;
    cudaError_t e = cudaIpcOpenMemHandle(&p, h[r], cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess)
      return gcs_set_error(ctx, GCS_ECOMM, "peer_xchg_connect: cudaIpcOpenMemHandle(rank %d): %s", r, cudaGetErrorString(e));
    x->peer[r] = (unsigned char*)p;
  }
  x->connected = 1;
  return GCS_OK;
}

int gcs_peer_xchg_reduce(gcs_ctx* ctx, gcs_peer_xchg* x, void* stream, double* pack, int64_t n_sum, int64_t n_max) {
  if (!ctx || !x) return GCS_EINVAL;
  GCS_CHECK_CUDA(ctx, cudaSetDevice(ctx->device));
  GCS_REQUIRE(ctx, x->connected && pack && n_sum >= 0 && n_max >= 0 && (uint64_t)(n_sum + n_max) * 8 <= x->bytes_per_rank &&
                       n_sum + n_max >= 1, "peer_xchg_reduce: not connected or %lld values exceed the slot", (long long)(n_sum + n_max));
  gcs::PeerPtrs P;
  for (int r = 0; r < 64; ++r) P.p[r] = r < x->world ? x->peer[r] : nullptr;
  ++x->epoch;
  const uint64_t window_off = (uint64_t)(x->epoch & 1u) * x->window_bytes;
  gcs::peer_exchange_kernel<<<1, 512, 0, (cudaStream_t)stream>>>(P, x->rank, x->world, window_off, x->bytes_per_rank, x->epoch, pack,
                                                                n_sum, n_max);
  GCS_LAUNCH_CHECK(ctx);
  return GCS_OK;
}

/* 0: fine; otherwise the epoch of the exchange whose poll timed out (a peer never signalled). Synchronises the stream's device. */
int gcs_peer_xchg_status(gcs_ctx* ctx, gcs_peer_xchg* x, uint32_t* out_epoch) {
  if (!ctx || !x || !out_epoch) return GCS_EINVAL;
  GCS_CHECK_CUDA(ctx, cudaSetDevice(ctx->device));
  unsigned s[2] = {0, 0};
  for (int wdw = 0; wdw < 2; ++wdw)
    GCS_CHECK_CUDA(ctx, cudaMemcpy(&s[wdw], x->local + wdw * x->window_bytes + (uint64_t)x->world * x->bytes_per_rank + 256, 4,
                                   cudaMemcpyDeviceToHost));
  *out_epoch = s[0] > s[1] ? s[0] : s[1];
  return GCS_OK;
}

int gcs_peer_xchg_destroy(gcs_ctx* ctx, gcs_peer_xchg* x) {
  if (!x) return GCS_OK;
  if (ctx) cudaSetDevice(ctx->device);
  cudaDeviceSynchronize();
  for (int r = 0; r < x->world; ++r)
    if (r != x->rank && x->peer[r]) cudaIpcCloseMemHandle(x->peer[r]);
  if (x->local) cudaFree(x->local);
  free(x);
  return GCS_OK;
}

}  // extern "C"
