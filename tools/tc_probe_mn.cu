// tc_probe_mn.cu -- checks the MN-major, no-swizzle ("interleaved") shared-memory operand layout used by the second
// generation of the tensor-core bin kernel (gcs_bins_tc.cu): kind::f16, A = [M x K] with M contiguous, B = [N x K] with N
// contiguous, core matrix = 8 K-rows x 16 bytes (8 fp16 along M/N) = 128 contiguous bytes.
//   address(mn, k) = (k >> 3) * KG + (mn >> 3) * 128 + (k & 7) * 16 + (mn & 7) * 2
// One lane = one point k writes its 8-element chunks with 16-byte stores; lanes 0..7 of a quarter warp cover one core
// matrix = 128 contiguous bytes (conflict-free).
// usage: tc_probe_mn <variant>   variant 0: LBO = K-group stride, SBO = MN-block stride;  1: swapped
//                                 +2: A tile with 12 M-blocks only (KG_A = 1536), rows >= 96 alias the next K group
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I gc-slam_b200/csrc -o tools/tc_probe_mn tools/tc_probe_mn.cu
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

#include "gcs_tc.cuh"

using namespace gcs::tc;

constexpr int M = 128, N = 40, K = 32;

__device__ __forceinline__ uint64_t desc_mn_none(uint32_t addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
  d |= (uint64_t)1 << 46;   // descriptor version (sm_100); layout type 0 = SWIZZLE_NONE
  return d;
}

__global__ void __launch_bounds__(128) probe(const __half* __restrict__ A, const __half* __restrict__ B, float* __restrict__ D,
                                             int variant, int reps, long long* __restrict__ clk) {
  extern __shared__ __align__(1024) unsigned char raw[];
  unsigned char* sm = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  const int kgA = (variant & 2) ? 12 * 128 : 16 * 128, kgB = 5 * 128;
  unsigned char* sA = sm;
  unsigned char* sB = sm + 4 * 2048;
  uint64_t* bar = reinterpret_cast<uint64_t*>(sm + 4 * 2048 + 4 * 1024);
  uint32_t* tm = reinterpret_cast<uint32_t*>(bar + 1);
  const int tid = threadIdx.x, wid = tid >> 5, lane = tid & 31;
  for (int i = tid; i < (4 * 2048 + 4 * 1024) / 4; i += 128) reinterpret_cast<uint32_t*>(sm)[i] = 0;
  __syncthreads();
  const int m_rows = (variant & 2) ? 96 : M;
  for (int idx = tid; idx < m_rows * K; idx += 128) {
    int m = idx / K, k = idx % K;
    *reinterpret_cast<__half*>(sA + (k >> 3) * kgA + (m >> 3) * 128 + (k & 7) * 16 + (m & 7) * 2) = A[m * K + k];
  }
  for (int idx = tid; idx < N * K; idx += 128) {
    int n = idx / K, k = idx % K;
    *reinterpret_cast<__half*>(sB + (k >> 3) * kgB + (n >> 3) * 128 + (k & 7) * 16 + (n & 7) * 2) = B[n * K + k];
  }
  if (tid == 0) { mbar_init(bar, 1); mbar_init_fence(); }
  fence_smem_to_async();
  if (wid == 0) tmem_alloc(tm, 512);
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem = *tm;
  const uint32_t idesc = idesc_f16(M, N) | (1u << 15) | (1u << 16);   // A and B MN-major
  if (tid == 0) {
    uint64_t da, db;
    if (variant & 1) { da = desc_mn_none(smem_u32(sA), 128, kgA); db = desc_mn_none(smem_u32(sB), 128, kgB); }
    else             { da = desc_mn_none(smem_u32(sA), kgA, 128); db = desc_mn_none(smem_u32(sB), kgB, 128); }
    const long long t0 = clock64();
    for (int r = 0; r < reps; ++r)
      for (int ks = 0; ks < 2; ++ks)   // K = 16 per MMA = two K groups: start address + 2 * KG
        mma_f16_ss(tmem, da + (uint64_t)((2 * ks * kgA) >> 4), db + (uint64_t)((2 * ks * kgB) >> 4), idesc, (r | ks) > 0);
    mma_commit(bar);
    mbar_wait(bar, 0);
    clk[0] = clock64() - t0;
  }
  mbar_wait(bar, 0);
  fence_after_sync();
  uint32_t v[8];
  for (int c = 0; c < N; c += 8) {
    tmem_ld_x8(tmem + ((uint32_t)(32 * wid) << 16) + c, v);
    tmem_ld_wait();
    for (int j = 0; j < 8; ++j) D[(32 * wid + lane) * N + c + j] = __uint_as_float(v[j]);
  }
  fence_before_sync();
  __syncthreads();
  if (wid == 0) tmem_free(tmem, 512);
}

int main(int argc, char** argv) {
  const int variant = argc > 1 ? atoi(argv[1]) : 0;
  std::vector<__half> A(M * K), B(N * K);
  std::vector<float> Af(M * K), Bf(N * K), D(M * N);
  __half *dA, *dB; float* dD; long long* dclk;
  cudaMalloc(&dA, A.size() * 2); cudaMalloc(&dB, B.size() * 2); cudaMalloc(&dD, D.size() * 4); cudaMalloc(&dclk, 8);
  const size_t smem = 4 * 2048 + 4 * 1024 + 64 + 1024;
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  srand(1);
  auto rnd = []() { return (float)rand() / RAND_MAX; };
  for (size_t i = 0; i < A.size(); ++i) { A[i] = __float2half(-1.f + 2.f * rnd()); Af[i] = __half2float(A[i]); }
  for (size_t i = 0; i < B.size(); ++i) { B[i] = __float2half(-1.f + 2.f * rnd()); Bf[i] = __half2float(B[i]); }
  cudaMemcpy(dA, A.data(), A.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, B.data(), B.size() * 2, cudaMemcpyHostToDevice);
  probe<<<1, 128, smem>>>(dA, dB, dD, variant, 1, dclk);
  cudaError_t e = cudaDeviceSynchronize();
  printf("probe_mn variant %d launch: %s\n", variant, cudaGetErrorString(e));
  if (e != cudaSuccess) return 1;
  cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
  double worst = 0;
  const int m_rows = (variant & 2) ? 96 : M;
  for (int m = 0; m < m_rows; ++m)
    for (int n = 0; n < 38; ++n) {
      double s = 0;
      for (int k = 0; k < K; ++k) s += (double)Af[m * K + k] * (double)Bf[n * K + k];
      worst = fmax(worst, fabs(s - D[m * N + n]));
    }
  printf("probe_mn variant %d (MN-major, no swizzle, f16, M=128 N=40 K=32, %d data rows): max |D - A.B^T| = %.3e  %s\n", variant,
         m_rows, worst, worst < 1e-4 ? "OK" : "MISMATCH");
  for (int reps : {128, 1024}) {
    long long clk = 0;
    probe<<<1, 128, smem>>>(dA, dB, dD, variant, reps, dclk);
    cudaDeviceSynchronize();
    cudaMemcpy(&clk, dclk, 8, cudaMemcpyDeviceToHost);
    printf("probe_mn variant %d: %5d MMAs (K=16) in %8lld clk -> %.1f clk/MMA\n", variant, 2 * reps, clk, (double)clk / (2 * reps));
  }
  return 0;
}
