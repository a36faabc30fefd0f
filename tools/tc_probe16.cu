// tc_probe16.cu -- checks the 16-bit (kind::f16, K-major SWIZZLE_64B) tcgen05 plumbing of gc-slam_b200/csrc/gcs_tc.cuh on a
// B200: layout / descriptor correctness at M = 128, N = 40, K = 32 (two MMAs), accumulator rounding over many steps,
// cycles per MMA.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I gc-slam_b200/csrc -o tools/tc_probe16 tools/tc_probe16.cu
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

#include "gcs_tc.cuh"

using namespace gcs::tc;

constexpr int M = 128, N = 40, NB = 64;

struct Smem {
  alignas(1024) unsigned char A[M * 64];
  alignas(1024) unsigned char B[NB * 64];
  alignas(8) uint64_t bar;
  uint32_t tmem;
};

// mode 0: D = A.B^T once; mode 1: accumulate the tile `reps` times; mode 2: timing of reps MMAs
__global__ void __launch_bounds__(128) probe16(const __half* __restrict__ A, const __half* __restrict__ B, float* __restrict__ D,
                                               int reps, int mode, long long* __restrict__ clk) {
  extern __shared__ __align__(1024) unsigned char raw[];
  Smem& sm = *reinterpret_cast<Smem*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  const int tid = threadIdx.x, wid = tid >> 5, lane = tid & 31;
  for (int idx = tid; idx < M * 32; idx += 128) {
    int r = idx >> 5, k = idx & 31;
    *reinterpret_cast<__half*>(sm.A + swz64_offset(r, 2 * k)) = A[idx];
  }
  for (int idx = tid; idx < NB * 32; idx += 128) {
    int r = idx >> 5, k = idx & 31;
    *reinterpret_cast<__half*>(sm.B + swz64_offset(r, 2 * k)) = r < N ? B[idx] : __float2half(0.f);
  }
  if (tid == 0) { mbar_init(&sm.bar, 1); mbar_init_fence(); }
  fence_smem_to_async();
  if (wid == 0) tmem_alloc(&sm.tmem, 512);
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem = sm.tmem;
  const uint32_t idesc = idesc_f16(M, N);
  long long t0 = 0, t1 = 0;
  if (tid == 0) {
    const uint64_t da = smem_desc_sw64(smem_u32(sm.A)), db = smem_desc_sw64(smem_u32(sm.B));
    t0 = clock64();
    if (mode == 2) {
      for (int r = 0; r < reps; ++r) mma_f16_ss(tmem + 64 * (r & 3), da + 2 * (r & 1), db + 2 * (r & 1), idesc, r > 3);
    } else {
      for (int r = 0; r < reps; ++r)
        for (int ks = 0; ks < 2; ++ks) mma_f16_ss(tmem, da + 2 * ks, db + 2 * ks, idesc, (r | ks) > 0);
    }
    mma_commit(&sm.bar);
  }
  mbar_wait(&sm.bar, 0);
  if (tid == 0) { t1 = clock64(); clk[0] = t1 - t0; }
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

int main() {
  std::vector<__half> A(M * 32), B(NB * 32);
  std::vector<float> Af(M * 32), Bf(NB * 32), D(M * N);
  __half *dA, *dB; float* dD; long long* dclk;
  cudaMalloc(&dA, A.size() * 2); cudaMalloc(&dB, B.size() * 2); cudaMalloc(&dD, D.size() * 4); cudaMalloc(&dclk, 8);
  const size_t smem = sizeof(Smem) + 1024;
  cudaFuncSetAttribute(probe16, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  srand(1);
  auto rnd = []() { return (float)rand() / RAND_MAX; };
  auto fill = [&](float lo, float hi) {
    for (size_t i = 0; i < A.size(); ++i) { A[i] = __float2half(lo + (hi - lo) * rnd()); Af[i] = __half2float(A[i]); }
    for (size_t i = 0; i < B.size(); ++i) { B[i] = __float2half(lo + (hi - lo) * rnd()); Bf[i] = __half2float(B[i]); }
    cudaMemcpy(dA, A.data(), A.size() * 2, cudaMemcpyHostToDevice);
    cudaMemcpy(dB, B.data(), B.size() * 2, cudaMemcpyHostToDevice);
  };
  fill(-1.f, 1.f);
  probe16<<<1, 128, smem>>>(dA, dB, dD, 1, 0, dclk);
  cudaError_t e = cudaDeviceSynchronize();
  printf("probe16-1 launch: %s\n", cudaGetErrorString(e));
  if (e != cudaSuccess) return 1;
  cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
  double worst = 0;
  for (int m = 0; m < M; ++m)
    for (int n = 0; n < N; ++n) {
      double s = 0;
      for (int k = 0; k < 32; ++k) s += (double)Af[m * 32 + k] * (double)Bf[n * 32 + k];
      worst = fmax(worst, fabs(s - D[m * N + n]));
    }
  printf("probe16-1 layout (SW64, f16, M=128 N=40 K=32): max |D - A.B^T| = %.3e  %s\n", worst, worst < 1e-4 ? "OK" : "MISMATCH");

  fill(0.5f, 1.5f);
  for (int reps : {1, 2, 4, 8, 16, 64}) {
    probe16<<<1, 128, smem>>>(dA, dB, dD, reps, 1, dclk);
    cudaDeviceSynchronize();
    cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
    double mean = 0, rms = 0;
    for (int m = 0; m < M; ++m)
      for (int n = 0; n < N; ++n) {
        double s = 0;
        for (int k = 0; k < 32; ++k) s += (double)Af[m * 32 + k] * (double)Bf[n * 32 + k];
        s *= reps;
        double rel = (D[m * N + n] - s) / s;
        mean += rel; rms += rel * rel;
      }
    mean /= M * N; rms = sqrt(rms / (M * N));
    printf("probe16-2 accumulate %3d tiles (%4d mma steps): mean rel err %+.3e  rms %.3e\n", reps, reps * 2, mean, rms);
  }
  for (int reps : {256, 2048}) {
    long long clk = 0;
    probe16<<<1, 128, smem>>>(dA, dB, dD, reps, 2, dclk);
    cudaDeviceSynchronize();
    cudaMemcpy(&clk, dclk, 8, cudaMemcpyDeviceToHost);
    printf("probe16-3 M=128 N=40 K=16: %5d MMAs in %8lld clk -> %.1f clk/MMA\n", reps, clk, (double)clk / reps);
  }
  return 0;
}
