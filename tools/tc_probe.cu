// tc_probe.cu -- checks the tcgen05 plumbing of gc-slam_b200/csrc/gcs_tc.cuh on a B200 before the bin kernel uses it:
//   1. correctness of the K-major SWIZZLE_128B operand layout + descriptors (M=128, N=48, K=32) vs a host product
//   2. how tf32 inputs are rounded (truncate / nearest) and how the f32 accumulator rounds over many steps
//   3. cycles per tcgen05.mma at this shape (operands from shared memory)
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I gc-slam_b200/csrc -o tools/tc_probe tools/tc_probe.cu
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

#include "gcs_tc.cuh"

using namespace gcs::tc;

constexpr int M = 128, N = 48;

struct Smem {
  alignas(1024) unsigned char A[M * 128];
  alignas(1024) unsigned char B[256 * 128];
  alignas(8) uint64_t bar;
  alignas(8) uint64_t bars[4];
  uint32_t tmem;
};

// mode 0: D = A.B^T once (4 k-steps).  mode 1: accumulate the same tile `reps` times.  mode 2: timing of reps MMAs.
__global__ void __launch_bounds__(128) probe_kernel(const float* __restrict__ A, const float* __restrict__ B,
                                                    float* __restrict__ D, int reps, int mode, int n_mma,
                                                    long long* __restrict__ clk) {
  extern __shared__ __align__(1024) unsigned char raw[];
  Smem& sm = *reinterpret_cast<Smem*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  const int tid = threadIdx.x, wid = tid >> 5, lane = tid & 31;
  for (int idx = tid; idx < M * 32; idx += 128) {
    int r = idx >> 5, k = idx & 31;
    *reinterpret_cast<float*>(sm.A + swz_offset(r, k)) = A[idx];
  }
  for (int idx = tid; idx < 256 * 32; idx += 128) {
    int r = idx >> 5, k = idx & 31;
    *reinterpret_cast<float*>(sm.B + swz_offset(r, k)) = r < N ? B[idx] : 0.f;
  }
  if (tid == 0) { mbar_init(&sm.bar, 1); for (int i = 0; i < 4; ++i) mbar_init(&sm.bars[i], 1); mbar_init_fence(); }
  fence_smem_to_async();
  if (wid == 0) tmem_alloc(&sm.tmem, 512);
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem = sm.tmem;
  const uint32_t idesc = idesc_tf32(M, n_mma);
  long long t0 = 0, t1 = 0;
  if (mode >= 100) {
    // mode = 100 + n_warps: lane 0 of the first n_warps warps each issue `reps` MMAs into their own accumulator
    const int nw = mode - 100;
    __syncthreads();
    if (tid == 0) t0 = clock64();
    if (lane == 0 && wid < nw) {
      const uint64_t da = smem_desc_sw128(smem_u32(sm.A)), db = smem_desc_sw128(smem_u32(sm.B));
      for (int r = 0; r < reps; ++r) mma_tf32_ss(tmem + 64 * wid, da + 2 * (r & 3), db + 2 * (r & 3), idesc, r > 0);
      mma_commit(&sm.bars[wid]);
    }
    for (int w = 0; w < nw; ++w) mbar_wait(&sm.bars[w], 0);
    if (tid == 0) { t1 = clock64(); clk[0] = t1 - t0; }
    if (tid == 0) mma_commit(&sm.bar);
  } else
  if (tid == 0) {
    const uint64_t da = smem_desc_sw128(smem_u32(sm.A)), db = smem_desc_sw128(smem_u32(sm.B));
    t0 = clock64();
    if (mode == 2) {
      for (int r = 0; r < reps; ++r) mma_tf32_ss(tmem, da + 2 * (r & 3), db + 2 * (r & 3), idesc, r > 0);
    } else if (mode >= 3) {
      // mode = 3 + log2(n_acc) + 8 * (M == 64): rotate over n_acc independent accumulators (64 columns apart)
      const int n_acc = 1 << ((mode - 3) & 7);
      const uint32_t id2 = idesc_tf32((mode - 3) & 8 ? 64 : 128, n_mma);
      for (int r = 0; r < reps; ++r)
        mma_tf32_ss(tmem + 64 * (r & (n_acc - 1)), da + 2 * (r & 3), db + 2 * (r & 3), id2, r >= n_acc);
    } else {
      for (int r = 0; r < reps; ++r)
        for (int ks = 0; ks < 4; ++ks) mma_tf32_ss(tmem, da + 2 * ks, db + 2 * ks, idesc, (r | ks) > 0);
    }
    mma_commit(&sm.bar);
  }
  mbar_wait(&sm.bar, 0);
  if (tid == 0) { t1 = clock64(); clk[0] = t1 - t0; }
  fence_after_sync();
  uint32_t v[16];
  for (int c = 0; c < N; c += 16) {
    tmem_ld_x16(tmem + ((uint32_t)(32 * wid) << 16) + c, v);
    tmem_ld_wait();
    for (int j = 0; j < 16; ++j) D[(32 * wid + lane) * N + c + j] = __uint_as_float(v[j]);
  }
  fence_before_sync();
  __syncthreads();
  if (wid == 0) tmem_free(tmem, 512);
}

static float tf32_trunc(float x) { uint32_t u; memcpy(&u, &x, 4); u &= 0xFFFFE000u; memcpy(&x, &u, 4); return x; }

int main() {
  std::vector<float> A(M * 32), B(64 * 32, 0.f), D(M * N);
  float *dA, *dB, *dD; long long* dclk;
  cudaMalloc(&dA, A.size() * 4); cudaMalloc(&dB, B.size() * 4); cudaMalloc(&dD, D.size() * 4); cudaMalloc(&dclk, 8);
  const size_t smem = sizeof(Smem) + 1024;
  cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  srand(1);
  auto rnd = []() { return (float)rand() / RAND_MAX; };

  // ---- 1. layout / descriptor correctness with tf32-exact inputs
  for (auto& x : A) x = tf32_trunc(rnd() * 2.f - 1.f);
  for (int i = 0; i < N * 32; ++i) B[i] = tf32_trunc(rnd() * 2.f - 1.f);
  cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice);
  probe_kernel<<<1, 128, smem>>>(dA, dB, dD, 1, 0, N, dclk);
  cudaError_t e = cudaDeviceSynchronize();
  printf("probe1 launch: %s\n", cudaGetErrorString(e));
  if (e != cudaSuccess) return 1;
  cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
  double worst = 0;
  for (int m = 0; m < M; ++m)
    for (int n = 0; n < N; ++n) {
      double s = 0;
      for (int k = 0; k < 32; ++k) s += (double)A[m * 32 + k] * (double)B[n * 32 + k];
      worst = fmax(worst, fabs(s - D[m * N + n]));
    }
  printf("probe1 layout: max |D - A.B^T| = %.3e  (expect ~1e-6)  %s\n", worst, worst < 1e-4 ? "OK" : "MISMATCH");

  // ---- 1b. is N = 40 (multiple of 8, not of 16) a legal shape at M = 128?
  {
    for (int i = 0; i < N * 32; ++i) B[i] = tf32_trunc(rnd() * 2.f - 1.f);   // fresh operands: stale TMEM cannot pass
    cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice);
    probe_kernel<<<1, 128, smem>>>(dA, dB, dD, 1, 0, 40, dclk);
    cudaError_t e2 = cudaDeviceSynchronize();
    printf("probe1b N=40 launch: %s\n", cudaGetErrorString(e2));
    if (e2 == cudaSuccess) {
      cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
      double w2 = 0;
      for (int m = 0; m < M; ++m)
        for (int n = 0; n < 40; ++n) {
          double s = 0;
          for (int k = 0; k < 32; ++k) s += (double)A[m * 32 + k] * (double)B[n * 32 + k];
          w2 = fmax(w2, fabs(s - D[m * N + n]));
        }
      printf("probe1b N=40: max |D - A.B^T| over 128x40 = %.3e %s\n", w2, w2 < 1e-4 ? "OK" : "MISMATCH");
    } else return 1;
  }

  // ---- 2a. input rounding: A[0,0] = 1 + 2^-11 + 2^-13, B[0,0] = 1, everything else 0
  std::fill(A.begin(), A.end(), 0.f); std::fill(B.begin(), B.end(), 0.f);
  A[0] = 1.f + ldexpf(1.f, -11) + ldexpf(1.f, -13); B[0] = 1.f;
  A[32] = 1.f + ldexpf(1.f, -11) - ldexpf(1.f, -13); B[32] = 1.f;   // row 1 x col 1 uses k=0 too
  B[32 + 0] = 1.f;
  cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice);
  probe_kernel<<<1, 128, smem>>>(dA, dB, dD, 1, 0, N, dclk);
  cudaDeviceSynchronize();
  cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
  printf("probe2a input rounding: (1+2^-11+2^-13)*1 -> 1 + %.3f * 2^-10 ; (1+2^-11-2^-13)*1 -> 1 + %.3f * 2^-10  (0 = truncation)\n",
         (D[0] - 1.0) * 1024.0, (D[1 * N + 0] - 1.0) * 1024.0);

  // ---- 2b. accumulator rounding: positive tf32 inputs, many accumulation steps, signed error vs exact
  for (auto& x : A) x = tf32_trunc(0.5f + rnd());
  for (int i = 0; i < 64 * 32; ++i) B[i] = i < N * 32 ? tf32_trunc(0.5f + rnd()) : 0.f;
  cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice);
  for (int reps : {1, 2, 4, 8, 16, 64, 256}) {
    probe_kernel<<<1, 128, smem>>>(dA, dB, dD, reps, 1, N, dclk);
    cudaDeviceSynchronize();
    cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
    double mean = 0, rms = 0;
    for (int m = 0; m < M; ++m)
      for (int n = 0; n < N; ++n) {
        double s = 0;
        for (int k = 0; k < 32; ++k) s += (double)A[m * 32 + k] * (double)B[n * 32 + k];
        s *= reps;
        double rel = (D[m * N + n] - s) / s;
        mean += rel; rms += rel * rel;
      }
    mean /= M * N; rms = sqrt(rms / (M * N));
    printf("probe2b accumulate %4d tiles (%5d mma steps): mean rel err %+.3e  rms %.3e\n", reps, reps * 4, mean, rms);
  }

  // ---- 3. cycles per MMA
  for (int n_mma : {16, 32, 48, 64, 96, 128, 256}) {
    
    long long clk = 0;
    for (int reps : {256, 2048}) {
      probe_kernel<<<1, 128, smem>>>(dA, dB, dD, reps, 2, n_mma, dclk);
      cudaDeviceSynchronize();
      cudaMemcpy(&clk, dclk, 8, cudaMemcpyDeviceToHost);
      printf("probe3 M=128 N=%3d: %5d MMAs in %8lld clk -> %.1f clk/MMA\n", n_mma, reps, clk, (double)clk / reps);
    }
  }
  for (int m64 = 0; m64 < 2; ++m64)
    for (int la = 0; la < 4; ++la)
      for (int n_mma : {16, 48, 64}) {
        long long clk = 0;
        probe_kernel<<<1, 128, smem>>>(dA, dB, dD, 2048, 3 + la + 8 * m64, n_mma, dclk);
        cudaDeviceSynchronize();
        cudaMemcpy(&clk, dclk, 8, cudaMemcpyDeviceToHost);
        printf("probe4 M=%3d N=%3d n_acc=%d: %.1f clk/MMA\n", m64 ? 64 : 128, n_mma, 1 << la, (double)clk / 2048);
      }
  for (int nw = 1; nw <= 4; ++nw)
    for (int n_mma : {48, 128, 256}) {
      long long clk = 0;
      probe_kernel<<<1, 128, smem>>>(dA, dB, dD, 1024, 100 + nw, n_mma, dclk);
      cudaDeviceSynchronize();
      cudaMemcpy(&clk, dclk, 8, cudaMemcpyDeviceToHost);
      printf("probe5 M=128 N=%3d issuing warps=%d: %.1f clk/MMA (aggregate)\n", n_mma, nw, (double)clk / (1024.0 * nw));
    }
  return 0;
}
