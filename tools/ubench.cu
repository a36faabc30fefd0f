// Pipe-throughput micro-benchmarks used to set compute roofs for the bin kernel:
//   DFMA, FFMA, MUFU.EX2, libdevice exp(double), LDS.64 broadcast.  Build: nvcc -arch=sm_100a -O3 -o ubench ubench.cu
#include <cstdio>
#include <cuda_runtime.h>
template <int OP>
__global__ void k(double* out, int iters) {
  double a[8]; float f[8];
  for (int i = 0; i < 8; ++i) { a[i] = threadIdx.x * 1e-3 + i; f[i] = (float)a[i]; }
  double c = 1.000001, d = 1e-7;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (OP == 0) a[i] = fma(a[i], c, d);
      if (OP == 1) f[i] = fmaf(f[i], 1.000001f, 1e-7f);
      if (OP == 2) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(f[i])); f[i] = y * 1e-3f; }
      if (OP == 3) a[i] = exp(-fabs(a[i]) * 1e-3) + i;
    }
  }
  double s = 0; for (int i = 0; i < 8; ++i) s += a[i] + f[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int OP> void run(const char* name, int iters, double ops_per_iter) {
  double* out; cudaMalloc(&out, 148 * 8 * 256 * 8);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<OP><<<148 * 8, 256>>>(out, iters); cudaDeviceSynchronize();
  cudaEventRecord(e0); k<OP><<<148 * 8, 256>>>(out, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  double total = 148.0 * 8 * 256 * iters * ops_per_iter;
  printf("%-12s %8.3f ms  %10.2f Gop/s  (%.1f op/clk/SM @1.965GHz)\n", name, ms, total / ms * 1e-6, total / (ms * 1e-3) / 148 / 1.965e9);
  cudaFree(out);
}
int main() {
  run<0>("DFMA", 4096, 8); run<1>("FFMA", 4096, 8); run<2>("MUFU.EX2", 4096, 8); run<3>("exp(f64)", 512, 8);
  return 0;
}
