// DFMA throughput vs resident warps per SM and ILP (independent chains per thread).
#include <cstdio>
#include <cuda_runtime.h>
template <int ILP>
__global__ void k(double* out, int iters) {
  double a[ILP];
  for (int i = 0; i < ILP; ++i) a[i] = threadIdx.x * 1e-3 + i;
  double c = 1.000001, d = 1e-7;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i) a[i] = fma(a[i], c, d);
  }
  double s = 0; for (int i = 0; i < ILP; ++i) s += a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int ILP> void run(int blocks_per_sm, int threads) {
  double* out; cudaMalloc(&out, 148 * 32 * 1024 * 8);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  int iters = 1 << 14;
  k<ILP><<<148 * blocks_per_sm, threads>>>(out, iters); cudaDeviceSynchronize();
  cudaEventRecord(e0); k<ILP><<<148 * blocks_per_sm, threads>>>(out, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  double total = 148.0 * blocks_per_sm * threads * (double)iters * ILP;
  printf("warps/SM %3d  ILP %2d : %6.1f DFMA/clk/SM\n", blocks_per_sm * threads / 32, ILP, total / (ms * 1e-3) / 148 / 1.965e9);
  cudaFree(out);
}
int main() {
  run<1>(1, 128); run<2>(1, 128); run<4>(1, 128); run<8>(1, 128); run<16>(1, 128);
  run<1>(2, 128); run<2>(2, 128); run<4>(2, 128); run<8>(2, 128); run<16>(2, 128);
  run<1>(4, 128); run<4>(4, 128); run<8>(4, 128);
  run<1>(8, 256); run<4>(8, 256); run<8>(8, 256);
  return 0;
}
