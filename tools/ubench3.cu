// Issue-rate micro-benchmarks for the instruction mix of the tensor-core bin kernel's producer warps (sm_100a):
// packed f32x2 math, 3-input max, fp16x2 pack / unpack, MUFU.EX2 beside packed FMAs, 16-byte shared-memory stores.
// Prints warp-instructions per clock per SM (4 schedulers: 4.0 is the issue limit).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ubench3 ubench3.cu
#include <cstdio>
#include <cstdint>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c) {
  float2 d;
  asm volatile("{\n\t.reg .b64 ra, rb, rc, rd;\n\tmov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\tmov.b64 rc, {%6, %7};\n\t"
               "fma.rn.f32x2 rd, ra, rb, rc;\n\tmov.b64 {%0, %1}, rd;\n\t}"
               : "=f"(d.x), "=f"(d.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y), "f"(c.x), "f"(c.y));
  return d;
}
__device__ __forceinline__ float2 add2(float2 a, float2 b) {
  float2 d;
  asm volatile("{\n\t.reg .b64 ra, rb, rd;\n\tmov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\t"
               "add.rn.f32x2 rd, ra, rb;\n\tmov.b64 {%0, %1}, rd;\n\t}"
               : "=f"(d.x), "=f"(d.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
  return d;
}

template <int OP>
__global__ void k(float* out, int iters) {
  extern __shared__ uint4 sm[];
  float2 a[8];
  for (int i = 0; i < 8; ++i) a[i] = make_float2(threadIdx.x * 1e-3f + i, threadIdx.x * 2e-3f + i);
  const float2 c = make_float2(1.000001f, 0.999999f), d = make_float2(1e-7f, 2e-7f);
  uint32_t acc = 0;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (OP == 0) a[i] = fma2(a[i], c, d);                                   // FFMA2
      if (OP == 1) a[i] = add2(a[i], d);                                      // FADD2
      if (OP == 2) { a[i].x = fmaf(a[i].x, c.x, d.x); a[i].y = fmaf(a[i].y, c.y, d.y); }   // 2 x FFMA
      if (OP == 3) { a[i].x = fmaxf(a[i].x, a[(i + 1) & 7].y); }              // FMNMX
      if (OP == 4) { float r; asm volatile("max.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a[i].x), "f"(a[(i + 1) & 7].y), "f"(a[(i + 2) & 7].x)); a[i].x = r; }  // FMNMX3
      if (OP == 5) { uint32_t r; asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(a[i].y), "f"(a[i].x)); acc ^= r; a[i].x += 1.0f; }   // F2FP + FADD
      if (OP == 6) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(a[i].x)); a[i].x = y * 1e-3f; }   // MUFU + FMUL
      if (OP == 7) {   // the soft-assign mix per bin pair: 5 packed FMA-pipe ops + 2 MUFU + 1 FMNMX3 + 2 F2FP
        float2 l = fma2(a[i], c, fma2(a[i], d, fma2(a[i], c, d)));
        float e0, e1;
        asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(e0) : "f"(l.x));
        asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(e1) : "f"(l.y));
        float2 e = make_float2(e0, e1);
        a[(i + 1) & 7] = add2(a[(i + 1) & 7], e);
        a[(i + 2) & 7] = fma2(e, l, a[(i + 2) & 7]);
        uint32_t h; asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(h) : "f"(e1), "f"(e0));
        float2 hf = __half22float2(*reinterpret_cast<__half2*>(&h));
        float2 r = add2(e, make_float2(-hf.x, -hf.y));
        uint32_t lo; asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(lo) : "f"(r.y), "f"(r.x));
        acc ^= h ^ lo;
      }
      if (OP == 8) { sm[threadIdx.x * 9 + i] = make_uint4(acc, it, i, 0); acc += it; }   // STS.128 (conflict-free stride)
    }
  }
  float s = 0; for (int i = 0; i < 8; ++i) s += a[i].x + a[i].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = s + __uint_as_float(acc);
}
template <int OP> void run(const char* name, int iters, double winstr_per_iter) {
  float* out; cudaMalloc(&out, 148 * 4 * 256 * 4);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int smem = 256 * 9 * 16 + 1024;
  cudaFuncSetAttribute(k<OP>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  k<OP><<<148 * 4, 256, smem>>>(out, iters); cudaDeviceSynchronize();
  cudaEventRecord(e0); k<OP><<<148 * 4, 256, smem>>>(out, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  double total = 148.0 * 4 * 8 * iters * winstr_per_iter;   // warp instructions of the named kind
  printf("%-28s %8.3f ms  %6.2f warp-instr/clk/SM @1.965GHz (%s)\n", name, ms, total / (ms * 1e-3) / 148 / 1.965e9,
         cudaGetErrorString(cudaGetLastError()));
  cudaFree(out);
}
int main() {
  run<0>("FFMA2", 4096, 8); run<1>("FADD2", 4096, 8); run<2>("FFMA x2 (scalar)", 4096, 16); run<3>("FMNMX", 4096, 8);
  run<4>("FMNMX3", 4096, 8); run<5>("F2FP.f16x2 (+FADD)", 4096, 16); run<6>("MUFU.EX2 (+FMUL)", 4096, 16);
  run<7>("soft-assign mix (17/pair)", 2048, 8 * 17); run<8>("STS.128", 2048, 8);
  return 0;
}
