// gcs_tc.cuh -- sm_100a tensor-core plumbing used by the tensor-core variant of the bin-moment kernel:
// tcgen05.mma (kind::tf32, operands in shared memory, accumulator in tensor memory), TMEM allocation and loads,
// mbarriers, and the K-major SWIZZLE_128B shared-memory operand layout.  Raw PTX, no library templates.
//
// Operand layout (both A and B are K-major, K = points): a tile is `rows` x 32 tf32 values (128 B per row).  Rows are
// stored 128 B apart, groups of 8 rows 1024 B apart (SBO), and the eight 16-byte chunks of a row are XOR-permuted by
// (row & 7) -- the 128-byte swizzle the tensor core expects.  A warp that writes one value per lane (lane = point k)
// into a fixed row touches every bank exactly once.  One tcgen05.mma consumes K = 8 points (32 B of every row); a
// 32-point tile is four MMAs whose descriptors advance the start address by 32 B.
#pragma once
#include <cuda_fp16.h>
#include <stdint.h>

namespace gcs {
namespace tc {

constexpr int kTileK = 32;           // points per operand tile (one 128-byte swizzle row)
constexpr int kRowBytes = 128;
constexpr int kGroupBytes = 1024;    // 8 rows

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// byte offset of element (row r, point k) inside a swizzled tile whose base is 1024-byte aligned
__device__ __forceinline__ uint32_t swz_offset(int r, int k) {
  return (uint32_t)((r >> 3) * kGroupBytes + (r & 7) * kRowBytes + ((((k >> 2) ^ (r & 7)) & 7) << 4) + ((k & 3) << 2));
}

// shared-memory matrix descriptor: K-major, SWIZZLE_128B, 8-row groups `sbo_bytes` apart, descriptor version 1
__device__ __forceinline__ uint64_t smem_desc_sw128(uint32_t smem_addr, uint32_t sbo_bytes = kGroupBytes) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)1 << 16;                                  // leading byte offset: unused for swizzled K-major
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
  d |= (uint64_t)1 << 46;                                  // descriptor version (sm_100)
  d |= (uint64_t)2 << 61;                                  // SWIZZLE_128B
  return d;
}

// ---- 16-bit operands (kind::f16): K-major rows of 32 points = 64 bytes, SWIZZLE_64B: groups of 8 rows 512 B apart, the
// four 16-byte chunks of a row XOR-permuted by bits 1-2 of the row index.  One tcgen05.mma consumes K = 16 points (32 B).
constexpr int kRowBytes16 = 64;
constexpr int kGroupBytes16 = 512;
__device__ __forceinline__ uint32_t swz64_offset(int r, int byte_in_row) {
  return (uint32_t)((r >> 3) * kGroupBytes16 + (r & 7) * kRowBytes16 + ((((byte_in_row >> 4) ^ (r >> 1)) & 3) << 4) +
                    (byte_in_row & 15));
}
__device__ __forceinline__ uint64_t smem_desc_sw64(uint32_t smem_addr, uint32_t sbo_bytes = kGroupBytes16) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)4 << 61;                                  // SWIZZLE_64B
  return d;
}
// ---- MN-major, unswizzled ("interleaved") operands: the 8 x 8 core matrix is 8 K rows of 16 bytes (8 consecutive M / N
// elements of 16 bits) = 128 contiguous bytes; M / N blocks of 8 elements are `sbo_bytes` apart, groups of 8 K rows
// `lbo_bytes` apart (tools/tc_probe_mn.cu: the other assignment faults).  One kind::f16 MMA (K = 16) reads two K groups.
__device__ __forceinline__ uint64_t smem_desc_mn(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
  d |= (uint64_t)1 << 46;                                  // descriptor version (sm_100); layout type 0: no swizzle
  return d;
}
constexpr uint32_t kIdescMnMajorA = 1u << 15, kIdescMnMajorB = 1u << 16;
// instruction descriptor, kind::f16: D f32, A/B f16 (format 0), both K-major, dense
__host__ __device__ constexpr uint32_t idesc_f16(int M, int N) {
  return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void mma_f16_ss(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                           uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}

// instruction descriptor, kind::tf32: D f32, A/B tf32, both K-major, dense
__host__ __device__ constexpr uint32_t idesc_tf32(int M, int N) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

__device__ __forceinline__ void mma_tf32_ss(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}

// all previously issued MMAs of this thread arrive (once) on the mbarrier when they have completed
__device__ __forceinline__ void mma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}

// one lane of a converged warp (the same lane every time for the same mask): lets the warp run issue code with
// warp-uniform operands -- tcgen05 instructions take uniform registers, and a thread-divergent `if (lane == 0)` around them
// makes the compiler wrap every one in a vote / R2UR loop
__device__ __forceinline__ bool elect_one() {
  uint32_t p;
  asm volatile(
      "{\n\t"
      ".reg .pred q;\n\t"
      "elect.sync _|q, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, q;\n\t"
      "}\n"
      : "=r"(p));
  return p != 0;
}

__device__ __forceinline__ void fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// generic-proxy shared-memory writes -> visible to the async proxy (tensor core operand reads)
__device__ __forceinline__ void fence_smem_to_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// whole warp; writes the TMEM base address to *holder (shared memory).  n_cols: power of two >= 32.
__device__ __forceinline__ void tmem_alloc(uint32_t* holder, uint32_t n_cols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(holder)), "r"(n_cols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_free(uint32_t addr, uint32_t n_cols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(n_cols) : "memory");
}

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_init_fence() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t"
      "}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {}
}
// non-blocking phase test (polling loops that watch several barriers)
__device__ __forceinline__ bool mbar_test_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t"
      "}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// TMEM -> registers: lane i of the warp reads TMEM lane (lane field of addr) + i, N consecutive 32-bit columns.
__device__ __forceinline__ void tmem_ld_x2(uint32_t addr, uint32_t (&v)[2]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x2.b32 {%0, %1}, [%2];" : "=r"(v[0]), "=r"(v[1]) : "r"(addr));
}
__device__ __forceinline__ void tmem_ld_x4(uint32_t addr, uint32_t (&v)[4]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3])
               : "r"(addr));
}
__device__ __forceinline__ void tmem_ld_x8(uint32_t addr, uint32_t (&v)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
               : "r"(addr));
}
__device__ __forceinline__ void tmem_ld_x16(uint32_t addr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(addr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// f32 -> (hi, lo) with hi exactly representable in tf32 (round to nearest, ties away) and lo = x - hi exact in f32.
// The tensor core ignores the 13 low mantissa bits of its tf32 inputs; hi + lo carries ~22 significant bits.
__device__ __forceinline__ void split_tf32(float x, float& hi, float& lo) {
  hi = __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xFFFFE000u);
  lo = x - hi;
}

// packed float32 pairs (sm_100 FFMA2 / FADD2): one instruction, two lanes of math
__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c) {
  float2 d;
  asm("{\n\t.reg .b64 ra, rb, rc, rd;\n\tmov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\tmov.b64 rc, {%6, %7};\n\t"
      "fma.rn.f32x2 rd, ra, rb, rc;\n\tmov.b64 {%0, %1}, rd;\n\t}"
      : "=f"(d.x), "=f"(d.y)
      : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y), "f"(c.x), "f"(c.y));
  return d;
}
__device__ __forceinline__ float2 add2(float2 a, float2 b) {
  float2 d;
  asm("{\n\t.reg .b64 ra, rb, rd;\n\tmov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\t"
      "add.rn.f32x2 rd, ra, rb;\n\tmov.b64 {%0, %1}, rd;\n\t}"
      : "=f"(d.x), "=f"(d.y)
      : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
  return d;
}
__device__ __forceinline__ float2 sub2(float2 a, float2 b) {
  float2 d;
  asm("{\n\t.reg .b64 ra, rb, rd;\n\tmov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\t"
      "sub.rn.f32x2 rd, ra, rb;\n\tmov.b64 {%0, %1}, rd;\n\t}"
      : "=f"(d.x), "=f"(d.y)
      : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
  return d;
}
__device__ __forceinline__ float2 mul2(float2 a, float2 b) {
  float2 d;
  asm("{\n\t.reg .b64 ra, rb, rd;\n\tmov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\t"
      "mul.rn.f32x2 rd, ra, rb;\n\tmov.b64 {%0, %1}, rd;\n\t}"
      : "=f"(d.x), "=f"(d.y)
      : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
  return d;
}
// two float32 -> packed fp16 pair (lo in bits 0-15, hi in bits 16-31), round to nearest; and back (exact)
__device__ __forceinline__ uint32_t pack_f16x2(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}
__device__ __forceinline__ float2 unpack_f16x2(uint32_t v) {
  return __half22float2(*reinterpret_cast<const __half2*>(&v));
}
// byte offset of row r of a 16-bit operand tile (before the in-row swizzled offset)
__host__ __device__ constexpr int row_base16(int r) { return (r >> 3) * kGroupBytes16 + (r & 7) * kRowBytes16; }
__device__ __forceinline__ float ex2f(float x) {
  float e;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(x));
  return e;
}
__device__ __forceinline__ float tf32_hi(float x) { return __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xFFFFE000u); }

}  // namespace tc
}  // namespace gcs
