// msort_umma.cuh — the few tcgen05 / TMEM primitives the kernels of this library use (sm_100a), written as inline
// PTX: shared-memory matrix descriptors for the canonical no-swizzle K-major layout, the single-thread MMA issue,
// commit -> mbarrier, TMEM allocation and tcgen05.ld.  Shared by the rollout policy kernel (msort_policy.cu), the
// fused rollout kernel and the step kernel's embedded-policy path (msort_kernels.cu).
//
// Operand layout used everywhere (fp16 operands, fp32 accumulation, M = 128 rows = one tile of envs):
//   a [rows x K] K-major operand is stored as 16-byte chunks of 8 consecutive K elements; chunk kc of row r sits at
//   byte offset (kc * rows + r) * 16.  Eight consecutive rows of one chunk column form a 128-byte "core matrix"
//   (SBO = 128 B between 8-row groups), the next K chunk starts rows*16 bytes later (LBO).  One tcgen05.mma
//   kind::f16 instruction consumes K = 16 = two chunks.
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace msort {
namespace umma {

// shared-memory matrix descriptor, no swizzle, K-major (descriptor version 1 = sm_100)
__device__ __forceinline__ uint64_t smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  return (uint64_t)((saddr >> 4) & 0x3fffu) | ((uint64_t)((lbo_bytes >> 4) & 0x3fffu) << 16) |
         ((uint64_t)((sbo_bytes >> 4) & 0x3fffu) << 32) | (1ull << 46);
}

// instruction descriptor: D fp32, A/B fp16, both K-major, M = 128, N = n (multiple of 16)
__device__ __forceinline__ constexpr uint32_t idesc_f16(int n) {
  return (1u << 4) | (0u << 7) | (0u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}

// D[tmem] (+)= A[smem] * B[smem]; issued by ONE thread
__device__ __forceinline__ void mma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      :: "r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}

// arrive on `bar` once every MMA this thread has issued so far is complete (implies fence::before_thread_sync)
__device__ __forceinline__ void commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];"
               :: "r"((uint32_t)__cvta_generic_to_shared(bar)) : "memory");
}

__device__ __forceinline__ void fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// generic-proxy shared-memory writes -> visible to the async proxy (MMA operand reads, bulk copies)
__device__ __forceinline__ void fence_async_proxy() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// one whole warp allocates `cols` (power of two >= 32) TMEM columns; the base address lands in *slot (shared memory)
__device__ __forceinline__ void tmem_alloc(uint32_t* slot, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
               :: "r"((uint32_t)__cvta_generic_to_shared(slot)), "r"(cols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t base, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(base), "r"(cols) : "memory");
}

// tcgen05.ld 32x32b: thread (lane l of warp w) reads `n` consecutive 32-bit columns of TMEM lane 32*(w%4)+l.
// taddr = base + (lane_row << 16) + column.  Whole-warp (.sync.aligned) instructions.
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr) : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
}

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr) : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(r[j]);
}

__device__ __forceinline__ void tmem_ld2(uint32_t taddr, float& v0, float& v1) {
  uint32_t r0, r1;
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x2.b32 {%0,%1}, [%2];" : "=r"(r0), "=r"(r1) : "r"(taddr) : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
  v0 = __uint_as_float(r0); v1 = __uint_as_float(r1);
}

// eight fp32 values -> one 16-byte chunk of fp16 (a K-major core-matrix row)
__device__ __forceinline__ uint4 pack8(const float (&v)[8]) {
  const __half2 a = __floats2half2_rn(v[0], v[1]), b = __floats2half2_rn(v[2], v[3]);
  const __half2 c = __floats2half2_rn(v[4], v[5]), d = __floats2half2_rn(v[6], v[7]);
  return make_uint4(*reinterpret_cast<const uint32_t*>(&a), *reinterpret_cast<const uint32_t*>(&b),
                    *reinterpret_cast<const uint32_t*>(&c), *reinterpret_cast<const uint32_t*>(&d));
}

// packed fp32 add (Blackwell FADD2: two independent IEEE adds per lane and issue slot)
__device__ __forceinline__ void fadd2(float a0, float a1, float b0, float b1, float& c0, float& c1) {
  asm("{\n\t.reg .b64 ra, rb, rc;\n\tmov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\tadd.rn.f32x2 rc, ra, rb;\n\tmov.b64 {%0, %1}, rc;\n\t}"
      : "=f"(c0), "=f"(c1) : "f"(a0), "f"(a1), "f"(b0), "f"(b1));
}
__device__ __forceinline__ void fsub2(float a0, float a1, float b0, float b1, float& c0, float& c1) {
  fadd2(a0, a1, -b0, -b1, c0, c1);     // the negations fold into the instruction's operand modifiers
}

// Two-term fp16 split of a pair of fp32 values: hi = fp16(x), lo = fp16(x - hi) — hi + lo carries 22 significand bits of
// x (absolute error <= 2^-25 for |x| <= 1 down to fp16's subnormal spacing 2^-24), so an fp16 tensor-core product
// against a three-term weight split reproduces the fp32 product to fp32 accuracy.  F2FP + 2 HADD2.F32 + 2 FADD + F2FP.
__device__ __forceinline__ void split2(float a, float b, uint32_t& hi, uint32_t& lo) {
  const __half2 h = __floats2half2_rn(a, b);
  const float2 nf = __half22float2(__hneg2(h));   // -hi (the sign flip folds into the conversion's operand modifier)
  float ra, rb;
  fadd2(a, b, nf.x, nf.y, ra, rb);
  const __half2 l = __floats2half2_rn(ra, rb);
  hi = *reinterpret_cast<const uint32_t*>(&h);
  lo = *reinterpret_cast<const uint32_t*>(&l);
}

}  // namespace umma
}  // namespace msort
