// msort_kernels.cu — hand-written sm_100a kernels: fused step (K1), reset (K2), observe,
// state export/import (K6), state statistics (K5).  One thread = one env; one CTA = one tile
// of 128 consecutive envs.  Row-major obs[N,D] / mask[N,A] are transposed through shared
// memory so every global store is a coalesced 4-byte-per-lane (128 B per warp) store.
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstring>
#include <type_traits>

#include "msort_device.cuh"
#include "msort_launch.h"
#include "msort_umma.cuh"

#ifndef MSORT_PREFETCH_TILES
#define MSORT_PREFETCH_TILES 296  // L2 prefetch distance of the step kernel in tiles (2 per SM; 0 = off): 148..518 measured alike, +3.4 %
#endif
#ifndef MSORT_TC_PERSIST
#define MSORT_TC_PERSIST 0        // Env_2's tensor-core kernel: 0 = one CTA per tile (weights by TMA per CTA; measured 104.6 vs 109.7 us per 1 048 576 envs), 1 = persistent like the FFMA2 form
#endif
#ifndef MSORT_HOT_PERSIST
#define MSORT_HOT_PERSIST 1       // Env_2's HOT step kernel: 1 = resident CTAs loop over tiles, next tile's state staged by TMA
#endif
#ifndef MSORT_PRESS_MIN_BLOCKS
#define MSORT_PRESS_MIN_BLOCKS 5  // Env_2 (embedded MLP: 32 activations + FFMA2 accumulator pairs in registers)
#endif
#ifndef MSORT_EXP_SIGMOID
#define MSORT_EXP_SIGMOID 0
#endif
#ifndef MSORT_TC_ROLL
#define MSORT_TC_ROLL 0           // experiment switch: the two 16-column halves of each MLP epilogue as a rolled loop (smaller footprint; measured 106.1 vs 104.4 us: slower)
#endif
#ifndef MSORT_PRESS_TC_MIN_BLOCKS
#define MSORT_PRESS_TC_MIN_BLOCKS 8  // Env_2 with the embedded policy on the tensor cores (TCMLP): 27.1 KB of shared memory per CTA, 64 registers
#endif
#ifndef MSORT_HOT_MONO_MIN_BLOCKS
#define MSORT_HOT_MONO_MIN_BLOCKS 8  // Env_3's HOT kernel fits 64 registers without a spill: 8 CTAs (32 warps) per SM, +3..5 %
#endif
#ifndef MSORT_FUSE_MIN_BLOCKS
#define MSORT_FUSE_MIN_BLOCKS 7  // Env_3's HOT kernel fused with the rollout policy: 30.4 KB of shared memory per CTA, 72 registers
#endif
#ifndef MSORT_HOT_SORT_MIN_BLOCKS
#define MSORT_HOT_SORT_MIN_BLOCKS 7  // Env_1's HOT kernel (72 registers)
#endif
#ifndef MSORT_STEP_MIN_BLOCKS
#define MSORT_STEP_MIN_BLOCKS 7  // resident CTAs per SM the step kernel is compiled for (register cap)
#endif

namespace msort {

template <int KIND> struct Dims;
template <> struct Dims<MSORT_ENV_SORT> { static constexpr int D = 13, A = 2; };
template <> struct Dims<MSORT_ENV_PRESS> { static constexpr int D = 16, A = 11; };
template <> struct Dims<MSORT_ENV_MONO> { static constexpr int D = 29, A = 22; };

// ---------------------------------------------------------------- tile output helpers
// The CTA's obs rows / mask rows are staged in shared memory in exactly the row-major layout of
// the global tensors, so the flush is a straight 16-byte-vector copy of a contiguous range
// (128 rows * D * 4 B and 128 * A B are multiples of 16 for every env kind).
__device__ __forceinline__ void flush_tile(const void* __restrict__ tile, void* __restrict__ dst, int bytes) {
  const int n16 = bytes >> 4;
  const uint4* s4 = reinterpret_cast<const uint4*>(tile);
  uint4* d4 = reinterpret_cast<uint4*>(dst);
  for (int e = threadIdx.x; e < n16; e += kTile) d4[e] = s4[e];
  const uint8_t* sb = reinterpret_cast<const uint8_t*>(tile);
  uint8_t* db = reinterpret_cast<uint8_t*>(dst);
  for (int e = (n16 << 4) + threadIdx.x; e < bytes; e += kTile) db[e] = sb[e];  // partial last tile only
}

// Full tiles leave through the TMA engine: one elected thread issues a bulk shared->global copy
// (cp.async.bulk, SASS UBLKCP) of the whole contiguous tile.  Writers must have executed
// fence.proxy.async + a CTA barrier before; the issuing thread waits until the engine has read
// the tile so the CTA's shared memory outlives the copy.
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void bulk_store(void* __restrict__ dst, const void* __restrict__ tile, uint32_t bytes) {
  const uint32_t src = (uint32_t)__cvta_generic_to_shared(tile);
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(src), "r"(bytes) : "memory");
}

__device__ __forceinline__ void bulk_commit_and_wait_read() {
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
  asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}

// every byte of x -> 0xff if its bit 7 is set, else 0x00 (PRMT with sign-replicating selectors;
// __byte_perm masks those selector bits away, hence the PTX)
__device__ __forceinline__ uint32_t byte_sign_mask(uint32_t x) {
  uint32_t m;
  asm("prmt.b32 %0, %1, %1, 0xba98;" : "=r"(m) : "r"(x));
  return m;
}

// bits 0..3 of x -> bytes 0..3 (0/1 each): bit i lands at 8i through the 2^(7i) term
__device__ __forceinline__ uint32_t spread4(uint32_t x) { return (x * 0x00204081u) & 0x01010101u; }

// Write one env's action-mask row (A bytes) into the dense shared tile.
template <int A>
__device__ __forceinline__ void put_mask_row(uint8_t* __restrict__ tile, int t, uint32_t bits) {
  if (A == 2) {
    reinterpret_cast<uint16_t*>(tile)[t] = 0x0101;
  } else {
    const uint32_t w0 = spread4(bits & 0xfu), w1 = spread4((bits >> 4) & 0xfu), w2 = spread4((bits >> 8) & 0x7u);
    if (A == 11) {
      uint8_t* r = tile + 11 * t;
#pragma unroll
      for (int k = 0; k < 4; ++k) { r[k] = (uint8_t)(w0 >> (8 * k)); r[4 + k] = (uint8_t)(w1 >> (8 * k)); }
#pragma unroll
      for (int k = 0; k < 3; ++k) r[8 + k] = (uint8_t)(w2 >> (8 * k));
    } else {  // A == 22: the 11 press bytes twice (monolith_action_masks env_super.py:887-898); row is 2-byte aligned
      uint16_t* r = reinterpret_cast<uint16_t*>(tile + 22 * t);
      r[0] = (uint16_t)w0; r[1] = (uint16_t)(w0 >> 16); r[2] = (uint16_t)w1; r[3] = (uint16_t)(w1 >> 16);
      r[4] = (uint16_t)w2;
      r[5] = (uint16_t)((w2 >> 16) | (w0 << 8));
      r[6] = (uint16_t)(w0 >> 8);
      r[7] = (uint16_t)((w0 >> 24) | (w1 << 8));
      r[8] = (uint16_t)(w1 >> 8);
      r[9] = (uint16_t)((w1 >> 24) | (w2 << 8));
      r[10] = (uint16_t)(w2 >> 8);
    }
  }
}

// ---------------------------------------------------------------- K1: fused step
struct StepArgs {
  uint4* state;
  const long long* actions;
  float* obs;
  float* reward;
  uint8_t* terminated;
  uint8_t* mask;  // nullable
  // info (nullable)
  long long* info_action;
  uint8_t* info_overflow;
  int8_t* info_overflow_mat;
  uint8_t* info_sort_mode;
  uint8_t* info_press_action;
  uint8_t* info_invalid;
  uint32_t* info_sorted_true;
  float* info_r_sort;
  float* info_r_press;
  int any_step_info;  // any of the six per-step info arrays above is present
  int act_tma;        // actions are 16-byte aligned: the persistent kernel may fetch a tile's actions by TMA
  const uint4* policy_tc;   // Env_2 TCMLP: packed tensor-core policy (kTcWords words, device memory) or nullptr
  uint8_t* mode_scratch;    // Env_2 split form: one byte per env of THIS launch for the policy kernel's sort modes, or nullptr (fused TCMLP kernel)
  // FUSE (Env_3 rollout kernel): the actor-critic packed by pack_fused_kernel and where the NEXT step's action goes
  const uint4* fused_w;     // kFwWords words (device memory, 16-byte aligned) or nullptr
  long long* next_actions;
  float* next_logp;
  float* next_value;
  unsigned draw_key0, draw_key1, draw_t;
  const unsigned* draw_t_dev;   // nullable: added to draw_t (a graph-replayed rollout bumps it between replays)
  int deterministic;
  float* terminal_obs;
  double* episode_return;
  int* episode_length;
  double* stats;
  // replay (REPLAY mode)
  const double* noise_u;
  const double* redis_u;
  long long redis_len;
  const uint32_t* input_counts;
  const uint8_t* press_choice;
  const uint8_t* sort_mode_in;
};

// stats slots accumulated by the step kernel (include/msort.h msort_info_out_t.stats)
enum { ST_EPISODES = 0, ST_RETURN = 1, ST_LENGTH = 2, ST_STEPS = 3, ST_REWARD = 4, ST_OVERFLOW = 5,
       ST_BALES = 6, ST_INVALID = 7, ST_CLAMPED = 8, ST_UNDERRUN = 9, ST_COUNT = 10 };

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Env_2's embedded policy travels as a kernel parameter (6.3 KB of the 32 KB parameter space): the
// weights then sit in the constant bank and every FFMA of the fully unrolled MLP takes its weight
// as a constant operand — no load instruction at all.
struct alignas(16) PolicyW { float w[(MSORT_POLICY_WEIGHTS + 3) / 4 * 4]; };   // paired layout (pack_policy_pairs), padded to whole LDCU.128s
struct NoPolicy {};
template <int KIND> using PolicyParam = typename std::conditional<KIND == MSORT_ENV_PRESS, PolicyW, NoPolicy>::type;

// ---------------------------------------------------------------- Env_2's embedded policy on the tensor cores (TCMLP)
// ref: sort_agent.predict(sort_obs, deterministic=True) env_2_press.py:106-109 — SB3 MlpPolicy 13 -> 32 -> 32 -> 2, tanh
// (training.py:115), argmax.  The CTA's tile of 128 envs is one UMMA M = 128 tile (TMEM lane = thread = env), so the
// three layers run as tcgen05.mma kind::f16 with fp32 accumulation in 32 TMEM columns.  To keep the fp32 reference's
// argmax (parity contract: identical except where |logit0 - logit1| < 1e-5) every operand is an fp16 SPLIT:
//   activation x = x_hi + x_lo (umma::split2, 22 significand bits), weight w = w_1 + w_2 + w_3 (host, 33 bits)
//   x*w ~ x_hi*w_1 + x_hi*w_2 + x_lo*w_1 + x_hi*w_3 + x_lo*w_2         (five MMAs per K step; each product exact,
//   fp32 accumulation, smallest terms first), error <= ~2^-23 relative + 2^-25 absolute per product.
// tanh is evaluated as 1 - 2r with r = 1 / (1 + 2^(c z)), c = 2 log2(e): the host folds c into this layer's weights
// and the affine map h = 1 - 2r into the NEXT layer's weights and bias (pack_policy_tc), so an activation costs
// MUFU.EX2 + FADD + MUFU.RCP and r (in (0, 1]) is what is split and stored as the next A operand.  Layer 1's bias rides
// in the three padding columns of the 13-wide observation (A = 1.0, B = the bias's three split terms); layer 2 adds
// its bias in the epilogue.  Layers 1 and 2 (13 -> 32 -> 32: 1 440 of the 1 504 MACs) are the MMAs; the 32 -> 2 output
// layer is evaluated in fp32 (FFMA2) on the layer-2 activations while they are still in registers, which saves the
// third operand store and the third MMA round trip.  Measured logit error: tests/test_tc_mlp_gpu.py.
// Env_2's observation row (16 floats) from the final state, written as four STS.128
__device__ __forceinline__ void put_press_row(const DevConfig& c, const Env& s, float* __restrict__ row) {
  float o[16];
  press_obs(c, s, o);
#pragma unroll
  for (int q = 0; q < 4; ++q) reinterpret_cast<float4*>(row)[q] = make_float4(o[4 * q], o[4 * q + 1], o[4 * q + 2], o[4 * q + 3]);
}

struct TcMlp {
  uint4* a;            // A operand: hi chunks [4][128] then lo chunks [4][128] (16 KB; the first 8 KB alias the obs tile)
  const uint32_t* w;   // packed weights in shared memory (kTcWords)
  uint64_t* bar;       // MMA completion
  uint32_t tmem;       // base address of the 32 accumulator columns ...
  const uint32_t* tmem_slot = nullptr;   // ... or where the allocating warp left it (read after the first CTA barrier of the MLP)
  uint64_t* wbar = nullptr;              // the weights' TMA barrier, awaited (phase 0) behind that barrier when given
};

__device__ __forceinline__ float tc_sigmoid2(float zp) {   // 1 / (1 + 2^zp); +inf -> 0, -inf -> 1
  float e, r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(zp));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(e + 1.0f));
  return r;
}

// two activations at once: the two "+ 1" share one packed add (FADD2)
__device__ __forceinline__ void tc_sigmoid2x2(float z0, float z1, float& r0, float& r1) {
  float e0, e1, d0, d1;
#if MSORT_EXP_SIGMOID == 1      // timing experiment: no MUFU at all (wrong values)
  umma::fadd2(z0, z1, 0.5f, 0.5f, r0, r1);
  return;
#elif MSORT_EXP_SIGMOID == 2    // timing experiment: EX2 only, the reciprocal replaced by one packed add (wrong values)
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e0) : "f"(z0));
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e1) : "f"(z1));
  umma::fadd2(e0, e1, 1.0f, 1.0f, r0, r1);
  return;
#endif
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e0) : "f"(z0));
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e1) : "f"(z1));
  umma::fadd2(e0, e1, 1.0f, 1.0f, d0, d1);
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r0) : "f"(d0));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r1) : "f"(d1));
}

// one thread: all split products of one layer, smallest terms first, then commit
template <int KSTEPS, int N, int NTERMS>
__device__ __forceinline__ void tc_issue_layer(const TcMlp& m, uint32_t tmem, int b_off_halves) {
  const uint32_t a0 = smem_u32(m.a), b0 = smem_u32(m.w) + 2u * (uint32_t)b_off_halves;
  constexpr uint32_t kTermBytes = (uint32_t)(KSTEPS * 16 * N * 2), kLoBytes = 4u * kTile * 16u;
  // the descriptors of one layer differ only in the 14-bit start-address field (bytes >> 4; shared memory is < 256 KB, so
  // base + offset never carries out of the field): two base words, then one immediate add per operand and MMA
  const uint64_t abase = umma::smem_desc(a0, kTile * 16u, 128u), bbase = umma::smem_desc(b0, N * 16u, 128u);
  uint32_t acc = 0u;
  auto mm = [&](int aterm, int bterm, int s) {
    const uint64_t ad = abase + (uint64_t)(((aterm ? kLoBytes : 0u) + (uint32_t)(2 * s) * kTile * 16u) >> 4);
    const uint64_t bd = bbase + (uint64_t)(((uint32_t)bterm * kTermBytes + (uint32_t)(2 * s) * N * 16u) >> 4);
    umma::mma_f16(tmem, ad, bd, umma::idesc_f16(N), acc);
    acc = 1u;
  };
#ifndef MSORT_TC_TERMS
#define MSORT_TC_TERMS 5   // 5 = all split products (product build); 3 / 1 = timing experiments only (lose accuracy)
#endif
  if (NTERMS == 3 && MSORT_TC_TERMS >= 5) {
#pragma unroll
    for (int s = 0; s < KSTEPS; ++s) { mm(1, 1, s); mm(0, 2, s); }
  }
  if (MSORT_TC_TERMS >= 3) {
#pragma unroll
    for (int s = 0; s < KSTEPS; ++s) { mm(1, 0, s); mm(0, 1, s); }
  }
#pragma unroll
  for (int s = 0; s < KSTEPS; ++s) mm(0, 0, s);
  umma::commit(m.bar);
}

// hidden-layer epilogue of this thread's env: 32 accumulators -> (+bias) -> r -> fp16 split -> the next A operand
__device__ __forceinline__ void tc_hidden_epilogue(const TcMlp& m, uint32_t tlane, int tid) {   // tlane: this thread's TMEM lane, column 0
#if MSORT_TC_ROLL
#pragma unroll 1
#else
#pragma unroll
#endif
  for (int half = 0; half < 2; ++half) {
    float v[16];
    umma::tmem_ld16(tlane + 16u * half, v);
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      uint32_t h[4], l[4];
#pragma unroll
      for (int p = 0; p < 4; ++p) {
        float r0, r1;
        tc_sigmoid2x2(v[8 * q + 2 * p], v[8 * q + 2 * p + 1], r0, r1);
        umma::split2(r0, r1, h[p], l[p]);
      }
      m.a[(2 * half + q) * kTile + tid] = make_uint4(h[0], h[1], h[2], h[3]);
      m.a[(4 + 2 * half + q) * kTile + tid] = make_uint4(l[0], l[1], l[2], l[3]);
    }
  }
}

// All 128 threads of a tile call this together (it contains CTA barriers); `so` = this env's 13-wide sort
// observation; l0 / l1 = the two logits.  Two MMA round trips (layers 1 and 2); the 32 -> 2 output layer is 32 packed
// fp32 FMAs on the activations that are still in registers (no split, no operand store, no third round trip).
#ifndef MSORT_EXP_TCPROF
#define MSORT_EXP_TCPROF 0
#endif
#if MSORT_EXP_TCPROF
__device__ long long g_tcprof[16];
#define TCPROF(k) do { if (tid == 0 && blockIdx.x == 0) g_tcprof[k] = clock64(); } while (0)
#else
#define TCPROF(k) do { } while (0)
#endif
__device__ __forceinline__ void tc_mlp_logits(const TcMlp& m, const float (&so)[13], int tid, uint32_t& phase, float& l0, float& l1) {
  uint32_t tmem = m.tmem, tlane = tmem + ((uint32_t)(tid & ~31) << 16);
  TCPROF(0);
  {   // layer-1 operand: obs columns 0..12, then 1, 1, 1 (bias terms); the lo part of an exact 1.0 is 0
    uint32_t h[8], l[8];
#pragma unroll
    for (int p = 0; p < 6; ++p) umma::split2(so[2 * p], so[2 * p + 1], h[p], l[p]);
    umma::split2(so[12], 1.0f, h[6], l[6]);
    umma::split2(1.0f, 1.0f, h[7], l[7]);
    m.a[tid] = make_uint4(h[0], h[1], h[2], h[3]);
    m.a[kTile + tid] = make_uint4(h[4], h[5], h[6], h[7]);
    m.a[4 * kTile + tid] = make_uint4(l[0], l[1], l[2], l[3]);
    m.a[5 * kTile + tid] = make_uint4(l[4], l[5], l[6], l[7]);
  }
#pragma unroll
  for (int layer = 0; layer < 2; ++layer) {
    TCPROF(1 + 5 * layer);
    umma::fence_async_proxy();       // this thread's operand writes -> visible to the MMA's async-proxy reads
    umma::fence_before_sync();       // ... and its TMEM reads are done before the next MMA overwrites the columns
    __syncthreads();
    TCPROF(2 + 5 * layer);
    if (layer == 0) {                // a CTA that set itself up just before (one CTA per tile): its barriers, TMEM columns and weights
      if (m.tmem_slot) { umma::fence_after_sync(); tmem = *m.tmem_slot; tlane = tmem + ((uint32_t)(tid & ~31) << 16); }
      if (m.wbar) mbar_wait(m.wbar, 0u);
    }
    if (tid == 0) {
      umma::fence_after_sync();
      if (layer == 0) tc_issue_layer<1, 32, 3>(m, tmem, kTcB1);
      else tc_issue_layer<2, 32, 3>(m, tmem, kTcB2);
    }
    TCPROF(3 + 5 * layer);
    mbar_wait(m.bar, phase); phase ^= 1u;
    umma::fence_after_sync();
    TCPROF(4 + 5 * layer);
    if (layer == 0) tc_hidden_epilogue(m, tlane, tid);
    TCPROF(5 + 5 * layer);
  }
  // layer-2 epilogue fused with the output layer: z -> +bias -> r2 (registers) -> logits += r2 * (-2 W3) (FFMA2)
  const float* const fw = reinterpret_cast<const float*>(m.w);
  const float4* const b2 = reinterpret_cast<const float4*>(fw + kTcBias2);
  const ulonglong2* const w3 = reinterpret_cast<const ulonglong2*>(fw + kTcW3);   // two (logit0, logit1) weight pairs per 16 bytes
  unsigned long long acc = *reinterpret_cast<const unsigned long long*>(fw + kTcBias3);
#if MSORT_TC_ROLL
#pragma unroll 1
#else
#pragma unroll
#endif
  for (int half = 0; half < 2; ++half) {
    float v[16];
    umma::tmem_ld16(tlane + 16u * half, v);
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const float4 bb = b2[4 * half + q];
      float z0, z1, z2, z3;
      umma::fadd2(v[4 * q], v[4 * q + 1], bb.x, bb.y, z0, z1);
      umma::fadd2(v[4 * q + 2], v[4 * q + 3], bb.z, bb.w, z2, z3);
      const ulonglong2 wa = w3[8 * half + 2 * q], wb = w3[8 * half + 2 * q + 1];
      float r0, r1, r2, r3;
      tc_sigmoid2x2(z0, z1, r0, r1);
      tc_sigmoid2x2(z2, z3, r2, r3);
      acc = ffma2_bcast(wa.x, r0, acc);
      acc = ffma2_bcast(wa.y, r1, acc);
      acc = ffma2_bcast(wb.x, r2, acc);
      acc = ffma2_bcast(wb.y, r3, acc);
    }
  }
  umma::fence_before_sync();         // TMEM reads done: ordered before the next tile's first MMA by the barriers in between
  f2unpack(acc, l0, l1);
  TCPROF(11);
}

// the sort mode: argmax of the two logits, ties -> 0 like np.argmax (sort_agent.predict(..., deterministic=True))
__device__ __forceinline__ int tc_mlp_mode(const TcMlp& m, const float (&so)[13], int tid, uint32_t& phase) {
  float l0, l1;
  tc_mlp_logits(m, so, tid, phase, l0, l1);
  return l1 > l0 ? 1 : 0;
}

// Diagnostics (msort_debug_policy_logits): the tensor-core policy alone on caller-given sort observations, logits out
// (unscaled), so tests can measure its error against an fp32 / fp64 evaluation of the same network.
__global__ void __launch_bounds__(kTile)
tc_logits_kernel(const float* __restrict__ obs13, const uint4* __restrict__ tcw, long long n, float* __restrict__ logits) {
  __shared__ __align__(128) uint4 s_a[8 * kTile];
  __shared__ __align__(128) uint32_t s_w[kTcWords];
  __shared__ __align__(8) uint64_t s_bar;
  __shared__ uint32_t s_tm;
  const int tid = threadIdx.x;
  if (tid == 0) { mbar_init(&s_bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  if (tid < 32) umma::tmem_alloc(&s_tm, 32u);
  for (int e = tid; e < kTcWords / 4; e += kTile) reinterpret_cast<uint4*>(s_w)[e] = tcw[e];
  umma::fence_async_proxy();
  umma::fence_before_sync();
  __syncthreads();
  umma::fence_after_sync();
  const TcMlp m{s_a, s_w, &s_bar, s_tm};
  uint32_t phase = 0;
  const long long ntiles = (n + kTile - 1) / kTile;
  for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const long long i = tile * kTile + tid;
    float so[13];
#pragma unroll
    for (int k = 0; k < 13; ++k) so[k] = i < n ? obs13[i * 13 + k] : 0.f;
    float l0, l1;
    tc_mlp_logits(m, so, tid, phase, l0, l1);
    if (i < n) { logits[2 * i] = l0; logits[2 * i + 1] = l1; }
    __syncthreads();   // every thread is done with the operand buffer before the next tile overwrites it
  }
  umma::fence_before_sync();
  __syncthreads();
  if (tid < 32) umma::tmem_dealloc(m.tmem, 32u);
#if MSORT_EXP_TCPROF
  if (blockIdx.x == 0 && tid == 0) for (int k = 0; k < 12; ++k) logits[k] = (float)(g_tcprof[k] - g_tcprof[0]);
#endif
}

// ---------------------------------------------------------------- FUSE: Env_3's step + the rollout policy in one kernel
// ref: the MaskablePPO rollout loop (training.py:118-143 -> sb3 collect_rollouts): per env-step the policy forward on the
// observation the previous step() returned, a masked categorical draw, then step().  The HOT Env_3 kernel with FUSE runs
// the policy of the NEXT step on the observation / mask tile it has just built in shared memory — same 128-env tile, same
// threads (thread = env = TMEM lane) — so the rollout is ONE launch per env-step and the policy never re-reads the 138 B of
// observation + mask per env from HBM.  The actor-critic (two tanh towers 29 -> 32 -> 32 -> {22 | 1}) runs on tcgen05:
//   layer 1  [128 x 32] x [32 x 64]   both towers side by side; K 29 = 1.0 carries the bias            -> TMEM columns 0..63
//   layer 2  [128 x 32] x [32 x 32]   twice (policy tower -> columns 0..31, value tower -> 32..63), bias in the epilogue
//   layer 3  [128 x 32] x [32 x 32]   the 22 logits; the value head (32 MACs) stays in fp32 registers
// fp16 operands in the canonical no-swizzle K-major layout, fp32 accumulation.  The 16 KB operand buffer aliases the
// observation tile (its bulk store has been read out by then), so a CTA needs 30.4 KB: 7 CTAs (28 warps) per SM.
constexpr int kFwB1 = 0;                        // fp16 elements: layer 1 [K 32][N 64]
constexpr int kFwB2p = kFwB1 + 32 * 64;         // layer 2, policy tower [32][32]
constexpr int kFwB2v = kFwB2p + 32 * 32;        // layer 2, value tower [32][32]
constexpr int kFwB3 = kFwB2v + 32 * 32;         // layer 3 [32][32] (rows 22..31 zero)
constexpr int kFwHalves = kFwB3 + 32 * 32;      // 5120 fp16 = 10 240 B
constexpr int kFwBias2 = kFwHalves / 2;         // 32-bit words from here: layer-2 bias, policy | value (64 floats)
constexpr int kFwW3v = kFwBias2 + 64;           // value head weights (32 floats)
constexpr int kFwBias3 = kFwW3v + 32;           // 22 logit biases, ..., [31] = value head bias (32 floats)
constexpr int kFwWords = kFwBias3 + 32;         // 2688 words = 10 752 B
static_assert(kFwWords == MSORT_ROLLOUT_WEIGHTS && (kFwWords * 4) % 16 == 0, "packed rollout policy size");

// flat fp32 parameters (msort_ppo_* order: pi W1 b1 W2 b2 W3 b3 | vf W1 b1 W2 b2 W3 b3, torch Linear layout) -> kFwWords
__global__ void __launch_bounds__(256)
pack_fused_kernel(const float* __restrict__ p, uint32_t* __restrict__ out) {
  constexpr int D = 29, A = 22, H = 32;
  constexpr int pi_w1 = 0, pi_b1 = pi_w1 + H * D, pi_w2 = pi_b1 + H, pi_b2 = pi_w2 + H * H, pi_w3 = pi_b2 + H, pi_b3 = pi_w3 + A * H;
  constexpr int vf_w1 = pi_b3 + A, vf_b1 = vf_w1 + H * D, vf_w2 = vf_b1 + H, vf_b2 = vf_w2 + H * H, vf_w3 = vf_b2 + H, vf_b3 = vf_w3 + H;
  const int e = blockIdx.x * 256 + threadIdx.x;
  if (e < kFwHalves / 2) {                       // two fp16 weights per word; canonical order [k/8][n][k%8]
    float v[2];
#pragma unroll
    for (int hh = 0; hh < 2; ++hh) {
      int x = 2 * e + hh;
      float w = 0.f;
      if (x < kFwB2p) {                          // layer 1, N = 64
        const int j = x & 7, n = (x >> 3) & 63, k = 8 * (x >> 9) + j;
        const int w1 = n < 32 ? pi_w1 : vf_w1, b1 = n < 32 ? pi_b1 : vf_b1, r = n & 31;
        w = k < D ? p[w1 + r * D + k] : (k == D ? p[b1 + r] : 0.f);
      } else {                                   // the three [32][32] tiles
        const int tile = (x - kFwB2p) >> 10; x = (x - kFwB2p) & 1023;
        const int j = x & 7, n = (x >> 3) & 31, k = 8 * (x >> 8) + j;
        w = tile == 0 ? p[pi_w2 + n * H + k] : (tile == 1 ? p[vf_w2 + n * H + k] : (n < A ? p[pi_w3 + n * H + k] : 0.f));
      }
      v[hh] = w;
    }
    const __half2 h = __floats2half2_rn(v[0], v[1]);
    out[e] = *reinterpret_cast<const uint32_t*>(&h);
  } else if (e < kFwWords) {
    const int f = e - kFwBias2;
    float w = 0.f;
    if (f < 64) w = f < 32 ? p[pi_b2 + f] : p[vf_b2 + f - 32];
    else if (f < 96) w = p[vf_w3 + f - 64];
    else { const int a = f - 96; w = a < A ? p[pi_b3 + a] : (a == 31 ? p[vf_b3] : 0.f); }
    out[e] = __float_as_uint(w);
  }
}

__device__ __forceinline__ uint32_t pack_h2(float a, float b) {
  const __half2 h = __floats2half2_rn(a, b);
  return *reinterpret_cast<const uint32_t*>(&h);
}
__device__ __forceinline__ float tanh_mufu(float x) {   // MUFU.TANH, |error| <~ 5e-4: the size of the operands' fp16 rounding
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// one thread: the K-steps of one fused-policy layer (A chunks from `a_chunk`, B tile at `b_halves`, D columns from `dcol`)
template <int N>
__device__ __forceinline__ void fused_issue(uint32_t a_saddr, uint32_t w_saddr, int a_chunk, int b_halves, uint32_t tmem, int dcol) {
  const uint64_t abase = umma::smem_desc(a_saddr + (uint32_t)a_chunk * kTile * 16u, kTile * 16u, 128u);
  const uint64_t bbase = umma::smem_desc(w_saddr + 2u * (uint32_t)b_halves, N * 16u, 128u);
#pragma unroll
  for (int s = 0; s < 2; ++s)       // K = 32 = two instructions of K 16 (two 16-byte chunks each)
    umma::mma_f16(tmem + (uint32_t)dcol, abase + (uint64_t)(((uint32_t)(2 * s) * kTile * 16u) >> 4),
                  bbase + (uint64_t)(((uint32_t)(2 * s) * N * 16u) >> 4), umma::idesc_f16(N), s > 0 ? 1u : 0u);
}

// The policy half of the rollout for one tile of 128 envs (thread = env = TMEM lane); shared by the fused step kernel (FUSE)
// and the stand-alone rollout_policy_kernel.  On entry: every thread holds its layer-1 operand words `a1` (fp16 pairs: K 0..28 =
// obs, K 29 = 1.0, K 30..31 = 0) and its 22-bit action mask `m22`; the 16 KB operand buffer `tile_raw` is free (a CTA barrier
// has passed since its last reader); the weights' TMA completes on `wbar` (phase 0); `tmem_slot` holds the base of 64 TMEM
// columns (allocated before a CTA barrier).  Contains CTA barriers: all 128 threads call it.
struct PolicyOut {
  long long* actions; float* logp; float* value;       // indexed by the tile-local row base the caller applied
  unsigned key0, key1, t; const unsigned* t_dev; int deterministic;
};
struct PolicyPhase { uint32_t mma, w; };      // parities of the MMA barrier / "weights seen" flag across the tiles of a persistent CTA
template <class AfterL3>
__device__ __forceinline__ void rollout_policy_tile(unsigned char* tile_raw, const uint32_t* wsm, uint64_t* wbar, uint64_t* mmabar,
                                                    uint32_t tmem, const uint32_t (&a1)[16], uint32_t m22, bool live,
                                                    long long i, long long gid, int tid, const PolicyOut& p, PolicyPhase& ph,
                                                    AfterL3 after_l3) {
  constexpr int A = 22;
    uint4* const A4 = reinterpret_cast<uint4*>(tile_raw);   // chunk kc (8 K values) of row r at (kc * 128 + r) * 16 B
#pragma unroll
    for (int kc = 0; kc < 4; ++kc) A4[kc * kTile + tid] = make_uint4(a1[4 * kc], a1[4 * kc + 1], a1[4 * kc + 2], a1[4 * kc + 3]);
    if (!ph.w) { mbar_wait(wbar, 0u); ph.w = 1u; }   // the weights (their fp32 part is read with plain loads below)
    const uint32_t a_saddr = smem_u32(tile_raw), w_saddr = smem_u32(wsm);
    const float* const fw = reinterpret_cast<const float*>(wsm);
    umma::fence_async_proxy();
    __syncthreads();
    umma::fence_after_sync();
    const uint32_t tlane = tmem + ((uint32_t)(tid & ~31) << 16);
    // ---- layer 1: both towers, N = 64
    if (tid == 0) { fused_issue<64>(a_saddr, w_saddr, 0, kFwB1, tmem, 0); umma::commit(mmabar); }
    mbar_wait(mmabar, ph.mma); ph.mma ^= 1u;
    umma::fence_after_sync();
    // (the epilogue loops stay rolled on purpose: the kernel is ~2 400 straight-line instructions per warp and seven CTAs
    //  sit at different places of it — instruction fetch is a measured stall here, `no_instruction` in profiles/ncu_r02_fused.md)
#pragma unroll 1
    for (int q = 0; q < 4; ++q) {             // 64 hidden units -> tanh -> chunks 0..3 (policy tower), 4..7 (value tower)
      float v[16];
      umma::tmem_ld16(tlane + 16u * q, v);
      uint32_t h[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) h[j] = pack_h2(tanh_mufu(v[2 * j]), tanh_mufu(v[2 * j + 1]));
      A4[(2 * q) * kTile + tid] = make_uint4(h[0], h[1], h[2], h[3]);
      A4[(2 * q + 1) * kTile + tid] = make_uint4(h[4], h[5], h[6], h[7]);
    }
    umma::fence_async_proxy();
    umma::fence_before_sync();
    __syncthreads();
    // ---- layer 2: one [32 x 32] product per tower
    if (tid == 0) {
      umma::fence_after_sync();
      fused_issue<32>(a_saddr, w_saddr, 0, kFwB2p, tmem, 0);
      fused_issue<32>(a_saddr, w_saddr, 4, kFwB2v, tmem, 32);
      umma::commit(mmabar);
    }
    mbar_wait(mmabar, ph.mma); ph.mma ^= 1u;
    umma::fence_after_sync();
#pragma unroll 1
    for (int q = 0; q < 2; ++q) {             // policy tower: + bias (LDS.128 broadcasts, packed adds) -> tanh -> chunks 0..3
      float v[16];
      umma::tmem_ld16(tlane + 16u * q, v);
      uint32_t h[8];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float4 bb = *reinterpret_cast<const float4*>(fw + kFwBias2 + 16 * q + 4 * j);
        float z0, z1, z2, z3;
        umma::fadd2(v[4 * j], v[4 * j + 1], bb.x, bb.y, z0, z1);
        umma::fadd2(v[4 * j + 2], v[4 * j + 3], bb.z, bb.w, z2, z3);
        h[2 * j] = pack_h2(tanh_mufu(z0), tanh_mufu(z1));
        h[2 * j + 1] = pack_h2(tanh_mufu(z2), tanh_mufu(z3));
      }
      A4[(2 * q) * kTile + tid] = make_uint4(h[0], h[1], h[2], h[3]);
      A4[(2 * q + 1) * kTile + tid] = make_uint4(h[4], h[5], h[6], h[7]);
    }
    float value = fw[kFwBias3 + 31];
#pragma unroll 1
    for (int q = 2; q < 4; ++q) {             // value tower: + bias -> tanh -> the value head's 32 MACs, in fp32
      float v[16];
      umma::tmem_ld16(tlane + 16u * q, v);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float4 bb = *reinterpret_cast<const float4*>(fw + kFwBias2 + 16 * q + 4 * j);
        const float4 ww = *reinterpret_cast<const float4*>(fw + kFwW3v + 16 * (q - 2) + 4 * j);
        float z0, z1, z2, z3;
        umma::fadd2(v[4 * j], v[4 * j + 1], bb.x, bb.y, z0, z1);
        umma::fadd2(v[4 * j + 2], v[4 * j + 3], bb.z, bb.w, z2, z3);
        value = fmaf(tanh_mufu(z0), ww.x, value);
        value = fmaf(tanh_mufu(z1), ww.y, value);
        value = fmaf(tanh_mufu(z2), ww.z, value);
        value = fmaf(tanh_mufu(z3), ww.w, value);
      }
    }
    umma::fence_async_proxy();
    umma::fence_before_sync();
    __syncthreads();
    // ---- layer 3: the 22 logits
    if (tid == 0) { umma::fence_after_sync(); fused_issue<32>(a_saddr, w_saddr, 0, kFwB3, tmem, 0); umma::commit(mmabar); }
    mbar_wait(mmabar, ph.mma); ph.mma ^= 1u;
    umma::fence_after_sync();
    after_l3();                               // the operand buffer has been consumed: the caller may refill it
    {
      float o[32];
      umma::tmem_ld16(tlane, *reinterpret_cast<float(*)[16]>(&o[0]));
      umma::tmem_ld16(tlane + 16u, *reinterpret_cast<float(*)[16]>(&o[16]));
      umma::fence_before_sync();              // TMEM reads done before the columns are given back
      if (live) {
        // masked log-softmax (sb3_contrib masks logits with -1e8 = probability 0 in fp32: here -inf; action a of Env_3 is
        // valid iff press action a % 11 is).  Everything in the exp2 domain: p_k ~ 2^((l_k - max) log2 e), 2^-inf = 0.
        float mx = -INFINITY;
#pragma unroll
        for (int k4 = 0; k4 < 6; ++k4) {
          const float4 bb = *reinterpret_cast<const float4*>(fw + kFwBias3 + 4 * k4);
          const float bk[4] = {bb.x, bb.y, bb.z, bb.w};
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int k = 4 * k4 + j;
            if (k < A) {
              o[k] = ((m22 >> k) & 1u) ? o[k] + bk[j] : -INFINITY;
              mx = fmaxf(mx, o[k]);
            }
          }
        }
        const float nmx = -mx * 1.4426950408889634f;
        float sum = 0.f;
#pragma unroll
        for (int k = 0; k < A; ++k) {
          asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(o[k]) : "f"(fmaf(o[k], 1.4426950408889634f, nmx)));   // unnormalised probability
          sum += o[k];
        }
        int na = 0;
        float pa = 0.f;
        if (p.deterministic) {
#pragma unroll
          for (int k = 0; k < A; ++k) if (o[k] > pa) { pa = o[k]; na = k; }
        } else {   // inverse-CDF draw with one Philox uniform keyed by (seed, draw index, global env id) — msort_policy_act's draw
          const unsigned long long g = (unsigned long long)gid;
          const unsigned dt = p.t + (p.t_dev ? *p.t_dev : 0u);
          const U4 r4 = philox4x32_10((uint32_t)g, (uint32_t)(g >> 32), 0xAC70u, dt, p.key0, p.key1);
          // uniform in (0, sum), kept strictly below sum (the cdf's last value is bit-identical to `sum`: same order of adds),
          // so the first k with cdf_k > u always exists and has p_k > 0
          const float u = fminf(((float)(r4.x >> 8) + 0.5f) * (1.0f / 16777216.0f) * sum, sum * 0.99999994f);
          float cdf = 0.f;
          bool found = false;
#pragma unroll
          for (int k = 0; k < A; ++k) {
            if (!found) { na = k; pa = o[k]; }
            cdf += o[k];
            found = found || cdf > u;
          }
        }
        p.actions[i] = na;
        p.logp[i] = (__log2f(pa) - __log2f(sum)) * 0.6931471805599453f;
        p.value[i] = value;
      }
    }
}

// The policy half alone (msort_rollout_policy): the rollout's first action, and the two-kernel form of the loop.  One CTA =
// one tile of 128 envs: the weights and the observation tile (14.5 KB, contiguous in the [N, 29] tensor) arrive by TMA bulk
// copies, the 22 mask bytes of each env by eleven 2-byte loads of its own thread (no mask tile: 26.6 KB per CTA = 8 CTAs per
// SM); each thread turns its row into the layer-1 operand words and its mask bytes into bits; the observation tile's space
// then becomes the operand buffer, exactly as in the fused kernel.
__global__ void __launch_bounds__(kTile, 8)
rollout_policy_kernel(const float* __restrict__ obs, const uint8_t* __restrict__ mask, const uint4* __restrict__ packed, long long n,
                      long long gid0, const __grid_constant__ PolicyOut po) {
  constexpr int D = 29, A = 22;
  __shared__ __align__(128) unsigned char s_tile_raw[8 * kTile * 16];
  __shared__ __align__(128) uint32_t s_w[kFwWords];
  __shared__ __align__(8) uint64_t s_wbar, s_mma, s_in;
  __shared__ uint32_t s_tmem;
  const int tid = threadIdx.x;
  float* const s_obs = reinterpret_cast<float*>(s_tile_raw);
  const long long ntiles = (n + kTile - 1) / kTile;
  // TMA needs 16-byte aligned global addresses: whole tiles of a 16-byte aligned tensor (a tile is a multiple of 16 bytes)
  const bool obs_aligned = (reinterpret_cast<uintptr_t>(obs) & 15u) == 0, mask_even = (reinterpret_cast<uintptr_t>(mask) & 1u) == 0;
  auto tile_by_tma = [&](long long tl) { return obs_aligned && (tl + 1) * kTile <= n; };
  auto fetch_obs = [&](long long tl) {      // one thread: the observation tile of `tl` into the (free) operand buffer
    if (tl < ntiles && tile_by_tma(tl)) {
      mbar_expect_tx(&s_in, (uint32_t)(kTile * D * 4));
      bulk_load(s_obs, obs + tl * kTile * D, kTile * D * 4u, &s_in);
    }
  };
  auto mask_bits = [&](long long tl) -> uint32_t {   // this thread's env of tile `tl`: its 22 mask bytes -> bits
    const long long r = tl * kTile + tid;
    uint32_t m = 0;
    if (tl < ntiles && r < n) {
      const uint8_t* mrow = mask + r * A;
      if (mask_even) {                       // rows are 22 bytes: 2-byte aligned
        const uint16_t* mr = reinterpret_cast<const uint16_t*>(mrow);
#pragma unroll
        for (int k = 0; k < A / 2; ++k) {
          const uint32_t w = mr[k];
          m |= ((w & 0xffu) ? 1u : 0u) << (2 * k) | ((w >> 8) ? 1u : 0u) << (2 * k + 1);
        }
      } else {
#pragma unroll
        for (int k = 0; k < A; ++k) m |= (mrow[k] ? 1u : 0u) << k;
      }
    }
    return m;
  };
  if (tid == 0) {
    mbar_init(&s_wbar, 1); mbar_init(&s_mma, 1); mbar_init(&s_in, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    fetch_obs(blockIdx.x);
    mbar_expect_tx(&s_wbar, kFwWords * 4u);
    bulk_load(s_w, packed, kFwWords * 4u, &s_wbar);
  }
  if (tid < 32) { __syncwarp(); umma::tmem_alloc(&s_tmem, 64u); umma::fence_before_sync(); }
  uint32_t m22 = mask_bits(blockIdx.x);
  __syncthreads();                            // barrier inits and the TMEM address are visible
  umma::fence_after_sync();
  const uint32_t tmem = s_tmem;
  PolicyPhase ph{0u, 0u};
  uint32_t in_phase = 0;
  // persistent: one wave of resident CTAs, each looping over tiles blockIdx.x, + gridDim.x, ...; the next tile's observations
  // land in the operand buffer (and its mask bits in a register) while the current tile's softmax / draw epilogue runs
  for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const long long row0 = tile * kTile, i = row0 + tid;
    const int rows = (int)min((long long)kTile, n - row0);
    const bool live = i < n;
    if (tile_by_tma(tile)) { mbar_wait(&s_in, in_phase); in_phase ^= 1u; }
    else {                                    // ragged last tile / unaligned tensor: plain loads (the buffer is free: barrier below)
      for (int e = tid; e < kTile * D; e += kTile) s_obs[e] = e < rows * D ? obs[row0 * D + e] : 0.f;
      __syncthreads();
    }
    uint32_t a1[16];
    {
      const float* x = &s_obs[tid * D];
#pragma unroll
      for (int q = 0; q < 14; ++q) a1[q] = pack_h2(x[2 * q], x[2 * q + 1]);
      a1[14] = pack_h2(x[28], 1.0f);
      a1[15] = 0u;
    }
    __syncthreads();                          // every row has been read: the tile's space becomes the operand buffer
    const long long next = tile + gridDim.x;
    uint32_t m22_next = 0;
    rollout_policy_tile(s_tile_raw, s_w, &s_wbar, &s_mma, tmem, a1, m22, live, i, gid0 + i, tid, po, ph,
                        [&]() {               // after the layer-3 MMAs: the operand buffer is free again
                          if (tid == 0) fetch_obs(next);
                          m22_next = mask_bits(next);
                        });
    m22 = m22_next;
    umma::fence_before_sync();
    __syncthreads();                          // all TMEM reads of this tile are done before the next tile's first MMA
  }
  if (tid < 32) umma::tmem_dealloc(tmem, 64u);
}

// ---------------------------------------------------------------- Env_2: the embedded policy as its own kernel (split form)
// ref: sort_agent.predict(self.get_sort_obs(), deterministic=True) env_2_press.py:106-109.  press_policy_kernel reads the 64-byte
// compact state of every env, rebuilds exactly the 13-wide sort observation the step kernel would show the policy (previous
// step's accuracies from the Philox counter, stages after the shift, container purities), evaluates the 13-32-32-2 policy on
// the tensor cores (tc_mlp_mode: the same function the fused TCMLP kernel calls, so the two forms agree bit for bit) and
// writes ONE mode byte per env; step_kernel<PRESS,...,HOT,EXTMODE> then steps the env with that mode.  Two small kernels
// instead of one 56 KB one: no spills, 8 CTAs per SM each, the hardware CTA scheduler for both.  Whole tiles only (the
// launcher gives a ragged tail to the FFMA2 kernel), FAST configurations only (HOT implies FAST).
__global__ void __launch_bounds__(kTile, 8)
press_policy_kernel(const __grid_constant__ DevConfig c, const uint4* __restrict__ state, const uint4* __restrict__ tcw,
                    uint8_t* __restrict__ modes) {
  __shared__ __align__(128) uint4 s_a[8 * kTile];
  __shared__ __align__(128) uint32_t s_w[kTcWords];
  __shared__ __align__(8) uint64_t s_bar, s_wbar;
  __shared__ uint32_t s_tm;
  const int tid = threadIdx.x;
  const long long i = (long long)blockIdx.x * kTile + tid;     // whole tiles: i < c.n
  if (tid == 0) {
    mbar_init(&s_bar, 1); mbar_init(&s_wbar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    mbar_expect_tx(&s_wbar, kTcWords * 4u);
    bulk_load(s_w, tcw, kTcWords * 4u, &s_wbar);
  }
  if (tid < 32) { __syncwarp(); umma::tmem_alloc(&s_tm, 32u); }
  umma::fence_before_sync();
  Env s;
  load_planes<LAYOUT_COMPACT>(state, c.n_pad, i, s);
  const unsigned long long gid = (unsigned long long)(c.gid0 + i);
  const uint32_t gid_lo = (uint32_t)gid, gid_hi = (uint32_t)(gid >> 32) & 0xffffu;
  // exactly the step kernel's preamble for Env_2 (FAST): accuracy_sorter of this step = accuracy_belt of the previous one
  double acc_a, acc_b;
  const int pm = s.mode;
  philox_accuracy2(c, gid_lo, gid_hi, s.episode, s.step - 1, pm, acc_a, acc_b);
  s.acc[0] = pm ? acc_a : 1.0; s.acc[1] = pm ? 1.0 : acc_a;
  s.acc[2] = pm ? acc_b : 1.0; s.acc[3] = pm ? 1.0 : acc_b;
  if (s.step == 0) {
#pragma unroll
    for (int m = 0; m < 4; ++m) s.acc[m] = c.base_acc[m];
  }
  s.sort4 = s.belt4; s.belt4 = s.in4;             // update_environment's shift (the new input does not enter the sort observation)
  float so[13];
  int kq[4];
  purity_ks(c, s, kq);
  sort_obs(c, s, kq, so);
  __syncthreads();                                // barrier inits and the TMEM address are visible
  umma::fence_after_sync();
  const TcMlp m{s_a, s_w, &s_bar, s_tm};
  mbar_wait(&s_wbar, 0u);
  uint32_t phase = 0;
  const int mode = tc_mlp_mode(m, so, tid, phase);
  modes[i] = (uint8_t)mode;
  umma::fence_before_sync();
  __syncthreads();
  if (tid < 32) umma::tmem_dealloc(s_tm, 32u);
}

// FAST (PHILOX only, chosen by the host when DevConfig::fast holds): boosted accuracies are exactly 1.0,
// unboosted ones need no clip, and the redistribution classes fit one register of packed bytes.
// Results are identical to the generic instantiation; only the instruction count differs.
// HOT (FAST + compact layout only): the per-call switches are the training configuration — action masking
// and auto-reset on, no overflow check, mask output wanted, no per-step info arrays, at most twelve
// redistribution draws per station (DevConfig::one_block); SMALL = levels <= 8192 (DevConfig::small_lv, e.g.
// max_steps 50), otherwise < 2^16 (compact layout, e.g. the reference's 200-step episodes) — and are
// compiled in, which removes ~20 uniform branches (and the basic-block boundaries they put in the
// scheduler's way).  Chosen per launch by launch_step_kind.
// TCMLP (Env_2's persistent HOT kernel only; full tiles only — the launcher gives a ragged tail to the plain HOT kernel):
// the embedded policy is evaluated on the tensor cores (tc_mlp_mode above) instead of per-thread FFMA2.
template <int KIND, int RNG, int LAYOUT, bool FAST, bool HOT = false, bool SMALL = HOT, bool TCMLP = false, bool FUSE = false, bool EXTMODE = false>
__global__ void __launch_bounds__(kTile, (KIND == MSORT_ENV_PRESS ? ((TCMLP || EXTMODE) ? MSORT_PRESS_TC_MIN_BLOCKS : MSORT_PRESS_MIN_BLOCKS)   // Env_2 (FFMA2 form) keeps 32 MLP activations in registers
                                           : FUSE ? MSORT_FUSE_MIN_BLOCKS
                                           : (KIND == MSORT_ENV_MONO && HOT) ? MSORT_HOT_MONO_MIN_BLOCKS
                                           : (KIND == MSORT_ENV_SORT && HOT) ? MSORT_HOT_SORT_MIN_BLOCKS : MSORT_STEP_MIN_BLOCKS) * (128 / kTile))
step_kernel(const __grid_constant__ DevConfig c, const __grid_constant__ StepArgs a,
            const __grid_constant__ PolicyParam<KIND> pw) {
  constexpr int D = Dims<KIND>::D, A = Dims<KIND>::A;
  static_assert(!TCMLP || !MSORT_HOT_PERSIST || (KIND == MSORT_ENV_PRESS && HOT && kTile == 128), "TCMLP specialises Env_2's persistent HOT kernel");   // (MSORT_HOT_PERSIST=0 experiment builds never launch it)
  static_assert(!FUSE || (KIND == MSORT_ENV_MONO && HOT && kTile == 128), "FUSE specialises Env_3's HOT kernel");
  // EXTMODE: Env_2's HOT kernel with the sort mode of every env given in a.sort_mode_in (written by press_policy_kernel just
  // before, on the same stream): no embedded policy in this kernel at all — one CTA per tile, 64 registers, 8 CTAs per SM
  static_assert(!EXTMODE || (KIND == MSORT_ENV_PRESS && HOT && !TCMLP), "EXTMODE specialises Env_2's HOT kernel");
  // obs tile; with TCMLP the 16 KB MMA A-operand buffer, whose first half the obs tile aliases (the operand is dead
  // once the last layer's MMAs are complete, long before the first obs entry of the tile is written)
  constexpr int kObsBytes = kTile * D * (int)sizeof(float);
  __shared__ __align__(128) unsigned char s_tile_raw[(TCMLP || FUSE) ? (8 * kTile * 16 > kObsBytes ? 8 * kTile * 16 : kObsBytes) : kObsBytes];
  float* const s_obs = reinterpret_cast<float*>(s_tile_raw);
  __shared__ __align__(16) uint8_t s_mask[kTile * A];
  __shared__ __align__(128) uint32_t s_tcw[TCMLP ? kTcWords : (FUSE ? kFwWords : 4)];   // TCMLP / FUSE: packed fp16 weight tiles + biases
  __shared__ __align__(8) uint64_t s_mma;                           // TCMLP / FUSE: MMA completion
  __shared__ __align__(8) uint64_t s_wbar;                          // FUSE: the weights have landed (TMA complete_tx)
  __shared__ uint32_t s_tmem;                                       // TCMLP / FUSE: TMEM base address
  __shared__ double s_accs[RNG == MSORT_RNG_REPLAY ? 4 : 1][RNG == MSORT_RNG_REPLAY ? kTile : 1];  // accuracy_sorter (REPLAY)
  __shared__ double s_stat[kTile / 32][ST_COUNT];   // per-warp partial sums (plain stores: no init, no atomics)

  static_assert(!HOT || (FAST && LAYOUT == LAYOUT_COMPACT), "HOT specialises the FAST compact kernel");
  const bool masking = HOT || (c.flags & MSORT_F_ACTION_MASKING);
  const bool auto_reset = HOT || (c.flags & MSORT_F_AUTO_RESET);
  const bool want_mask = HOT || a.mask != nullptr;
  const bool use_mlp = !EXTMODE && (TCMLP || (KIND == MSORT_ENV_PRESS && (c.flags & MSORT_F_SORT_POLICY_MLP) &&
                                                !(RNG == MSORT_RNG_REPLAY && a.sort_mode_in)));
  const int tid = threadIdx.x;

  // PERSIST (Env_2's HOT instantiation): the grid is one wave of resident CTAs, each looping over tiles
  // blockIdx.x, blockIdx.x + gridDim.x, ...  While a tile is computed the TMA engine already copies the next
  // tile's four state planes (and its actions) into shared memory — issued by one thread, completion counted
  // on an mbarrier — so no warp waits for DRAM at the top of a tile.  Measured: Env_2 (long, MLP-heavy tiles)
  // +7 %; Env_1 / Env_3 10-17 % SLOWER than one CTA per tile with the hardware scheduler refilling the SM
  // (second barrier per tile, ~50 more instructions per warp), so they keep that form plus the L2 prefetch.
  constexpr bool PERSIST = HOT && KIND == MSORT_ENV_PRESS && !EXTMODE && (TCMLP ? MSORT_TC_PERSIST != 0 : MSORT_HOT_PERSIST != 0);
  // TCMLP keeps shared memory at 27.1 KB per CTA (8 resident CTAs per SM): the staging buffer IS the lo half of the MMA
  // operand buffer (free from the completion of the layer-2 MMAs until the next tile's layer-1 operand is built — the
  // next tile is staged in exactly that window), and the actions come by plain coalesced loads.
  __shared__ __align__(16) uint4 s_in_own[PERSIST && !TCMLP ? 4 * kTile : 1];   // staged state planes C0..C3 of the next tile
  __shared__ __align__(16) long long s_act[PERSIST && !TCMLP ? kTile : 2];      // ... and its actions
  uint4* const s_in = TCMLP ? reinterpret_cast<uint4*>(s_tile_raw + 4 * kTile * 16) : s_in_own;
  __shared__ __align__(8) uint64_t s_full;                             // "staged tile has landed"
  const long long ntiles = (c.n + kTile - 1) / kTile;
  auto stage_tile = [&](long long t) {                                  // one thread: start the copies of tile t
    const bool whole = !TCMLP && (t + 1) * kTile <= c.n && a.act_tma;
    mbar_expect_tx(&s_full, 4u * kTile * 16u + (whole ? kTile * 8u : 0u));
#pragma unroll
    for (int p = 0; p < 4; ++p) bulk_load(&s_in[p * kTile], a.state + p * c.n_pad + t * kTile, kTile * 16u, &s_full);
    if (whole) bulk_load(s_act, a.actions + t * kTile, kTile * 8u, &s_full);
  };
  // The once-per-tile serial jobs go to three different warps so that no warp is the CTA's straggler at the
  // two barriers: lane 0 of warp 1 stages the next tile, warp 2 adds the statistics, lane 0 of warp 3 issues
  // (and later waits for) the bulk stores.  Without PERSIST everything stays with warp 0: the other warps exit.
  constexpr int kStageTid = 32 % kTile, kStatTid = PERSIST ? 64 % kTile : 0, kStoreTid = PERSIST ? 96 % kTile : 0;
  uint32_t phase = 0, mma_phase = 0;
  bool weights_seen = false;
  TcMlp tcm{reinterpret_cast<uint4*>(s_tile_raw), s_tcw, &s_mma, 0u};
  if (PERSIST) {
    if (tid == kStageTid) {
      mbar_init(&s_full, 1);
      if (TCMLP) mbar_init(&s_mma, 1);
      if (TCMLP) mbar_init(&s_wbar, 1);
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
      stage_tile(blockIdx.x);
      if (TCMLP) {   // once per resident CTA: the packed weights by one TMA bulk copy (lands while the first tile is staged)
        mbar_expect_tx(&s_wbar, kTcWords * 4u);
        bulk_load(s_tcw, a.policy_tc, kTcWords * 4u, &s_wbar);
      }
    }
    if (TCMLP) {   // ... and 32 TMEM columns (warp 0)
      if (tid < 32) umma::tmem_alloc(&s_tmem, 32u);
      umma::fence_before_sync();
    }
    __syncthreads();
    if (TCMLP) { umma::fence_after_sync(); tcm.tmem = s_tmem; }
  }
  if (TCMLP && !PERSIST) {   // one CTA per tile: barriers, the weights' TMA and 32 TMEM columns per CTA, behind the state loads
    if (tid == 0) {
      mbar_init(&s_mma, 1);
      mbar_init(&s_wbar, 1);
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
      mbar_expect_tx(&s_wbar, kTcWords * 4u);
      bulk_load(s_tcw, a.policy_tc, kTcWords * 4u, &s_wbar);
    }
    if (tid < 32) { __syncwarp(); umma::tmem_alloc(&s_tmem, 32u); }
    umma::fence_before_sync();
    __syncthreads();       // (taking the TMEM address and the weights behind the MLP's first barrier instead — TcMlp::tmem_slot /
    umma::fence_after_sync();   //  wbar — measured 3 % SLOWER: 107.3 vs 104.1 us per 1 048 576 envs, more spills at 64 registers)
    tcm.tmem = s_tmem;
  }
  if (FUSE) {   // the weights start their way into shared memory now and land behind the step; 64 TMEM columns
    if (tid == 0) {
      mbar_init(&s_wbar, 1);
      mbar_init(&s_mma, 1);
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
      mbar_expect_tx(&s_wbar, kFwWords * 4u);
      bulk_load(s_tcw, a.fused_w, kFwWords * 4u, &s_wbar);
    }
    if (tid < 32) { __syncwarp(); umma::tmem_alloc(&s_tmem, 64u); umma::fence_before_sync(); }
  }
  uint32_t mbits = 0;   // FUSE: this env's press-mask bits for the policy phase
  // TCMLP: the tile's actions come by plain loads issued one tile ahead (whole tiles only), the weights' TMA is awaited once
  long long act_ahead = 0;
  if (TCMLP && PERSIST) act_ahead = a.actions[(long long)blockIdx.x * kTile + tid];
  long long tile = blockIdx.x;
  do {   // one pass unless PERSIST
  const long long row0 = tile * kTile;
  const long long i = row0 + tid;
  const bool live = i < c.n;
  const int rows = (int)min((long long)kTile, c.n - row0);

  // per-thread statistics (reduced per warp at the end)
  uint32_t st_flags = 0;   // done | overflow<<8 | invalid<<16 | clamped<<24
  uint32_t st_bales = 0, st_len = 0, st_underrun = 0;
  double st_reward = 0.0, st_return = 0.0;

#if MSORT_PREFETCH_TILES > 0
  // Pull the state planes and actions of the tile that runs MSORT_PREFETCH_TILES CTAs later (about one wave of
  // resident CTAs) into L2, so that its initial loads — the largest single stall of the kernel — see L2
  // latency instead of DRAM latency.  64 state lines + 8 action lines of 128 B per tile.
  if (LAYOUT == LAYOUT_COMPACT && !PERSIST) {
    const long long pt = (long long)blockIdx.x + MSORT_PREFETCH_TILES;
    if (pt < (long long)gridDim.x && (tid < 64 || (tid < 72 && (pt + 1) * kTile <= c.n))) {   // actions are not padded: whole tiles only
      const char* p = tid < 64 ? reinterpret_cast<const char*>(a.state + (tid >> 4) * c.n_pad + pt * kTile) + (tid & 15) * 128
                               : reinterpret_cast<const char*>(a.actions + pt * kTile) + (tid - 64) * 128;
      asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
    }
  }
#endif
  Env s;
  long long act = 0;
  if (PERSIST) {
    if (TCMLP) act = act_ahead;
    mbar_wait(&s_full, phase); phase ^= 1u;
    if (live) {
      load_planes<LAYOUT>(s_in, kTile, tid, s);
      if (!TCMLP) act = (a.act_tma && rows == kTile) ? s_act[tid] : a.actions[i];
    }
    if (tid == kStoreTid) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // the previous tile's obs / mask have left shared memory
    __syncthreads();                                                                       // every thread has taken its env out of the staging buffer
    if (!TCMLP && tid == kStageTid && tile + gridDim.x < ntiles) stage_tile(tile + gridDim.x);   // (TCMLP: after the policy's last MMA)
  } else if (live) {
    load_planes<LAYOUT>(a.state, c.n_pad, i, s);
    act = a.actions[i];
  }
  if (live) {
    const unsigned long long gid = (unsigned long long)(c.gid0 + i);
    const uint32_t gid_lo = (uint32_t)gid, gid_hi = (uint32_t)(gid >> 32) & 0xffffu;

    const uint32_t ep = s.episode, stp = s.step;
    float* const orow = &s_obs[tid * D];          // this env's row of the dense obs tile
    float* const prow = orow + (KIND == MSORT_ENV_MONO ? 13 : 0);  // press part of the row
    if ((unsigned long long)act >= (unsigned long long)A) {   // outside Discrete(A): clamp and count (one unsigned compare)
      act = act < 0 ? 0 : A - 1; st_flags += 1u << 24;
    }

    // accuracy_sorter <- accuracy_belt (env_super.py:457).  REPLAY: stored in planes P6/P7 (and parked in
    // shared memory for the dynamically indexed loop).  PHILOX: recomputed from the previous step's
    // counter and mode — 64 B less state traffic per env-step than storing four float64.
    double acc_sorter[4];
    double acc_a = 0.0, acc_b = 0.0;              // FAST: accuracies of the two unboosted stations of the previous mode
    U4 hotA = {0, 0, 0, 0}, hotB = {0, 0, 0, 0};  // HOT: the two redistribution blocks
    const int pm = s.mode;                        // previous step's sensor mode
    if (RNG == MSORT_RNG_REPLAY) {
#pragma unroll
      for (int m = 0; m < 4; ++m) { acc_sorter[m] = s.acc[m]; s_accs[m][tid] = s.acc[m]; }
    } else if (FAST) {
      // stp == 0 (counter wraps): the sorting stage is empty right after a reset, so the values are never
      // multiplied by anything but 0; only Env_2's embedded policy observes them (set below)
      philox_accuracy2(c, gid_lo, gid_hi, ep, stp - 1, pm, acc_a, acc_b);   // unconditional: no branch between the Philox chains
      if (HOT) {   // one block per station is enough (c.one_block): start both redistribution blocks now, next to the accuracy chains
        hotA = env_draw(c, gid_lo, gid_hi, kBlkRedis + (pm ? 0u : 64u), ep, stp);
        hotB = env_draw(c, gid_lo, gid_hi, kBlkRedis + 128u, ep, stp);
      }
      if (KIND == MSORT_ENV_PRESS) {
        s.acc[0] = pm ? acc_a : 1.0; s.acc[1] = pm ? 1.0 : acc_a;
        s.acc[2] = pm ? acc_b : 1.0; s.acc[3] = pm ? 1.0 : acc_b;
      }
    } else if (stp == 0) {
#pragma unroll
      for (int m = 0; m < 4; ++m) acc_sorter[m] = c.base_acc[m];  // right after reset (env_super.py:395-396)
    } else {
      philox_accuracy(c, gid_lo, gid_hi, ep, stp - 1, s.mode, acc_sorter);
      if (KIND == MSORT_ENV_PRESS) {
#pragma unroll
        for (int m = 0; m < 4; ++m) s.acc[m] = acc_sorter[m];  // the embedded policy observes the old accuracies
      }
    }
    if (RNG != MSORT_RNG_REPLAY && KIND == MSORT_ENV_PRESS && stp == 0) {
#pragma unroll
      for (int m = 0; m < 4; ++m) s.acc[m] = c.base_acc[m];
    }

    // 1: material flow (update_environment env_super.py:440-442)
    s.sort4 = s.belt4; s.belt4 = s.in4;
    // 2: seasonal generator (input_generator.py:37-64); counts only
    if (RNG == MSORT_RNG_REPLAY && a.input_counts) {
      s.in4 = a.input_counts[i];
    } else {
      if (s.gcount >= c.spp) { s.gidx ^= 1; s.gcount = 0; }
      s.in4 = c.pat[s.gidx ^ s.gfirst];
      if (!FAST && c.pat_remainder > 0) {   // FAST implies no remainder
        U4 r4 = {0, 0, 0, 0};
        for (int k = 0; k < c.pat_remainder; ++k) {
          if ((k & 3) == 0) r4 = env_draw(c, gid_lo, gid_hi, kBlkInput + 0x100u * (uint32_t)(k >> 2), ep, stp);
          uint32_t x = (k & 3) == 0 ? r4.x : ((k & 3) == 1 ? r4.y : ((k & 3) == 2 ? r4.z : r4.w));
          s.in4 += 1u << (8 * (x & 3u));
        }
      }
      s.gcount += 1;
    }
    if (FAST) {
      // without an input remainder every stage holds pattern 1, pattern 2 or nothing (right after a
      // reset): those parts of the observation come from host-built tables (same float32 operations)
      // (a stage holding anything else can only come from an imported state: computed directly)
      if (KIND != MSORT_ENV_PRESS) {
        const bool b0 = s.belt4 == c.pat[0], b1 = s.belt4 == c.pat[1];
        if (b0 || b1 || s.belt4 == 0u) {
          const float* tb = c.obs_belt_tab[b0 ? 0 : (b1 ? 1 : 2)];   // one indexed constant load per entry
#pragma unroll
          for (int k = 0; k < 5; ++k) orow[k] = tb[k];
        } else {
          obs_belt(s, orow);
        }
      }
      if (KIND != MSORT_ENV_SORT && !TCMLP) {   // TCMLP: the obs tile is still the MMA operand buffer; written after the policy
        const bool s0 = s.sort4 == c.pat[0], s1 = s.sort4 == c.pat[1];
        if (s0 || s1 || s.sort4 == 0u) {
          const float* ts = c.obs_sort_tab[s0 ? 0 : (s1 ? 1 : 2)];
#pragma unroll
          for (int k = 0; k < 4; ++k) prow[10 + k] = ts[k];
        } else {
          obs_sorting(c, s, prow);
        }
      }
    } else {
      if (KIND != MSORT_ENV_PRESS) obs_belt(s, orow);      // final for this step: write now
      if (KIND != MSORT_ENV_SORT) obs_sorting(c, s, prow);
    }

    // 3: decode the action
    int mode = 0, pa = 0;
    bool skip_press = false, invalid = false;
    if (KIND == MSORT_ENV_SORT) {
      mode = (int)act;
    } else if (KIND == MSORT_ENV_MONO) {
      mode = (int)act >= 11 ? 1 : 0; pa = (int)act - 11 * mode;
      // validity is judged on the levels BEFORE this step's sort (env_monolith.py:132-138)
      if (!masking && !press_action_valid(c, s, pa)) { pa = 0; skip_press = true; invalid = true; }
    } else {
      pa = (int)act;
      if ((RNG == MSORT_RNG_REPLAY || EXTMODE) && a.sort_mode_in) {
        mode = a.sort_mode_in[i] & 1;
      } else if (use_mlp) {
        float so[13];
        int kq[4];
        purity_ks(c, s, kq);
        sort_obs(c, s, kq, so);
        if constexpr (TCMLP) {
          if (!weights_seen) { mbar_wait(&s_wbar, 0u); weights_seen = true; }
          mode = tc_mlp_mode(tcm, so, tid, mma_phase);   // every thread of the (full) tile is here: CTA barriers inside
          // this thread has seen the layer-2 MMAs complete: the operand buffer is dead, its lo half takes the next tile
          if (PERSIST && tile + gridDim.x < ntiles) {
            if (tid == kStageTid) stage_tile(tile + gridDim.x);
            act_ahead = a.actions[(tile + gridDim.x) * kTile + tid];
          }
        } else if constexpr (KIND == MSORT_ENV_PRESS) mode = mlp_sort_mode<HOT>(pw.w, so);
      } else {  // sorting_rules env_super.py:469-482: pA+pC > pB+pD on float64 proportions
        int ac = b4(s.belt4, 0) + b4(s.belt4, 2), bd = b4(s.belt4, 1) + b4(s.belt4, 3);
        if (ac != bd) mode = ac > bd ? 0 : 1;  // strict integer inequality survives the float64 rounding
        else {                                  // integer tie: the float64 sums decide, as in the reference
          int bt = ac + bd;
          double p[4];
#pragma unroll
          for (int m = 0; m < 4; ++m) p[m] = bt > 0 ? ddiv((double)b4(s.belt4, m), (double)bt) : 0.0;
          mode = dadd(p[0], p[2]) > dadd(p[1], p[3]) ? 0 : 1;
        }
      }
    }
    s.mode = mode;

    // 4: update_accuracy env_super.py:492-509.  PHILOX mode does not keep the new accuracies: they go
    //    straight into the observation and are recomputed next step (Env_2's obs has none at all).
    if (RNG == MSORT_RNG_REPLAY) {
      const double2* nz = reinterpret_cast<const double2*>(a.noise_u) + 2 * i;
      const double2 n0 = nz[0], n1 = nz[1];
      s.acc[0] = accuracy_of(c, 0, mode, n0.x); s.acc[1] = accuracy_of(c, 1, mode, n0.y);
      s.acc[2] = accuracy_of(c, 2, mode, n1.x); s.acc[3] = accuracy_of(c, 3, mode, n1.y);
      if (KIND != MSORT_ENV_PRESS) obs_acc(s.acc, orow);
    } else if (KIND != MSORT_ENV_PRESS) {
      if (FAST) {
        double na, nb;
        philox_accuracy2(c, gid_lo, gid_hi, ep, stp, mode, na, nb);
        const float fa = (float)na, fb = (float)nb;
        orow[5] = mode ? fa : 1.f; orow[6] = mode ? 1.f : fa;
        orow[7] = mode ? fb : 1.f; orow[8] = mode ? 1.f : fb;
      } else {
        double acc_new[4];
        philox_accuracy(c, gid_lo, gid_hi, ep, stp, mode, acc_new);
        obs_acc(acc_new, orow);
      }
    }

    // 5: sort_material env_super.py:511-609.
    uint32_t sorted_true4 = 0;   // correctly sorted units per station this step (bytes A..D); reported only
    if (RNG == MSORT_RNG_REPLAY) {
      // REPLAY: the reference's loop nest literally (per-class leftovers, numpy's float64 cdf search),
      // one recorded uniform per draw.
      uint32_t L = s.sort4, T4 = 0, F4 = 0;  // leftover / true / false counts, packed bytes
      int m = 0, rem = 0, tot = sum4(L);
      uint32_t sh = 0;
      while (true) {
        if (rem == 0 && m < 4) {
          const int t = (int)((L >> sh) & 0xffu);
          const int tv = __double2int_rn(dmul((double)t, s_accs[m][tid]));  // int(round(t*acc)) (:539)
          T4 += (uint32_t)tv << sh;
          L -= (uint32_t)tv << sh;  // leftover[m] = false_val (:546)
          tot -= tv;
          rem = t - tv;
          F4 += (uint32_t)rem << sh;
          ++m; sh += 8;
        }
        if (rem == 0) { if (m == 4) break; continue; }
        uint32_t j;
        if ((long long)s.cursor >= a.redis_len) {
          st_underrun += 1;
          j = 0; while (((L >> (8 * j)) & 0xffu) == 0) ++j;  // defined fallback; reported as an under-run
        } else {
          const double uu = a.redis_u[i * a.redis_len + s.cursor];
          s.cursor += 1;
          double cdf[3], cs = 0.0;  // numpy Generator.choice(4, p=): cumsum, /= last, searchsorted right
#pragma unroll
          for (int q = 0; q < 4; ++q) { cs = dadd(cs, ddiv((double)b4(L, q), (double)tot)); if (q < 3) cdf[q] = cs; }
          j = 0;
#pragma unroll
          for (int q = 0; q < 3; ++q) j += ddiv(cdf[q], cs) <= uu ? 1u : 0u;
        }
        L -= 1u << (8 * j); --rem; --tot;
      }
      s.e += sum4(L);                                            // :579,597
#pragma unroll
      for (int q = 0; q < 4; ++q) { s.tr[q] += b4(T4, q); s.fl[q] += b4(F4, q); }  // :600-602
      sorted_true4 = T4;
    } else if (FAST) {
      // PHILOX, FAST form of the block below (same draws, same class selection, same results).
      // The four classes live in one register as bytes  lump | X<<8 | L2<<16 | L3<<24  (every prefix
      // sum <= batch <= 127).  With P = C*0x01010101 (byte k = prefix sum through class k) and the
      // draw r replicated as 0x7f-r in every byte, bit 7 of byte k of P + (0x7f-r) is set iff r < P_k;
      // the hits are monotone in k, H = those bits as 0/1 bytes, and C + 255*H = C - (1 << 8j) for
      // the first hit j: one multiply-add removes the unit from the selected class.
      // With the boosted stations exact (accuracy 1.0: all true, no draws) exactly one of stations
      // 0/1 and at most station 2 draw, as selected by the previous mode pm.
      uint32_t C = s.sort4 & 0xffffff00u;
      int tot = sum4(s.sort4), rem;
      uint32_t T4, F4, blk0 = kBlkRedis;
      {                                   // station 0 (unboosted iff pm == 1)
        const int t = b4(s.sort4, 0);
        const int tv = pm ? __double2int_rn(dmul((double)t, acc_a)) : t;   // int(round(t*acc)) half-to-even (:539)
        rem = t - tv;
        T4 = (uint32_t)tv; F4 = (uint32_t)rem;
        tot -= tv; C += (uint32_t)rem;    // leftover[0] = false_val (:546)
      }
      if (!pm) {                          // station 0 had nothing to redistribute: station 1 draws in the same loop
        const int t = b4(C, 1);
        const int tv = __double2int_rn(dmul((double)t, acc_a));
        rem = t - tv;
        T4 |= (uint32_t)tv << 8; F4 |= (uint32_t)rem << 8;
        tot -= tv; C = (C & 0xffff00ffu) + (uint32_t)rem;
        blk0 = kBlkRedis + 64u;
      }
#define MSORT_PDRAW(xl, xh, k)                                                                         \
      {                                                                                                \
        const uint32_t r = draw64(xl, xh, (uint32_t)(tot - (k)));                                      \
        const uint32_t Tq = (C - r) * 0x01010101u + 0x7f7f7f7fu;   /* == C*0x01010101 + (0x7f - r) per byte */ \
        const uint32_t M = byte_sign_mask(Tq);                     /* byte k -> 0xff if its bit 7 is set: 255*H */ \
        if (rem > (k)) C += M;                                                                         \
      }
      // draw k of a block sees tot - k units (every earlier draw of an active lane removed one); an
      // inactive lane's products are never used
      if (HOT) {                          // at most twelve draws (host-proved): one straight, predicated block —
        U4 r4 = hotA;                     // unconditional: the unboosted station almost always mis-sorts something
        MSORT_PDRAW(r4.x, r4.y, 0); MSORT_PDRAW(r4.x, r4.y, 1); MSORT_PDRAW(r4.x, r4.y, 2);
        MSORT_PDRAW(r4.x, r4.y, 3); MSORT_PDRAW(r4.x, r4.y, 4); MSORT_PDRAW(r4.x, r4.y, 5);
        MSORT_PDRAW(r4.z, r4.w, 6); MSORT_PDRAW(r4.z, r4.w, 7); MSORT_PDRAW(r4.z, r4.w, 8);
        MSORT_PDRAW(r4.z, r4.w, 9); MSORT_PDRAW(r4.z, r4.w, 10); MSORT_PDRAW(r4.z, r4.w, 11);
        tot -= rem;
      } else
      for (uint32_t b = 0; rem > 0; ++b, tot -= min(rem, 12), rem -= 12) {
        U4 r4 = env_draw(c, gid_lo, gid_hi, blk0 + b, ep, stp);
        MSORT_PDRAW(r4.x, r4.y, 0); MSORT_PDRAW(r4.x, r4.y, 1); MSORT_PDRAW(r4.x, r4.y, 2);
        MSORT_PDRAW(r4.x, r4.y, 3); MSORT_PDRAW(r4.x, r4.y, 4); MSORT_PDRAW(r4.x, r4.y, 5);
        if (rem <= 6) { tot -= rem; break; }
        MSORT_PDRAW(r4.z, r4.w, 6); MSORT_PDRAW(r4.z, r4.w, 7); MSORT_PDRAW(r4.z, r4.w, 8);
        MSORT_PDRAW(r4.z, r4.w, 9); MSORT_PDRAW(r4.z, r4.w, 10); MSORT_PDRAW(r4.z, r4.w, 11);
      }
#undef MSORT_PDRAW
      if (pm) {                           // station 1 after station 0's draws: boosted, everything left is true
        T4 |= (C & 0xff00u);
        tot -= b4(C, 1); C &= 0xffff00ffu;
      }
      int lump;
      {                                   // station 2 (unboosted iff pm == 1): classes (lump, L3), tot == lump + L3
        const int t = b4(C, 2);
        const int tv = pm ? __double2int_rn(dmul((double)t, acc_b)) : t;
        rem = t - tv;
        T4 |= (uint32_t)tv << 16; F4 |= (uint32_t)rem << 16;
        tot -= tv; lump = b4(C, 0) + rem;
      }
#define MSORT_PDRAW2(xl, xh, k)                                                                        \
      {                                                                                                \
        const int r = (int)draw64(xl, xh, (uint32_t)(tot - (k)));                                      \
        if (rem > (k) && r < lump) lump -= 1;                                                          \
      }
      if (HOT) {
        if (rem > 0) {
          U4 r4 = hotB;
          MSORT_PDRAW2(r4.x, r4.y, 0); MSORT_PDRAW2(r4.x, r4.y, 1); MSORT_PDRAW2(r4.x, r4.y, 2);
          MSORT_PDRAW2(r4.x, r4.y, 3); MSORT_PDRAW2(r4.x, r4.y, 4); MSORT_PDRAW2(r4.x, r4.y, 5);
          MSORT_PDRAW2(r4.z, r4.w, 6); MSORT_PDRAW2(r4.z, r4.w, 7); MSORT_PDRAW2(r4.z, r4.w, 8);
          MSORT_PDRAW2(r4.z, r4.w, 9); MSORT_PDRAW2(r4.z, r4.w, 10); MSORT_PDRAW2(r4.z, r4.w, 11);
          tot -= rem;
        }
      } else
      for (uint32_t b = 0; rem > 0; ++b, tot -= min(rem, 12), rem -= 12) {
        U4 r4 = env_draw(c, gid_lo, gid_hi, kBlkRedis + 128u + b, ep, stp);
        MSORT_PDRAW2(r4.x, r4.y, 0); MSORT_PDRAW2(r4.x, r4.y, 1); MSORT_PDRAW2(r4.x, r4.y, 2);
        MSORT_PDRAW2(r4.x, r4.y, 3); MSORT_PDRAW2(r4.x, r4.y, 4); MSORT_PDRAW2(r4.x, r4.y, 5);
        if (rem <= 6) { tot -= rem; break; }
        MSORT_PDRAW2(r4.z, r4.w, 6); MSORT_PDRAW2(r4.z, r4.w, 7); MSORT_PDRAW2(r4.z, r4.w, 8);
        MSORT_PDRAW2(r4.z, r4.w, 9); MSORT_PDRAW2(r4.z, r4.w, 10); MSORT_PDRAW2(r4.z, r4.w, 11);
      }
#undef MSORT_PDRAW2
      {                                   // station 3 (unboosted iff pm == 0); its draws leave the pool sum unchanged
        const int t = tot - lump;
        const int tv = pm ? t : __double2int_rn(dmul((double)t, acc_b));
        T4 |= (uint32_t)tv << 24; F4 |= (uint32_t)(t - tv) << 24;
      }
      s.e += lump;                                               // :579,597
#pragma unroll
      for (int q = 0; q < 4; ++q) { s.tr[q] += b4(T4, q); s.fl[q] += b4(F4, q); }  // :600-602
      sorted_true4 = T4;
    } else {
      // PHILOX: the same random process in the form DESIGN.md §4 "Sorting" derives.
      // Removing a unit from a station that has already been processed (or from the current
      // station's own false units) only ever matters through the SUM of those leftovers — it feeds
      // later selection probabilities and finally container E — so they are kept as one `lump`;
      // only not-yet-processed stations are tracked individually.  The draws of the last station
      // cannot influence anything but that sum, which drops by exactly the number of draws, so
      // they are not simulated.  Draw k of station S uses Philox block kBlkRedis + 64*S + k/12: the block's
      // words form two 64-bit lanes (y:x) and (w:z); draws k%12 = 0..5 come from the first, 6..11 from the
      // second, each as r = hi64(lane*tot), lane = lo64(lane*tot) (draw64; the j-th draw of a lane is
      // biased by at most tot^j / 2^64 <= 6e-8).
      // Stations 0 and 1 share one warp loop: a lane whose station 0 has no false units starts
      // station 1 up front, so lanes drawing for station 0 and lanes drawing for station 1 run
      // side by side (with the default boosts every env has exactly one of the two to draw for).
      int X = b4(s.sort4, 1), L2 = b4(s.sort4, 2), L3 = b4(s.sort4, 3);
      int tot = b4(s.sort4, 0) + X + L2 + L3, lump = 0, rem;
      uint32_t T4 = 0, F4 = 0;            // per-station true / false counts (bytes), added to the containers below
      uint32_t blk0 = kBlkRedis;          // first Philox block of the station this lane draws for
      bool started1 = false;
      {                                   // station 0
        const int t = b4(s.sort4, 0);
        const int tv = __double2int_rn(dmul((double)t, acc_sorter[0]));  // int(round(t*acc)) half-to-even (:539)
        rem = t - tv;
        T4 = (uint32_t)tv; F4 = (uint32_t)rem;
        tot -= tv; lump += rem;           // leftover[0] = false_val (:546)
      }
#define MSORT_START_STATION1(cond)                                                                    \
      if (cond) {                                                                                      \
        const int t = X;                                                                               \
        const int tv = __double2int_rn(dmul((double)t, acc_sorter[1]));                                \
        rem = t - tv;                                                                                  \
        T4 |= (uint32_t)tv << 8; F4 |= (uint32_t)rem << 8;                                             \
        tot -= tv; lump += rem; X = 0;                                                                 \
        blk0 = kBlkRedis + 64u; started1 = true;                                                       \
      }
      // one draw over the classes (lump, X, L2, L3) in prefix order; `x` is replaced by the low product
#define MSORT_DRAW4(xl, xh)                                                                            \
      {                                                                                                \
        const bool act = rem > 0;                                                                      \
        const int r = (int)draw64(xl, xh, (uint32_t)tot);                                              \
        const int c0 = lump, c1 = c0 + X, c2 = c1 + L2;                                                \
        const bool h0 = r < c0, h1 = r < c1, h2 = r < c2;                                              \
        lump -= (act && h0) ? 1 : 0;                                                                   \
        X -= (act && !h0 && h1) ? 1 : 0;                                                               \
        L2 -= (act && !h1 && h2) ? 1 : 0;                                                              \
        L3 -= (act && !h2) ? 1 : 0;                                                                    \
        tot -= act ? 1 : 0; rem -= act ? 1 : 0;                                                        \
      }
#define MSORT_DRAW_LOOP4()                                                                             \
      for (uint32_t b = 0; rem > 0; ++b) {                                                             \
        U4 r4 = env_draw(c, gid_lo, gid_hi, blk0 + b, ep, stp);                                        \
        MSORT_DRAW4(r4.x, r4.y); MSORT_DRAW4(r4.x, r4.y); MSORT_DRAW4(r4.x, r4.y);                     \
        MSORT_DRAW4(r4.x, r4.y); MSORT_DRAW4(r4.x, r4.y); MSORT_DRAW4(r4.x, r4.y);                     \
        if (rem <= 0) break;                                                                           \
        MSORT_DRAW4(r4.z, r4.w); MSORT_DRAW4(r4.z, r4.w); MSORT_DRAW4(r4.z, r4.w);                     \
        MSORT_DRAW4(r4.z, r4.w); MSORT_DRAW4(r4.z, r4.w); MSORT_DRAW4(r4.z, r4.w);                     \
      }
      MSORT_START_STATION1(rem == 0);
      MSORT_DRAW_LOOP4();
      MSORT_START_STATION1(!started1);    // lanes that drew for station 0 move on to station 1 ...
      MSORT_DRAW_LOOP4();                 // ... whose draws (none with the default boosts) run here
      {                                   // station 2: classes (lump, L3)
        const int t = L2;
        const int tv = __double2int_rn(dmul((double)t, acc_sorter[2]));
        rem = t - tv;
        T4 |= (uint32_t)tv << 16; F4 |= (uint32_t)rem << 16;
        tot -= tv; lump += rem;
      }
#define MSORT_DRAW2(xl, xh)                                                                            \
      {                                                                                                \
        const bool act = rem > 0;                                                                      \
        const int r = (int)draw64(xl, xh, (uint32_t)tot);                                              \
        const bool h0 = r < lump;                                                                      \
        lump -= (act && h0) ? 1 : 0;                                                                   \
        L3 -= (act && !h0) ? 1 : 0;                                                                    \
        tot -= act ? 1 : 0; rem -= act ? 1 : 0;                                                        \
      }
      for (uint32_t b = 0; rem > 0; ++b) {
        U4 r4 = env_draw(c, gid_lo, gid_hi, kBlkRedis + 128u + b, ep, stp);
        MSORT_DRAW2(r4.x, r4.y); MSORT_DRAW2(r4.x, r4.y); MSORT_DRAW2(r4.x, r4.y);
        MSORT_DRAW2(r4.x, r4.y); MSORT_DRAW2(r4.x, r4.y); MSORT_DRAW2(r4.x, r4.y);
        if (rem <= 0) break;
        MSORT_DRAW2(r4.z, r4.w); MSORT_DRAW2(r4.z, r4.w); MSORT_DRAW2(r4.z, r4.w);
        MSORT_DRAW2(r4.z, r4.w); MSORT_DRAW2(r4.z, r4.w); MSORT_DRAW2(r4.z, r4.w);
      }
      {                                   // station 3: its draws leave `lump` (the sum of all leftovers) unchanged
        const int t = L3;
        const int tv = __double2int_rn(dmul((double)t, acc_sorter[3]));
        T4 |= (uint32_t)tv << 24; F4 |= (uint32_t)(t - tv) << 24;
      }
#undef MSORT_DRAW2
#undef MSORT_DRAW_LOOP4
#undef MSORT_DRAW4
#undef MSORT_START_STATION1
      s.e += lump;                                               // :579,597
#pragma unroll
      for (int q = 0; q < 4; ++q) { s.tr[q] += b4(T4, q); s.fl[q] += b4(F4, q); }  // :600-602
      sorted_true4 = T4;
    }

    // 6: Env_1 samples its own press action under the mask (env_super.py:291-300)
    if (KIND == MSORT_ENV_SORT) {
      if (RNG == MSORT_RNG_REPLAY) pa = a.press_choice[i];
      else {
        uint32_t vb = press_mask_bits(c, s);
        U4 r4 = env_draw(c, gid_lo, gid_hi, kBlkPress, ep, stp);
        int pick = (int)__umulhi(r4.x, (uint32_t)__popc(vb));
        for (int q = 0; q < pick; ++q) vb &= vb - 1;  // drop the `pick` lowest valid actions
        pa = __ffs(vb) - 1;
      }
      if (pa > 10) pa = 0;
    }
    if (KIND == MSORT_ENV_PRESS && !masking && !press_action_valid(c, s, pa)) { pa = 0; invalid = true; }

    // 7: press_action_rules env_super.py:626-640
    if (!skip_press) {
#pragma unroll
      for (int p = 0; p < 2; ++p) {  // check_press_status :642-659
        if (s.timer[p] > 0) {
          s.timer[p] -= 1;
          if (s.timer[p] == 0) {
            st_bales += (uint32_t)press_bale(c, a.state, c.n_pad, i, s.mat[p], s.pn[p], s.pq[p]);
            s.mat[p] = 0; s.pn[p] = 0; s.pq[p] = 0;
          }
        }
      }
      if (pa != 0) {  // use_press :722-769 (static indexing only: keeps the env in registers)
        const bool second = pa > 5;
        const int m = pa - (second ? 6 : 1);
        if ((second ? s.timer[1] : s.timer[0]) == 0) {
          int tv = 0, amt = s.e;
#pragma unroll
          for (int q = 0; q < 4; ++q) if (m == q) { tv = s.tr[q]; amt = s.tr[q] + s.fl[q]; s.tr[q] = 0; s.fl[q] = 0; }
          if (m == 4) s.e = 0;
          s.started = 1; s.last_amt = amt;
          const int qk = (m < 4 && amt > 0) ? purity_k<SMALL, HOT && !SMALL>(c, tv, amt) : 0;  // round(true/total, 2) (:754)
          if (!second) { s.timer[0] = c.press_time[0]; s.mat[0] = m; s.pn[0] = amt; s.pq[0] = qk; }
          else { s.timer[1] = c.press_time[1]; s.mat[1] = m; s.pn[1] = amt; s.pq[1] = qk; }
        }
      }
    }

    // container levels after sorting and pressing
    int lv[5];
#pragma unroll
    for (int m = 0; m < 4; ++m) lv[m] = s.tr[m] + s.fl[m];
    lv[4] = s.e;

    // 8: overflow termination (detect_overflow :900-905)
    int overflow_mat = -1;
    if (!HOT && (c.flags & MSORT_F_CHECK_OVERFLOW)) {
#pragma unroll
      for (int m = 4; m >= 0; --m) if (lv[m] > c.cap) overflow_mat = m;
    }
    const bool overflow = overflow_mat >= 0;

    // 9: rewards (float64 sums of exactly-rounded terms; emitted as float32)
    int kq[4] = {-1, -1, -1, -1};
    if (KIND != MSORT_ENV_PRESS) {
#pragma unroll
      for (int m = 0; m < 4; ++m) kq[m] = lv[m] > 0 ? purity_k<SMALL, HOT && !SMALL>(c, s.tr[m], lv[m]) : -1;
    }
    double reward, rs_term = 0.0, rp_term = 0.0;   // the two terms are only reported (telemetry)
    bool terminated;
    if (overflow) {
      reward = c.ovf_pen;
      // logged split of the penalty: Env_3 halves it (env_monolith.py:271), Env_1/2 book it as pressing
      if (KIND == MSORT_ENV_MONO) { rs_term = rp_term = 0.5 * c.ovf_pen; } else { rp_term = c.ovf_pen; }
      s.step += 1;
      terminated = true;
    } else {
      double r_sort = 0.0, r_press = 0.0;
      if (KIND != MSORT_ENV_PRESS) {  // calculate_sorting_reward :963-1003
        if (FAST || c.fast_pdiff) {
          int kt = 0;
#pragma unroll
          for (int m = 0; m < 4; ++m) kt += kq[m] >= 0 ? kq[m] : c.qthr100[m];
          r_sort = __ldg(&c.sort_lut[min(max(kt, 0), kSortLut - 1)]);  // 3.2 KB table, L1-resident
        } else {
          r_sort = sort_reward_f64(c, kq[0], kq[1], kq[2], kq[3]);
        }
      }
      if (KIND != MSORT_ENV_SORT) {  // calculate_press_reward :1006-1080
        int mx = lv[4], tl = lv[4];
        double max_pen;        // min(0, severe if any fill>0.95, mild if any fill in (0.90,0.95]) (:1024-1027)
        if (FAST) {            // penalties ordered severe <= mild (host-checked): only the fullest container matters
#pragma unroll
          for (int m = 0; m < 4; ++m) { mx = max(mx, lv[m]); tl += lv[m]; }
          max_pen = mx >= c.lvl_sev ? c.pen_sev0 : (mx >= c.lvl_mild ? c.pen_mild0 : 0.0);
        } else {
          bool sev = lv[4] >= c.lvl_sev, mild = lv[4] >= c.lvl_mild && lv[4] < c.lvl_sev;
#pragma unroll
          for (int m = 0; m < 4; ++m) {
            mx = max(mx, lv[m]); tl += lv[m];
            sev |= lv[m] >= c.lvl_sev; mild |= lv[m] >= c.lvl_mild && lv[m] < c.lvl_sev;
          }
          max_pen = 0.0;
          if (sev && c.pen_sev < max_pen) max_pen = c.pen_sev;
          if (mild && c.pen_mild < max_pen) max_pen = c.pen_mild;
        }
        if (mx >= c.lvl_cat) r_press = c.pen_cat;            // :1022-1023
        else if (max_pen < 0.0) r_press = max_pen;           // :1029-1030 (started flag NOT cleared)
        else {
          double rr = (double)tl * c.c_state;                // overall fill ratio * max_state_reward (:1049-1050)
          if (s.started) {                                   // :1054-1075
            const int S = c.S, amount = s.last_amt;
            const int nb = LAYOUT == LAYOUT_COMPACT ? (int)__umulhi((uint32_t)amount, c.S_magic) : amount / S;  // amount < 2^16 in the compact layout
            const int rm = amount - nb * S;
            const int d = min(rm, S - rm);
            const double eff = (1.0 - (double)d * c.c_eff) * c.bef;
            const double peak = (double)min(nb, 3) * (1.0 / 3.0);
            rr += eff + (peak - c.bef);
            s.started = 0; s.last_amt = 0;
          }
          r_press = clipd(rr, -1.0, 1.0);
        }
      }
      reward = KIND == MSORT_ENV_SORT ? r_sort : (KIND == MSORT_ENV_PRESS ? r_press : r_sort + r_press);
      rs_term = r_sort; rp_term = r_press;
      s.step += 1;
      terminated = s.step >= (uint32_t)c.max_steps;
    }
    s.ep_ret = dadd(s.ep_ret, reward);

    // 10: observation (this env's row of the dense shared tile), outputs, auto-reset
    if (KIND != MSORT_ENV_PRESS) {
      if (FAST) {   // whole-percent thresholds in [0,1] (host-checked): (k - 100*thr)/100 is already inside [-1,1]
#pragma unroll
        for (int m = 0; m < 4; ++m) orow[9 + m] = kq[m] >= 0 ? (float)(kq[m] - c.qthr100[m]) * 0.01f : 0.f;
      } else {
        obs_pdiff(c, kq, orow);
      }
    }
    if (KIND != MSORT_ENV_SORT && !TCMLP) obs_levels_timers(c, s, prow);

    a.reward[i] = (float)reward;
    a.terminated[i] = terminated ? 1 : 0;
    if (!HOT && a.any_step_info) {
      if (a.info_action) a.info_action[i] = act;
      if (a.info_overflow) a.info_overflow[i] = overflow ? 1 : 0;
      if (a.info_overflow_mat) a.info_overflow_mat[i] = (int8_t)overflow_mat;
      if (a.info_sort_mode) a.info_sort_mode[i] = (uint8_t)mode;
      if (a.info_press_action) a.info_press_action[i] = (uint8_t)pa;
      if (a.info_invalid) a.info_invalid[i] = invalid ? 1 : 0;
      if (a.info_sorted_true) a.info_sorted_true[i] = sorted_true4;
      if (a.info_r_sort) a.info_r_sort[i] = (float)rs_term;
      if (a.info_r_press) a.info_r_press[i] = (float)rp_term;
    }
    st_reward = reward;
    st_flags += (overflow ? 1u << 8 : 0u) + (invalid ? 1u << 16 : 0u);
    if (terminated) {
      st_flags += 1u; st_return = s.ep_ret; st_len = s.step;
      if (a.episode_return) a.episode_return[i] = s.ep_ret;
      if (a.episode_length) a.episode_length[i] = (int)s.step;
      if (auto_reset) {
        if (a.terminal_obs) {
          if (TCMLP) put_press_row(c, s, orow);
          // Episodes of one warp normally end together: then the warp's 32 rows are one contiguous
          // range in the tile and in the tensor, copied with coalesced stores.  (All 32 lanes reach
          // this point in that case, so __activemask() names the whole warp.)
          const unsigned tm = __activemask();
          if (tm == 0xffffffffu) {
            __syncwarp();                 // the rows of all 32 lanes are complete in shared memory
            const int w0 = tid & ~31, lane = tid & 31;
            const float* src = &s_obs[w0 * D];
            float* dst = a.terminal_obs + (row0 + w0) * D;
#pragma unroll
            for (int k = 0; k < D; ++k) dst[k * 32 + lane] = src[k * 32 + lane];
            __syncwarp();                 // before any lane overwrites its row with the reset observation
          } else {
            float* t = a.terminal_obs + i * D;
#pragma unroll
            for (int k = 0; k < D; ++k) t[k] = orow[k];
          }
        }
        // unseeded reset (env_super.py:365-420): streams run on, generator re-seeded
        reset_env(c, s);
        s.episode = ep + 1;
        s.gfirst = (int)(env_draw(c, gid_lo, gid_hi, kBlkReset, s.episode, 0u).x & 1u);
        zero_cold(a.state, c.n_pad, i);
        if (!TCMLP) env_obs<KIND>(c, s, orow);
      }
    }
    // TCMLP: the whole 16-wide row is a function of the final state — assembled in registers and stored as four 16-byte
    // vectors (scalar stores at the row pitch of 64 B are 16-way bank conflicts: they were 57 % of the kernel's
    // shared-memory store wavefronts)
    if (TCMLP) put_press_row(c, s, orow);
    if (want_mask) { mbits = press_mask_bits(c, s); put_mask_row<A>(s_mask, tid, mbits); }
    store_planes<LAYOUT>(a.state, c.n_pad, i, s);
  }

  if (a.stats) {  // warp reduce (REDUX for the integer counters) -> shared -> one atomic per CTA per slot
    const uint32_t f = __reduce_add_sync(0xffffffffu, st_flags);
    const uint32_t nb = __reduce_add_sync(0xffffffffu, st_bales);
    const uint32_t nl = __reduce_add_sync(0xffffffffu, st_len);
    const uint32_t nu = RNG == MSORT_RNG_REPLAY ? __reduce_add_sync(0xffffffffu, st_underrun) : 0u;
    const uint32_t ns = __popc(__ballot_sync(0xffffffffu, live));
    const double rw = warp_sum(st_reward);
    const double rt = (f & 0xffu) ? warp_sum(st_return) : 0.0;  // warp-uniform condition
    if ((tid & 31) == 0) {
      double* w = s_stat[tid >> 5];
      w[ST_EPISODES] = (double)(f & 0xffu); w[ST_RETURN] = rt; w[ST_LENGTH] = (double)nl;
      w[ST_STEPS] = (double)ns; w[ST_REWARD] = rw;
      w[ST_OVERFLOW] = (double)((f >> 8) & 0xffu); w[ST_BALES] = (double)nb;
      w[ST_INVALID] = (double)((f >> 16) & 0xffu); w[ST_CLAMPED] = (double)(f >> 24); w[ST_UNDERRUN] = (double)nu;
    }
  }
  fence_async_smem();  // order this thread's tile writes before the async-proxy reads below
  __syncthreads();
  if (a.stats && tid >= kStatTid && tid < kStatTid + ST_COUNT) {
    double v = 0.0;
#pragma unroll
    for (int w = 0; w < kTile / 32; ++w) v += s_stat[w][tid - kStatTid];
    if (v != 0.0) atomicAdd(&a.stats[tid - kStatTid], v);
  }
  if (rows == kTile) {
    if (tid == kStoreTid) {
      bulk_store(a.obs + row0 * D, s_obs, kTile * D * (uint32_t)sizeof(float));
      if (want_mask) bulk_store(a.mask + row0 * A, s_mask, kTile * A);
      if (PERSIST || FUSE) asm volatile("cp.async.bulk.commit_group;" ::: "memory");   // waited for before the tile is written again
      else bulk_commit_and_wait_read();
    }
  } else {  // partial last tile
    flush_tile(s_obs, a.obs + row0 * D, rows * D * (int)sizeof(float));
    if (want_mask) flush_tile(s_mask, a.mask + row0 * A, rows * A);
  }
  if constexpr (FUSE) {
    // ================================================================ the NEXT step's policy on the tile just built
    // (1) this env's observation row -> the fp16 words of the layer-1 operand (K 0..28 = obs, K 29 = 1.0 for the bias)
    uint32_t a1[16];
    {
      const float* x = &s_obs[tid * D];       // pitch 29 floats: conflict-free
#pragma unroll
      for (int q = 0; q < 14; ++q) a1[q] = pack_h2(x[2 * q], x[2 * q + 1]);
      a1[14] = pack_h2(x[28], 1.0f);
      a1[15] = 0u;
    }
    // (2) the operand buffer aliases the observation tile: its bulk store must have read it out first
    if (rows == kTile && tid == kStoreTid) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    __syncthreads();
    // (the barrier inside rollout_policy_tile, before its first MMA, also orders the TMEM allocation before this read's use)
    const uint32_t tmem = s_tmem;
    {
      const PolicyOut po{a.next_actions, a.next_logp, a.next_value, a.draw_key0, a.draw_key1, a.draw_t, a.draw_t_dev, a.deterministic};
      PolicyPhase ph{0u, 0u};
      rollout_policy_tile(s_tile_raw, s_tcw, &s_wbar, &s_mma, tmem, a1, mbits | (mbits << 11), live, i, c.gid0 + i, tid, po, ph, []() {});
    }
    __syncthreads();
    if (tid < 32) umma::tmem_dealloc(tmem, 64u);
    if (rows == kTile && tid == kStoreTid) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
  }
  tile += gridDim.x;
  } while (PERSIST && tile < ntiles);
  if (PERSIST && tid == kStoreTid) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // shared memory outlives the last copies
  if (TCMLP) {
    umma::fence_before_sync();
    __syncthreads();
    if (tid < 32) umma::tmem_dealloc(s_tmem, 32u);
  }
}

// ---------------------------------------------------------------- K2: reset
template <int KIND>
__global__ void __launch_bounds__(kTile)
reset_kernel(const __grid_constant__ DevConfig c, uint4* __restrict__ state, const uint8_t* __restrict__ which,
             const uint8_t* __restrict__ first_pattern, float* __restrict__ obs, uint8_t* __restrict__ mask,
             uint32_t reset_flags) {
  constexpr int D = Dims<KIND>::D, A = Dims<KIND>::A;
  __shared__ __align__(16) float s_obs[kTile * D];
  __shared__ __align__(16) uint8_t s_mask[kTile * A];
  const long long row0 = (long long)blockIdx.x * kTile;
  const long long i = row0 + threadIdx.x;
  const int rows = (int)min((long long)kTile, c.n - row0);
  const bool selected = i < c.n && (!which || which[i]);
  if (selected) {
    Env s;
    s.episode = 0; s.cursor = 0;
    if (reset_flags & MSORT_RESET_KEEP_STREAMS) {  // reset(seed=None): streams run on (env_super.py:377)
      load_env(c, state, i, s);
      s.episode += 1;
    }
    reset_env(c, s);
    int fp = first_pattern ? first_pattern[i] : 0;
    if (fp == 1 || fp == 2) s.gfirst = fp - 1;
    else {
      const unsigned long long g = (unsigned long long)(c.gid0 + i);
      s.gfirst = (int)(env_draw(c, (uint32_t)g, (uint32_t)(g >> 32) & 0xffffu, kBlkReset, s.episode, 0u).x & 1u);
    }
    zero_cold(state, c.n_pad, i);
    store_env(c, state, i, s);
    float* orow = &s_obs[threadIdx.x * D];
    env_obs<KIND>(c, s, orow);
    const uint32_t bits = press_mask_bits(c, s);
    put_mask_row<A>(s_mask, threadIdx.x, bits);
    if (which) {  // partial reset: only the selected rows may be written
      if (obs) for (int k = 0; k < D; ++k) obs[i * D + k] = orow[k];
      if (mask) for (int k = 0; k < A; ++k) mask[i * A + k] = s_mask[threadIdx.x * A + k];
    }
  }
  if (!which) {  // full reset: coalesced flush of the whole tile
    __syncthreads();
    if (obs) flush_tile(s_obs, obs + row0 * D, rows * D * (int)sizeof(float));
    if (mask) flush_tile(s_mask, mask + row0 * A, rows * A);
  }
}

// ---------------------------------------------------------------- observe (no transition)
template <int KIND>
__global__ void __launch_bounds__(kTile)
observe_kernel(const __grid_constant__ DevConfig c, const uint4* __restrict__ state, float* __restrict__ obs,
               uint8_t* __restrict__ mask, int after_shift) {
  constexpr int D = Dims<KIND>::D, A = Dims<KIND>::A;
  __shared__ __align__(16) float s_obs[kTile * D];
  __shared__ __align__(16) uint8_t s_mask[kTile * A];
  const long long row0 = (long long)blockIdx.x * kTile;
  const long long i = row0 + threadIdx.x;
  const int rows = (int)min((long long)kTile, c.n - row0);
  if (i < c.n) {
    Env s;
    load_env(c, state, i, s);
    // what the agents of Env_3.step(mode='model') see (env_monolith.py:114-115,186-221): the plant after
    // update_environment has moved input -> belt -> sorting, before anything else of the step happened
    if (after_shift) { s.sort4 = s.belt4; s.belt4 = s.in4; }
    env_obs<KIND>(c, s, &s_obs[threadIdx.x * D]);
    put_mask_row<A>(s_mask, threadIdx.x, press_mask_bits(c, s));
  }
  __syncthreads();
  if (obs) flush_tile(s_obs, obs + row0 * D, rows * D * (int)sizeof(float));
  if (mask) flush_tile(s_mask, mask + row0 * A, rows * A);
}

// ---------------------------------------------------------------- masked-random action source
// One thread per env; the CTA's mask rows (contiguous kTile*A bytes) are staged through shared
// memory with coalesced 4-byte loads, then each thread scans its own row.
template <int A>
__global__ void __launch_bounds__(kTile)
sample_kernel(const __grid_constant__ DevConfig c, const uint8_t* __restrict__ mask, long long* __restrict__ actions,
              unsigned key0, unsigned key1, unsigned t) {
  __shared__ uint32_t s_rows[kTile * A / 4];
  const long long row0 = (long long)blockIdx.x * kTile;
  const int rows = (int)min((long long)kTile, c.n - row0);
  const uint8_t* src = mask + row0 * A;
  const int total = rows * A, words = total >> 2;
  for (int w = threadIdx.x; w < words; w += kTile) s_rows[w] = reinterpret_cast<const uint32_t*>(src)[w];
  for (int e = 4 * words + threadIdx.x; e < total; e += kTile) reinterpret_cast<uint8_t*>(s_rows)[e] = src[e];
  __syncthreads();
  if ((int)threadIdx.x >= rows) return;
  const uint8_t* row = reinterpret_cast<const uint8_t*>(s_rows) + threadIdx.x * A;
  uint32_t bits = 0;
#pragma unroll
  for (int k = 0; k < A; ++k) bits |= (uint32_t)(row[k] != 0) << k;
  unsigned long long g = (unsigned long long)(c.gid0 + row0 + threadIdx.x);
  U4 r = philox4x32_10((uint32_t)g, (uint32_t)(g >> 32), 0x5A3Cu, t, key0, key1);
  int nv = __popc(bits), a = 0;
  if (nv > 0) {
    int pick = (int)__umulhi(r.x, (uint32_t)nv);
    for (int q = 0; q < pick; ++q) bits &= bits - 1;
    a = __ffs(bits) - 1;
  }
  actions[row0 + threadIdx.x] = a;
}

// ---------------------------------------------------------------- K3: generator stream export
// The state-independent random inputs of `num_steps` steps of one episode, written in the REPLAY descriptor's form
// (ref: SeasonalInputGenerator.generate_input input_generator.py:37-64 for the batches, rng_noise.uniform
// env_super.py:508 for the noise): what the PHILOX step kernel will consume for (env, episode, step).  Together with the
// oracle's recording of the redistribution draws (which depend on the plant state) a PHILOX trajectory can be replayed
// through the REPLAY instantiation and through the reference.  One thread per env, coalesced [t][env] stores.
__global__ void __launch_bounds__(256)
generate_streams_kernel(const __grid_constant__ DevConfig c, uint32_t episode, uint32_t first_step, uint32_t num_steps,
                        uint32_t* __restrict__ input_counts, double* __restrict__ noise_u, uint32_t* __restrict__ draw_words,
                        uint8_t* __restrict__ first_pattern) {
  const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
  if (i >= c.n) return;
  const unsigned long long g = (unsigned long long)(c.gid0 + i);
  const uint32_t gid_lo = (uint32_t)g, gid_hi = (uint32_t)(g >> 32) & 0xffffu;
  const uint32_t gfirst = env_draw(c, gid_lo, gid_hi, kBlkReset, episode, 0u).x & 1u;   // pattern_sequence[0] - 1 of this episode
  if (first_pattern) first_pattern[i] = (uint8_t)(1u + gfirst);
  for (uint32_t k = 0; k < num_steps; ++k) {
    const uint32_t t = first_step + k;
    const long long o = (long long)k * c.n + i;
    if (input_counts) {   // step t emits batch number t of the episode: the pattern flips every steps_per_pattern batches
      uint32_t in4 = c.pat[((t / (uint32_t)c.spp) & 1u) ^ gfirst];
      for (int r = 0; r < c.pat_remainder; ++r) {
        const U4 r4 = env_draw(c, gid_lo, gid_hi, kBlkInput + 0x100u * (uint32_t)(r >> 2), episode, t);
        const uint32_t x = (r & 3) == 0 ? r4.x : ((r & 3) == 1 ? r4.y : ((r & 3) == 2 ? r4.z : r4.w));
        in4 += 1u << (8 * (x & 3u));
      }
      input_counts[o] = in4;
    }
    if (noise_u) {
      const U4 r4 = env_draw(c, gid_lo, gid_hi, kBlkNoise, episode, t);
      double2* d = reinterpret_cast<double2*>(noise_u) + 2 * o;
      d[0] = make_double2((double)r4.x * 2.3283064365386963e-10, (double)r4.y * 2.3283064365386963e-10);
      d[1] = make_double2((double)r4.z * 2.3283064365386963e-10, (double)r4.w * 2.3283064365386963e-10);
    }
    if (draw_words) {     // first block of redistribution words of stations 0, 1, 2 (station 3's draws are never simulated)
      uint4* w = reinterpret_cast<uint4*>(draw_words) + 3 * o;
#pragma unroll
      for (int s = 0; s < 3; ++s) {
        const U4 r4 = env_draw(c, gid_lo, gid_hi, kBlkRedis + 64u * (uint32_t)s, episode, t);
        w[s] = make_uint4(r4.x, r4.y, r4.z, r4.w);
      }
    }
  }
}

// ---------------------------------------------------------------- host-buffer step (msort_step_host): tiny helpers
// actions arrive from the host as one byte per env (Discrete(22) fits): widen to the int64 the step kernel reads
__global__ void __launch_bounds__(256)
widen_actions_kernel(const uint8_t* __restrict__ in, long long* __restrict__ out, long long n) {
  const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
  if (i < n) out[i] = in[i];
}

// one 16-bit word per env for the trip back: bits 0..10 = the 11 press-mask bits (Env_1: its two always-valid
// actions), bit 15 = terminated.  Mask rows are staged through shared memory like sample_kernel does.
template <int A>
__global__ void __launch_bounds__(kTile)
pack_flags_kernel(const uint8_t* __restrict__ mask, const uint8_t* __restrict__ terminated, uint16_t* __restrict__ flags, long long n) {
  __shared__ uint32_t s_rows[kTile * A / 4];
  const long long row0 = (long long)blockIdx.x * kTile;
  const int rows = (int)min((long long)kTile, n - row0);
  const uint8_t* src = mask + row0 * A;
  const int total = rows * A, words = total >> 2;
  for (int w = threadIdx.x; w < words; w += kTile) s_rows[w] = reinterpret_cast<const uint32_t*>(src)[w];
  for (int e = 4 * words + threadIdx.x; e < total; e += kTile) reinterpret_cast<uint8_t*>(s_rows)[e] = src[e];
  __syncthreads();
  if ((int)threadIdx.x >= rows) return;
  const uint8_t* row = reinterpret_cast<const uint8_t*>(s_rows) + threadIdx.x * A;
  uint32_t bits = 0;
#pragma unroll
  for (int k = 0; k < (A < 11 ? A : 11); ++k) bits |= (uint32_t)(row[k] != 0) << k;
  flags[row0 + threadIdx.x] = (uint16_t)(bits | (terminated[row0 + threadIdx.x] ? 0x8000u : 0u));
}

// ---------------------------------------------------------------- rule-based action source
// ref: Env_3.step(mode='rule_based') env_monolith.py:166-184
__global__ void __launch_bounds__(kTile)
rule_actions_kernel(const __grid_constant__ DevConfig c, const uint4* __restrict__ state, int after_shift,
                    long long* __restrict__ actions) {
  const long long i = (long long)blockIdx.x * kTile + threadIdx.x;
  if (i >= c.n) return;
  Env s;
  load_env(c, state, i, s);
  // sorting_rules env_super.py:469-482: mode 0 iff pA+pC > pB+pD (float64 proportions; empty belt -> 1)
  const uint32_t belt = after_shift ? s.in4 : s.belt4;
  const int ac = b4(belt, 0) + b4(belt, 2), bd = b4(belt, 1) + b4(belt, 3), bt = ac + bd;
  int mode;
  if (ac != bd) mode = ac > bd ? 0 : 1;
  else {
    double p[4];
    for (int m = 0; m < 4; ++m) p[m] = bt > 0 ? ddiv((double)b4(belt, m), (double)bt) : 0.0;
    mode = dadd(p[0], p[2]) > dadd(p[1], p[3]) ? 0 : 1;
  }
  // check_container_level env_super.py:689-720
  int press = 0;
  const int free_press = s.timer[0] == 0 ? 1 : (s.timer[1] == 0 ? 2 : 0);
  if (free_press) {
    int best = 0, idx = -1;
    for (int m = 0; m < 4; ++m) { const int l = s.tr[m] + s.fl[m]; if (l > best) { best = l; idx = m; } }
    if (s.e > best) { best = s.e; idx = 4; }
    if (best > 0) press = (free_press - 1) * 5 + idx + 1;   // press_action_to_discrete :864-867
  }
  actions[i] = c.kind == MSORT_ENV_SORT ? mode : (c.kind == MSORT_ENV_PRESS ? press : 11 * mode + press);
}

// ---------------------------------------------------------------- K6: export / import
__device__ __forceinline__ void export_env(const DevConfig& c, const uint4* __restrict__ state, long long i,
                                           msort_env_state_t* __restrict__ dst) {
  Env s;
  load_env(c, state, i, s);
  msort_env_state_t o;
  for (int m = 0; m < 4; ++m) {
    o.input[m] = b4(s.in4, m); o.belt[m] = b4(s.belt4, m); o.sorting[m] = b4(s.sort4, m);
    o.cont_true[m] = s.tr[m]; o.cont_false[m] = s.fl[m]; o.acc_belt[m] = s.acc[m];
  }
  o.cont_e = s.e;
  for (int p = 0; p < 2; ++p) {
    o.press_timer[p] = s.timer[p]; o.press_mat[p] = s.mat[p]; o.press_n[p] = s.pn[p]; o.press_q[p] = s.pq[p];
  }
  o.last_press_started = s.started; o.last_press_amount = s.last_amt;
  o.gen_first = s.gfirst + 1; o.gen_idx = s.gidx; o.gen_counter = s.gcount;
  o.step = (int)s.step; o.episode = (int)s.episode; o.sensor_mode = s.mode; o.replay_cursor = s.cursor;
  for (int m = 0; m < 5; ++m) {
    uint4 v = state[(kHotPlanes + m) * c.n_pad + i];
    o.bale_n[m] = (int)v.x; o.bale_sum[m] = (int)v.y;
    o.bale_last_size[m] = (int)(v.z & 0xffffffu); o.bale_last_q[m] = (int)(v.z >> 24);
  }
  o.reserved = 0;
  o.ep_return = s.ep_ret;
  *dst = o;
}

__global__ void __launch_bounds__(kTile)
export_kernel(const __grid_constant__ DevConfig c, const uint4* __restrict__ state, msort_env_state_t* __restrict__ out) {
  const long long i = (long long)blockIdx.x * kTile + threadIdx.x;
  if (i >= c.n) return;
  export_env(c, state, i, &out[i]);
}

// telemetry snapshot: plain state of the listed envs (ids outside [0, n) give a zeroed record)
__global__ void __launch_bounds__(kTile)
gather_kernel(const __grid_constant__ DevConfig c, const uint4* __restrict__ state, const long long* __restrict__ ids,
              long long count, msort_env_state_t* __restrict__ out) {
  const long long j = (long long)blockIdx.x * kTile + threadIdx.x;
  if (j >= count) return;
  const long long i = ids[j];
  if (i < 0 || i >= c.n) { memset(&out[j], 0, sizeof(msort_env_state_t)); return; }
  export_env(c, state, i, &out[j]);
}

__global__ void __launch_bounds__(kTile)
import_kernel(const __grid_constant__ DevConfig c, uint4* __restrict__ state, const msort_env_state_t* __restrict__ in) {
  const long long i = (long long)blockIdx.x * kTile + threadIdx.x;
  if (i >= c.n) return;
  const msort_env_state_t o = in[i];
  Env s;
  s.in4 = s.belt4 = s.sort4 = 0;
  for (int m = 0; m < 4; ++m) {
    s.in4 |= (uint32_t)(o.input[m] & 0xff) << (8 * m);
    s.belt4 |= (uint32_t)(o.belt[m] & 0xff) << (8 * m);
    s.sort4 |= (uint32_t)(o.sorting[m] & 0xff) << (8 * m);
    s.tr[m] = o.cont_true[m]; s.fl[m] = o.cont_false[m]; s.acc[m] = o.acc_belt[m];
  }
  s.e = o.cont_e;
  for (int p = 0; p < 2; ++p) {
    s.timer[p] = o.press_timer[p]; s.mat[p] = o.press_mat[p]; s.pn[p] = o.press_n[p]; s.pq[p] = o.press_q[p];
  }
  s.started = o.last_press_started != 0; s.last_amt = o.last_press_amount;
  s.gfirst = o.gen_first == 2 ? 1 : 0; s.gidx = o.gen_idx & 1; s.gcount = o.gen_counter;
  s.step = (uint32_t)o.step; s.episode = (uint32_t)o.episode; s.mode = o.sensor_mode & 1; s.cursor = o.replay_cursor;
  s.ep_ret = o.ep_return;
  store_env(c, state, i, s);
  for (int m = 0; m < 5; ++m)
    state[(kHotPlanes + m) * c.n_pad + i] =
        make_uint4((uint32_t)o.bale_n[m], (uint32_t)o.bale_sum[m],
                   ((uint32_t)o.bale_last_size[m] & 0xffffffu) | ((uint32_t)o.bale_last_q[m] << 24), 0u);
}

// ---------------------------------------------------------------- K5: state statistics
__global__ void __launch_bounds__(256)
stats_kernel(const __grid_constant__ DevConfig c, const uint4* __restrict__ state, double* __restrict__ out16) {
  __shared__ double sh[MSORT_NUM_STATS][8];
  double v[MSORT_NUM_STATS];
#pragma unroll
  for (int k = 0; k < MSORT_NUM_STATS; ++k) v[k] = 0.0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < c.n; i += (long long)gridDim.x * blockDim.x) {
    Env s;
    load_env(c, state, i, s);
    v[0] += 1.0;
    double lvl = (double)s.e, pm = 0.0;
    int kq[4];
    purity_ks(c, s, kq);
#pragma unroll
    for (int m = 0; m < 4; ++m) { lvl += (double)(s.tr[m] + s.fl[m]); pm += kq[m] >= 0 ? 0.01 * (double)kq[m] : c.qthr_empty[m]; }
    v[1] += lvl;
#pragma unroll
    for (int m = 0; m < 5; ++m) {
      uint4 b = state[(kHotPlanes + m) * c.n_pad + i];
      v[2 + m] += (double)b.x; v[7 + m] += (double)b.y;
    }
    v[12] += pm * 0.25;
    v[13] += (double)((s.timer[0] > 0) + (s.timer[1] > 0));
    v[14] += s.ep_ret;
    v[15] += (double)s.step;
  }
#pragma unroll
  for (int k = 0; k < MSORT_NUM_STATS; ++k) {
    double x = v[k];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
    if ((threadIdx.x & 31) == 0) sh[k][threadIdx.x >> 5] = x;
  }
  __syncthreads();
  if (threadIdx.x < MSORT_NUM_STATS) {
    double x = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) x += sh[threadIdx.x][w];
    atomicAdd(&out16[threadIdx.x], x);
  }
}

// ---------------------------------------------------------------- launch wrappers
// Embedded-policy weights: SB3 order (include/msort.h: W1[32][13] b1[32] W2[32][32] b2[32] W3[2][32] b3[2]) ->
// the paired layout mlp_sort_mode() consumes: for every pair of output neurons (2p, 2p+1) and every input k
// the two weights side by side (one FFMA2 operand); the last layer pairs the two logits.  Section offsets
// and the bias sections are unchanged.
void pack_policy_pairs(const float* sb3, float* paired) {
  constexpr int W1 = 0, b1 = 416, W2 = 448, b2 = 1472, W3 = 1504, b3 = 1568;
  for (int jp = 0; jp < 16; ++jp)
    for (int k = 0; k < 13; ++k) {
      paired[W1 + (jp * 13 + k) * 2] = sb3[W1 + (2 * jp) * 13 + k];
      paired[W1 + (jp * 13 + k) * 2 + 1] = sb3[W1 + (2 * jp + 1) * 13 + k];
    }
  for (int jp = 0; jp < 16; ++jp)
    for (int k = 0; k < 32; ++k) {
      paired[W2 + (jp * 32 + k) * 2] = sb3[W2 + (2 * jp) * 32 + k];
      paired[W2 + (jp * 32 + k) * 2 + 1] = sb3[W2 + (2 * jp + 1) * 32 + k];
    }
  for (int j = 0; j < 32; ++j) { paired[W3 + 2 * j] = sb3[W3 + j]; paired[W3 + 2 * j + 1] = sb3[W3 + 32 + j]; }
  for (int j = 0; j < 32; ++j) { paired[b1 + j] = sb3[b1 + j]; paired[b2 + j] = sb3[b2 + j]; }
  paired[b3] = sb3[b3]; paired[b3 + 1] = sb3[b3 + 1];
}

static inline unsigned tiles(long long n) { return (unsigned)((n + kTile - 1) / kTile); }
int policy_tc_words() { return kTcWords; }

// Env_2's embedded policy for the tensor-core path (tc_mlp_mode): SB3 order -> kTcWords packed words.
//   layer 1: B[n][k] = c*W1[n][k] (k < 13), rows 13..15 of term 0 = the three split terms of c*b1[n]      (c = 2 log2 e)
//   layer 2: B[n][k] = -2c*W2[n][k],  bias2'[n] = c*(b2[n] + sum_k W2[n][k])      (input r1 with h1 = 1 - 2 r1)
//   layer 3: fp32 pairs (-2 W3[0][k], -2 W3[1][k]),  bias3'[n] = b3[n] + sum_k W3[n][k]      (FFMA2 in the kernel)
// every MMA weight v (float64) as fp16 terms t1 = fp16(v), t2 = fp16(v - t1), t3 = fp16(v - t1 - t2).  Returns false when a
// value does not fit fp16's range (the caller keeps the FFMA2 kernel for such a policy).
bool pack_policy_tc(const float* sb3, uint32_t* words) {
  constexpr int W1 = 0, b1 = 416, W2 = 448, b2 = 1472, W3 = 1504, b3 = 1568;
  const double cc = 2.0 * 1.4426950408889634;
  __half* hw = reinterpret_cast<__half*>(words);
  float* fw = reinterpret_cast<float*>(words);
  memset(words, 0, sizeof(uint32_t) * kTcWords);
  bool ok = true;
  auto put = [&](int base, int K, int N, int nterms, int n, int k, double v) {   // tile element [k][n] of every term
    if (!(std::fabs(v) < 32768.0)) { ok = false; return; }
    double r = v;
    for (int t = 0; t < nterms; ++t) {
      const __half q = __float2half_rn((float)r);
      hw[base + t * K * N + ((k >> 3) * N + n) * 8 + (k & 7)] = q;
      r -= (double)__half2float(q);
    }
  };
  for (int n = 0; n < 32; ++n) {
    for (int k = 0; k < 13; ++k) put(kTcB1, 16, 32, 3, n, k, cc * (double)sb3[W1 + n * 13 + k]);
    double r = cc * (double)sb3[b1 + n];           // the bias: its three terms sit in rows 13..15 of term 0 (A holds 1.0 there)
    if (!(std::fabs(r) < 32768.0)) ok = false;
    for (int t = 0; t < 3 && ok; ++t) {
      const __half q = __float2half_rn((float)r);
      hw[kTcB1 + ((13 + t) / 8 * 32 + n) * 8 + ((13 + t) & 7)] = q;
      r -= (double)__half2float(q);
    }
  }
  for (int n = 0; n < 32; ++n) {
    double sum = (double)sb3[b2 + n];
    for (int k = 0; k < 32; ++k) {
      const double w = (double)sb3[W2 + n * 32 + k];
      put(kTcB2, 32, 32, 3, n, k, -2.0 * cc * w);
      sum += w;
    }
    fw[kTcBias2 + n] = (float)(cc * sum);
  }
  for (int j = 0; j < 32; ++j) {                   // output layer, fp32: pair j = (-2 W3[0][j], -2 W3[1][j]) (exact scaling)
    fw[kTcW3 + 2 * j] = -2.0f * sb3[W3 + j];
    fw[kTcW3 + 2 * j + 1] = -2.0f * sb3[W3 + 32 + j];
  }
  for (int n = 0; n < 2; ++n) {
    double sum = (double)sb3[b3 + n];
    for (int k = 0; k < 32; ++k) sum += (double)sb3[W3 + n * 32 + k];
    fw[kTcBias3 + n] = (float)sum;
  }
  for (int j = 0; j < MSORT_POLICY_WEIGHTS; ++j) if (!std::isfinite(sb3[j])) ok = false;
  return ok;
}

// StepArgs of the env range starting `first` envs into the launch (whole-tile starts; D / A of the env kind)
static StepArgs shift_args(StepArgs a, long long first, int D, int A) {
  a.state += first; a.actions += first; a.obs += first * D; a.reward += first; a.terminated += first;
  if (a.mask) a.mask += first * A;
  if (a.info_action) a.info_action += first;
  if (a.info_overflow) a.info_overflow += first;
  if (a.info_overflow_mat) a.info_overflow_mat += first;
  if (a.info_sort_mode) a.info_sort_mode += first;
  if (a.info_press_action) a.info_press_action += first;
  if (a.info_invalid) a.info_invalid += first;
  if (a.info_sorted_true) a.info_sorted_true += first;
  if (a.info_r_sort) a.info_r_sort += first;
  if (a.info_r_press) a.info_r_press += first;
  if (a.terminal_obs) a.terminal_obs += first * D;
  if (a.episode_return) a.episode_return += first;
  if (a.episode_length) a.episode_length += first;
  a.act_tma = ((uintptr_t)a.actions & 15u) == 0;
  return a;
}

// Resident CTAs per SM of the four persistent Env_2 kernels (index = SMALL + 2 * TCMLP), asked from the occupancy
// calculator on the CURRENT device (msort_create); also sets the shared-memory carve-out those kernels want.
void query_persist_occupancy(int per_sm[4]) {
  int dev = 0;
  cudaDeviceProp prop;
  cudaGetDevice(&dev);
  const bool have_prop = cudaGetDeviceProperties(&prop, dev) == cudaSuccess;
  auto ask = [&](auto kern) {
    int v = 0;
    cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&v, kern, kTile, 0) != cudaSuccess) v = 0;
    // the same bound from the kernel's own resource use (registers, static shared memory + 1 KB the driver reserves
    // per CTA): the calculator has been seen to answer for the default carve-out rather than the preferred one
    cudaFuncAttributes fa;
    if (have_prop && cudaFuncGetAttributes(&fa, kern) == cudaSuccess) {
      const int by_regs = prop.regsPerMultiprocessor / std::max(1, ((fa.numRegs + 7) / 8 * 8) * kTile);
      const int by_smem = (int)(prop.sharedMemPerMultiprocessor / (fa.sharedSizeBytes + 1024));
      const int mine = std::min(std::min(by_regs, by_smem), prop.maxBlocksPerMultiProcessor);
      v = std::max(v, mine);
    }
    return std::max(v, 1);
  };
  per_sm[0] = ask(step_kernel<MSORT_ENV_PRESS, MSORT_RNG_PHILOX, LAYOUT_COMPACT, true, true, false, false>);
  per_sm[1] = ask(step_kernel<MSORT_ENV_PRESS, MSORT_RNG_PHILOX, LAYOUT_COMPACT, true, true, true, false>);
  per_sm[2] = ask(step_kernel<MSORT_ENV_PRESS, MSORT_RNG_PHILOX, LAYOUT_COMPACT, true, true, false, true>);
  per_sm[3] = ask(step_kernel<MSORT_ENV_PRESS, MSORT_RNG_PHILOX, LAYOUT_COMPACT, true, true, true, true>);
  cudaGetLastError();
}

template <int KIND>
static cudaError_t launch_step_kind(const DevConfig& c, const StepArgs& a, const float* policy_host, int rng, cudaStream_t st, int* variant,
                                    int allow_hot, const int* persist_per_sm) {
  const unsigned g = tiles(c.n);
  int dummy;
  int& var = variant ? *variant : dummy;
  PolicyParam<KIND> pw;
  if constexpr (KIND == MSORT_ENV_PRESS) {
    memset(pw.w, 0, sizeof(pw.w));
    if (policy_host) memcpy(pw.w, policy_host, MSORT_POLICY_WEIGHTS * sizeof(float));
  }
  var = rng == MSORT_RNG_REPLAY ? MSORT_STEP_REPLAY : (c.fast ? MSORT_STEP_FAST : MSORT_STEP_GENERIC);
  if (a.fused_w && (rng == MSORT_RNG_REPLAY || c.layout != LAYOUT_COMPACT)) return cudaErrorNotSupported;
  if (rng == MSORT_RNG_REPLAY) step_kernel<KIND, MSORT_RNG_REPLAY, LAYOUT_REPLAY, false><<<g, kTile, 0, st>>>(c, a, pw);
  else if (c.layout == LAYOUT_COMPACT) {
    const unsigned want = MSORT_F_ACTION_MASKING | MSORT_F_AUTO_RESET, never = MSORT_F_CHECK_OVERFLOW;
    const bool hot = allow_hot && c.fast && c.one_block && (c.flags & want) == want && !(c.flags & never) && a.mask && !a.any_step_info;
    if (a.fused_w && !(hot && KIND == MSORT_ENV_MONO)) return cudaErrorNotSupported;   // FUSE exists for Env_3's HOT configuration only
    if (hot) {
      auto kern = c.small_lv ? step_kernel<KIND, MSORT_RNG_PHILOX, LAYOUT_COMPACT, true, true, true>
                             : step_kernel<KIND, MSORT_RNG_PHILOX, LAYOUT_COMPACT, true, true, false>;
      var = MSORT_STEP_HOT;
      if (a.fused_w) {   // Env_3 rollout kernel: step + the next step's policy (launch_step refuses other kinds)
        if constexpr (KIND == MSORT_ENV_MONO) {
          var = MSORT_STEP_HOT_FUSED;
          if (c.small_lv) step_kernel<KIND, MSORT_RNG_PHILOX, LAYOUT_COMPACT, true, true, true, false, true><<<g, kTile, 0, st>>>(c, a, pw);
          else step_kernel<KIND, MSORT_RNG_PHILOX, LAYOUT_COMPACT, true, true, false, false, true><<<g, kTile, 0, st>>>(c, a, pw);
          return cudaGetLastError();
        }
      }
      if constexpr (MSORT_HOT_PERSIST && KIND == MSORT_ENV_PRESS) {
        // persistent: exactly one wave of resident CTAs (per-SM count from the occupancy calculator, cached in the handle)
        var = MSORT_STEP_HOT_PERSISTENT;
        const int sm = c.small_lv ? 1 : 0;
        auto wave = [&](long long n, int tc) { return std::min(tiles(n), (unsigned)(c.sm_count * std::max(1, persist_per_sm[sm + 2 * tc]))); };
        const long long n_full = c.n / kTile * kTile;
        if (a.policy_tc && (c.flags & MSORT_F_SORT_POLICY_MLP)) {
          // embedded policy on the tensor cores: whole tiles only; a ragged tail goes to the FFMA2 kernel below
          var = MSORT_STEP_HOT_TENSOR;
          if (n_full > 0 && a.mode_scratch) {   // split form: the policy kernel writes one mode byte per env, the policy-free step kernel reads it
            DevConfig cf = c; cf.n = n_full;
            StepArgs as = a; as.sort_mode_in = a.mode_scratch;
            press_policy_kernel<<<tiles(n_full), kTile, 0, st>>>(cf, a.state, a.policy_tc, a.mode_scratch);
            auto sk = c.small_lv ? step_kernel<KIND, MSORT_RNG_PHILOX, LAYOUT_COMPACT, true, true, true, false, false, true>
                                 : step_kernel<KIND, MSORT_RNG_PHILOX, LAYOUT_COMPACT, true, true, false, false, false, true>;
            sk<<<tiles(n_full), kTile, 0, st>>>(cf, as, pw);
            var = MSORT_STEP_HOT_TENSOR_SPLIT;
          } else if (n_full > 0) {
            DevConfig cf = c; cf.n = n_full;
            auto tk = c.small_lv ? step_kernel<KIND, MSORT_RNG_PHILOX, LAYOUT_COMPACT, true, true, true, true>
                                 : step_kernel<KIND, MSORT_RNG_PHILOX, LAYOUT_COMPACT, true, true, false, true>;
            tk<<<MSORT_TC_PERSIST ? wave(n_full, 1) : tiles(n_full), kTile, 0, st>>>(cf, a, pw);
          }
          if (c.n > n_full) {
            DevConfig ct = c; ct.n = c.n - n_full; ct.gid0 += n_full;
            kern<<<1, kTile, 0, st>>>(ct, shift_args(a, n_full, Dims<KIND>::D, Dims<KIND>::A), pw);
          }
        } else {
          kern<<<wave(c.n, 0), kTile, 0, st>>>(c, a, pw);
        }
      } else {
        kern<<<g, kTile, 0, st>>>(c, a, pw);
      }
    }
    else if (c.fast) step_kernel<KIND, MSORT_RNG_PHILOX, LAYOUT_COMPACT, true><<<g, kTile, 0, st>>>(c, a, pw);
    else step_kernel<KIND, MSORT_RNG_PHILOX, LAYOUT_COMPACT, false><<<g, kTile, 0, st>>>(c, a, pw);
  } else {
    if (c.fast) step_kernel<KIND, MSORT_RNG_PHILOX, LAYOUT_WIDE, true><<<g, kTile, 0, st>>>(c, a, pw);
    else step_kernel<KIND, MSORT_RNG_PHILOX, LAYOUT_WIDE, false><<<g, kTile, 0, st>>>(c, a, pw);
  }
  return cudaGetLastError();
}

cudaError_t launch_step(const DevConfig& c, const StepLaunch& l, int rng, cudaStream_t st) {
  StepArgs a;
  a.state = (uint4*)l.state; a.actions = (const long long*)l.actions; a.obs = l.obs; a.reward = l.reward;
  a.terminated = l.terminated; a.mask = l.mask;
  const msort_info_out_t* f = l.info;
  a.info_action = f ? (long long*)f->action : nullptr;
  a.info_overflow = f ? f->overflow : nullptr;
  a.info_overflow_mat = f ? f->overflow_material : nullptr;
  a.info_sort_mode = f ? f->sort_mode : nullptr;
  a.info_press_action = f ? f->press_action : nullptr;
  a.info_invalid = f ? f->invalid_action : nullptr;
  a.info_sorted_true = f ? f->sorted_true : nullptr;
  a.info_r_sort = f ? f->reward_sort : nullptr;
  a.info_r_press = f ? f->reward_press : nullptr;
  a.any_step_info = a.info_action || a.info_overflow || a.info_overflow_mat || a.info_sort_mode ||
                    a.info_press_action || a.info_invalid || a.info_r_sort || a.info_r_press || a.info_sorted_true;
  a.act_tma = ((uintptr_t)l.actions & 15u) == 0;
  a.policy_tc = reinterpret_cast<const uint4*>(l.policy_tc);
  a.mode_scratch = l.mode_scratch;
  const FusedLaunch* fz = l.fused;
  a.fused_w = fz ? reinterpret_cast<const uint4*>(fz->packed) : nullptr;
  a.next_actions = fz ? (long long*)fz->next_actions : nullptr; a.next_logp = fz ? fz->next_logp : nullptr; a.next_value = fz ? fz->next_value : nullptr;
  a.draw_key0 = fz ? (unsigned)(fz->seed & 0xffffffffu) : 0u; a.draw_key1 = fz ? (unsigned)(fz->seed >> 32) : 0u;
  a.draw_t = fz ? fz->t : 0u; a.draw_t_dev = fz ? fz->t_dev : nullptr; a.deterministic = fz ? fz->deterministic : 0;
  a.terminal_obs = f ? f->terminal_obs : nullptr;
  a.episode_return = f ? f->episode_return : nullptr;
  a.episode_length = f ? f->episode_length : nullptr;
  a.stats = f ? f->stats : nullptr;
  const msort_replay_t* r = l.replay;
  a.noise_u = r ? r->noise_u : nullptr; a.redis_u = r ? r->redis_u : nullptr;
  a.redis_len = r ? r->redis_len : 0; a.input_counts = r ? r->input_counts : nullptr;
  a.press_choice = r ? r->press_choice : nullptr; a.sort_mode_in = r ? r->sort_mode : nullptr;
  switch (c.kind) {
    case MSORT_ENV_SORT: return launch_step_kind<MSORT_ENV_SORT>(c, a, l.policy_host, rng, st, l.variant, l.allow_hot, l.persist_per_sm);
    case MSORT_ENV_PRESS: return launch_step_kind<MSORT_ENV_PRESS>(c, a, l.policy_host, rng, st, l.variant, l.allow_hot, l.persist_per_sm);
    default: return launch_step_kind<MSORT_ENV_MONO>(c, a, l.policy_host, rng, st, l.variant, l.allow_hot, l.persist_per_sm);
  }
}

cudaError_t launch_rollout_policy(const DevConfig& c, const float* obs, const uint8_t* mask, const uint32_t* packed, uint64_t seed,
                                  uint32_t t, const uint32_t* t_dev, int deterministic, int64_t* actions, float* logp, float* value,
                                  cudaStream_t st) {
  if (c.n <= 0) return cudaSuccess;
  const PolicyOut po{(long long*)actions, logp, value, (unsigned)(seed & 0xffffffffu), (unsigned)(seed >> 32), t, t_dev, deterministic};
  const unsigned grid = std::min(tiles(c.n), (unsigned)(8 * c.sm_count));     // one wave of resident CTAs (26.6 KB, 64 registers: 8 per SM)
  rollout_policy_kernel<<<grid, kTile, 0, st>>>(obs, mask, reinterpret_cast<const uint4*>(packed), c.n, c.gid0, po);
  return cudaGetLastError();
}

cudaError_t launch_pack_fused(const float* params, uint32_t* packed, cudaStream_t st) {
  pack_fused_kernel<<<(kFwWords + 255) / 256, 256, 0, st>>>(params, packed);
  return cudaGetLastError();
}

cudaError_t launch_tc_logits(const float* obs13, const uint32_t* tcw, long long n, float* logits, int sm_count, cudaStream_t st) {
  if (n <= 0) return cudaSuccess;
  #ifndef MSORT_EXP_TC_CTAS
#define MSORT_EXP_TC_CTAS 4
#endif
  const unsigned grid = (unsigned)std::min<long long>((n + kTile - 1) / kTile, (long long)MSORT_EXP_TC_CTAS * sm_count);
  tc_logits_kernel<<<grid, kTile, 0, st>>>(obs13, reinterpret_cast<const uint4*>(tcw), n, logits);
  return cudaGetLastError();
}

cudaError_t launch_reset(const DevConfig& c, void* state, const uint8_t* which, const uint8_t* first_pattern,
                         float* obs, uint8_t* mask, uint32_t reset_flags, cudaStream_t st) {
  switch (c.kind) {
    case MSORT_ENV_SORT: reset_kernel<MSORT_ENV_SORT><<<tiles(c.n), kTile, 0, st>>>(c, (uint4*)state, which, first_pattern, obs, mask, reset_flags); break;
    case MSORT_ENV_PRESS: reset_kernel<MSORT_ENV_PRESS><<<tiles(c.n), kTile, 0, st>>>(c, (uint4*)state, which, first_pattern, obs, mask, reset_flags); break;
    default: reset_kernel<MSORT_ENV_MONO><<<tiles(c.n), kTile, 0, st>>>(c, (uint4*)state, which, first_pattern, obs, mask, reset_flags); break;
  }
  return cudaGetLastError();
}

cudaError_t launch_observe(const DevConfig& c, const void* state, float* obs, uint8_t* mask, int after_shift, cudaStream_t st) {
  switch (c.kind) {
    case MSORT_ENV_SORT: observe_kernel<MSORT_ENV_SORT><<<tiles(c.n), kTile, 0, st>>>(c, (const uint4*)state, obs, mask, after_shift); break;
    case MSORT_ENV_PRESS: observe_kernel<MSORT_ENV_PRESS><<<tiles(c.n), kTile, 0, st>>>(c, (const uint4*)state, obs, mask, after_shift); break;
    default: observe_kernel<MSORT_ENV_MONO><<<tiles(c.n), kTile, 0, st>>>(c, (const uint4*)state, obs, mask, after_shift); break;
  }
  return cudaGetLastError();
}

cudaError_t launch_sample(const DevConfig& c, const uint8_t* mask, int64_t* actions, uint64_t seed, uint32_t t,
                          cudaStream_t st) {
  unsigned k0 = (unsigned)(seed & 0xffffffffu), k1 = (unsigned)(seed >> 32);
  switch (c.kind) {
    case MSORT_ENV_SORT: sample_kernel<2><<<tiles(c.n), kTile, 0, st>>>(c, mask, (long long*)actions, k0, k1, t); break;
    case MSORT_ENV_PRESS: sample_kernel<11><<<tiles(c.n), kTile, 0, st>>>(c, mask, (long long*)actions, k0, k1, t); break;
    default: sample_kernel<22><<<tiles(c.n), kTile, 0, st>>>(c, mask, (long long*)actions, k0, k1, t); break;
  }
  return cudaGetLastError();
}

cudaError_t launch_generate_streams(const DevConfig& c, uint32_t episode, uint32_t first_step, uint32_t num_steps, uint32_t* input_counts,
                                    double* noise_u, uint32_t* draw_words, uint8_t* first_pattern, cudaStream_t st) {
  generate_streams_kernel<<<(unsigned)((c.n + 255) / 256), 256, 0, st>>>(c, episode, first_step, num_steps, input_counts, noise_u,
                                                                         draw_words, first_pattern);
  return cudaGetLastError();
}

cudaError_t launch_widen_actions(const uint8_t* in, int64_t* out, long long n, cudaStream_t st) {
  if (n <= 0) return cudaSuccess;
  widen_actions_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(in, (long long*)out, n);
  return cudaGetLastError();
}

cudaError_t launch_pack_flags(int kind, const uint8_t* mask, const uint8_t* terminated, uint16_t* flags, long long n, cudaStream_t st) {
  if (n <= 0) return cudaSuccess;
  switch (kind) {
    case MSORT_ENV_SORT: pack_flags_kernel<2><<<tiles(n), kTile, 0, st>>>(mask, terminated, flags, n); break;
    case MSORT_ENV_PRESS: pack_flags_kernel<11><<<tiles(n), kTile, 0, st>>>(mask, terminated, flags, n); break;
    default: pack_flags_kernel<22><<<tiles(n), kTile, 0, st>>>(mask, terminated, flags, n); break;
  }
  return cudaGetLastError();
}

cudaError_t launch_rule_actions(const DevConfig& c, const void* state, int after_shift, int64_t* actions, cudaStream_t st) {
  rule_actions_kernel<<<tiles(c.n), kTile, 0, st>>>(c, (const uint4*)state, after_shift, (long long*)actions);
  return cudaGetLastError();
}

cudaError_t launch_export(const DevConfig& c, const void* state, msort_env_state_t* out, cudaStream_t st) {
  export_kernel<<<tiles(c.n), kTile, 0, st>>>(c, (const uint4*)state, out);
  return cudaGetLastError();
}

cudaError_t launch_gather(const DevConfig& c, const void* state, const int64_t* env_ids, long long count,
                          msort_env_state_t* out, cudaStream_t st) {
  if (count <= 0) return cudaSuccess;
  gather_kernel<<<tiles(count), kTile, 0, st>>>(c, (const uint4*)state, (const long long*)env_ids, count, out);
  return cudaGetLastError();
}

cudaError_t launch_import(const DevConfig& c, void* state, const msort_env_state_t* in, cudaStream_t st) {
  import_kernel<<<tiles(c.n), kTile, 0, st>>>(c, (uint4*)state, in);
  return cudaGetLastError();
}

cudaError_t launch_stats(const DevConfig& c, const void* state, double* out16, int sm_count, cudaStream_t st) {
  cudaError_t e = cudaMemsetAsync(out16, 0, sizeof(double) * MSORT_NUM_STATS, st);
  if (e != cudaSuccess) return e;
  long long want = (c.n + 255) / 256;
  unsigned grid = (unsigned)(want < (long long)sm_count * 8 ? want : (long long)sm_count * 8);
  if (grid == 0) grid = 1;
  stats_kernel<<<grid, 256, 0, st>>>(c, (const uint4*)state, out16);
  return cudaGetLastError();
}

}  // namespace msort
