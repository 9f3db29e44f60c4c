// msort_policy.cu — fused actor-critic inference + masked categorical sampling for the GPU-resident
// MaskablePPO rollout (ref: MaskablePPO(policy_kwargs=dict(net_arch=dict(pi=[32,32],
// vf=[32,32]))) training.py:115-131 — two tanh towers obs -> 32 -> 32 -> {A logits | 1 value}; logits
// masked with -1e8 like sb3_contrib, env_monolith.py:152-158 for the masked draw).
//
// Unlike Env_2's embedded 13->32->32->2 policy (argmax parity with an fp32 reference: CUDA cores), this
// IS a dense contraction with no bit-parity requirement — M = number of envs (1e6), three layers — so it
// runs on the 5th-generation tensor cores:
//   * one CTA = one tile of 128 envs = one UMMA M=128 tile; the two towers are evaluated as ONE network
//     with concatenated / block-diagonal weights:  [128x32]·[32x64] -> tanh -> [128x64]·[64x64] -> tanh
//     -> [128x64]·[64x32]  (columns 0..A-1 logits, column A value);
//   * tcgen05.mma.cta_group::1.kind::f16 issued by one thread, A (activations) and B (weights) as fp16 in
//     shared memory in the canonical no-swizzle K-major layout (fp16 keeps the 10 mantissa bits tf32 would;
//     every operand is O(1)), fp32 accumulators in 64 TMEM columns;
//   * TMEM lane = tile row = env, so after tcgen05.ld (32x32b) thread r holds row r's outputs: bias + tanh
//     in registers, written straight back as the next layer's A operand (16-byte chunk kc = 8 columns of row r at
//     (kc*128 + r)*16 B: consecutive threads -> consecutive 16 B, conflict-free STS.128); two threads per
//     env, one per tower (accumulator columns 0..31 / 32..63), 256 threads per CTA;
//   * tcgen05.commit -> mbarrier tells the CTA when a layer's accumulators are complete;
//   * persistent CTAs (4 per SM): the 17 KB of packed weights are staged into shared memory once; the obs
//     and mask tiles (contiguous 14.8 KB / 2.8 KB) arrive by TMA bulk copies (cp.async.bulk + mbarrier
//     complete_tx); the next tile's copies are issued as soon as layer 1 has consumed the obs tile and
//     land behind the three layers of the current one.
// Epilogue per env: masked log-softmax over the A logits, inverse-CDF draw with one Philox uniform keyed by
// (seed, t, global env id) (or argmax), outputs action / log-prob / value with coalesced stores.
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <algorithm>

#include "msort_device.cuh"
#include "msort_launch.h"
#include "msort_umma.cuh"

namespace msort {

namespace {

constexpr int kRows = 128;                 // envs per tile == UMMA M
constexpr int kThreads = 256;              // two threads per env: one per tower
constexpr int kK1 = 32, kN1 = 64;          // layer 1: obs (padded to 32) -> pi hidden | vf hidden
constexpr int kK2 = 64, kN2 = 64;          // layer 2: block-diagonal
constexpr int kK3 = 64, kN3 = 32;          // layer 3: A logits | value | padding
// packed buffer (float32 words): fp16 weights B1 | B2 | B3 (two per word), then the fp32 biases
constexpr int kB1 = 0, kB2 = kB1 + kK1 * kN1, kB3 = kB2 + kK2 * kN2;       // offsets in fp16 elements
constexpr int kWeightHalves = kB3 + kK3 * kN3;
constexpr int kPacked = kWeightHalves / 2 + kN1 + kN2 + kN3;               // == MSORT_POLICY_ACT_WEIGHTS
static_assert(kPacked == MSORT_POLICY_ACT_WEIGHTS, "packed actor-critic size");
constexpr int kTmemCols = 64;


using umma::fence_after_sync;
using umma::fence_async_proxy;
using umma::fence_before_sync;
using umma::pack8;
using umma::tmem_ld32;

template <int D, int A>
struct __align__(16) PolicySmem {
  __half b[kWeightHalves];         // B1 | B2 | B3 in canonical K-major order (16 KB)
  float bias[kN1 + kN2 + kN3];
  uint4 a[kK2 / 8 * kRows];        // A operand: 8 chunks (8 fp16 columns each) x 128 rows x 16 B (16 KB; layer 1 uses 4)
  float stage[kRows * D];          // obs tile (row-major), filled by a TMA bulk copy; free again once layer 1's A operand is built
  uint8_t mask[2][kRows * A];      // mask tile ring (needed until the end of the tile)
  uint64_t bar;                    // MMA completion (tcgen05.commit)
  uint64_t full_obs;               // obs tile landed (TMA complete_tx)
  uint64_t full_mask[2];           // mask tile landed
  uint32_t tmem_base;
};

// one layer: all threads have written their A chunks; thread 0 issues the K-steps and commits
template <int K, int N, class Smem>
__device__ __forceinline__ void issue_layer(const Smem& sm, int b_off, uint32_t tmem_d, uint64_t* bar) {
  const uint32_t a0 = smem_u32(sm.a), b0 = smem_u32(sm.b + b_off);
  // the K-steps' descriptors differ only in the 14-bit start-address field (bytes >> 4; never carries out of the field in
  // < 256 KB of shared memory): one base word per operand, then an immediate add per MMA
  const uint64_t abase = umma::smem_desc(a0, kRows * 16u, 128u), bbase = umma::smem_desc(b0, N * 16u, 128u);
#pragma unroll
  for (int s = 0; s < K / 16; ++s) {   // one instruction = K 16 (fp16) = two 16-byte chunks
    const uint64_t ad = abase + (uint64_t)(((uint32_t)(2 * s) * kRows * 16u) >> 4);
    const uint64_t bd = bbase + (uint64_t)(((uint32_t)(2 * s) * N * 16u) >> 4);
    umma::mma_f16(tmem_d, ad, bd, umma::idesc_f16(N), s > 0 ? 1u : 0u);
  }
  umma::commit(bar);
}

}  // namespace

// tanh for the hidden layers.  Default: the hardware tanh (MUFU.TANH, |error| <~ 5e-4 — the same size as the
// tf32 rounding of the operands); -DMSORT_POLICY_TANH_APPROX=0 selects 1 - 2/(exp2(c*x)+1) (error ~1e-7).
#ifndef MSORT_POLICY_TANH_APPROX
#define MSORT_POLICY_TANH_APPROX 1
#endif
__device__ __forceinline__ float policy_tanh(float x) {
#if MSORT_POLICY_TANH_APPROX
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
#else
  float e, r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(x * 2.8853900817779268f));   // exp(2x)
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(e + 1.0f));
  return fmaf(-2.0f, r, 1.0f);
#endif
}

// 256 threads: thread (row, half) — row = tid & 127 is the env / TMEM lane, half = tid >> 7 the tower:
// warps 0-3 own accumulator columns 0..31 (policy tower), warps 4-7 columns 32..63 (value tower); both
// warp groups reach the same TMEM lane quarter (warp % 4).
template <int D, int A>
__global__ void __launch_bounds__(kThreads, 4)
policy_act_kernel(const float* __restrict__ obs, const uint8_t* __restrict__ mask, const float* __restrict__ packed,
                  long long n, long long gid0, unsigned key0, unsigned key1, unsigned t_in, const unsigned* __restrict__ t_dev,
                  int deterministic, int use_tma, long long ld_obs, long long ld_mask,
                  long long* __restrict__ actions, float* __restrict__ logp_out, float* __restrict__ value_out) {
  static_assert(D <= 32 && A <= 31, "padded layer sizes");
  extern __shared__ __align__(16) unsigned char smem_raw[];
  using Smem = PolicySmem<D, A>;
  Smem& sm = *reinterpret_cast<Smem*>(smem_raw);
  const int tid = threadIdx.x, warp = tid >> 5, row = tid & (kRows - 1), half = tid >> 7;
  const unsigned t = t_in + (t_dev ? *t_dev : 0u);   // draw index: a graph-replayed rollout bumps *t_dev between replays

  // ---- one-time setup: TMEM columns, mbarriers, weights
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(&sm.tmem_base)), "r"(kTmemCols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (tid == 0) {
    mbar_init(&sm.bar, 1);
    mbar_init(&sm.full_obs, 1);
    mbar_init(&sm.full_mask[0], 1);
    mbar_init(&sm.full_mask[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  {
    const float4* src = reinterpret_cast<const float4*>(packed);
    float4* dst = reinterpret_cast<float4*>(sm.b);       // b[] and bias[] are contiguous
    for (int e = tid; e < kPacked / 4; e += kThreads) dst[e] = src[e];
  }
  fence_async_proxy();            // weights were written through the generic proxy, the MMA reads through the async one
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem = sm.tmem_base;
  const uint32_t tcol = tmem + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(32 * half);   // this thread's lane, its tower's columns
  uint32_t phase = 0;

  const long long ntiles = (n + kRows - 1) / kRows;
  constexpr uint32_t obs_bytes = (uint32_t)(kRows * D) * 4u, mask_bytes = (uint32_t)(kRows * A);
  static_assert(obs_bytes % 16 == 0 && mask_bytes % 16 == 0, "TMA bulk copies move multiples of 16 bytes");
  // a full tile is one contiguous, 16-byte-aligned range of each tensor: one TMA bulk copy each
  auto prefetch = [&](long long tl, int slot) {
    if (use_tma && tl < ntiles && (tl + 1) * kRows <= n) {
      mbar_expect_tx(&sm.full_obs, obs_bytes);
      bulk_load(sm.stage, obs + tl * kRows * D, obs_bytes, &sm.full_obs);
      mbar_expect_tx(&sm.full_mask[slot], mask_bytes);
      bulk_load(sm.mask[slot], mask + tl * kRows * A, mask_bytes, &sm.full_mask[slot]);
    }
  };
  // plain-load fallback for a tile TMA cannot fetch (ragged last tile: its byte count need not be a multiple
  // of 16; or unaligned tensors), executed by `nthr` threads numbered `t0`
  auto plain_load = [&](long long tl, int slot, int t0, int nthr) {
    // (also the path of msort_policy_eval: rows `ld_obs` floats / `ld_mask` bytes apart — e.g. the 13-wide sort part of a
    //  29-wide observation — and no mask at all = every action valid)
    const long long r0 = tl * kRows;
    const int rws = (int)min((long long)kRows, n - r0);
    for (int e = t0; e < kRows * D; e += nthr) {
      const int r = e / D, k = e - r * D;
      sm.stage[e] = r < rws ? obs[(r0 + r) * ld_obs + k] : 0.f;
    }
    for (int e = t0; e < kRows * A; e += nthr) {
      const int r = e / A, k = e - r * A;
      sm.mask[slot][e] = r < rws ? (mask ? mask[(r0 + r) * ld_mask + k] : (uint8_t)1) : (uint8_t)0;
    }
  };
  // layer-1 A operand of this thread's row: obs zero-padded to 32 columns, chunks [kc0, kc0 + nkc)
  auto build_a1 = [&](int kc0, int nkc) {
    const float* x = sm.stage + row * D;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      if (q < nkc) {
        const int kc = kc0 + q;
        float v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = 8 * kc + j < D ? x[8 * kc + j] : 0.f;
        sm.a[kc * kRows + row] = pack8(v);
      }
    }
  };
  auto tile_by_tma = [&](long long tl) { return use_tma && (tl + 1) * kRows <= n; };

  // ---- prologue: the first tile's layer-1 operand (afterwards the value-tower warps build the next tile's
  //      operand while the policy-tower warps run the softmax epilogue)
  if (tid == 0) prefetch(blockIdx.x, 0);
  if ((long long)blockIdx.x < ntiles) {
    if (tile_by_tma(blockIdx.x)) mbar_wait(&sm.full_obs, 0u);
    else { plain_load(blockIdx.x, 0, tid, kThreads); __syncthreads(); }
    build_a1(2 * half, 2);
  }
  uint32_t it = 0;
  for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
    const long long row0 = tile * kRows;
    const int rows = (int)min((long long)kRows, n - row0);
    const int slot = (int)(it & 1u);
    const uint8_t* mtile = sm.mask[slot];
    const bool by_tma = tile_by_tma(tile);
    fence_async_proxy();
    fence_before_sync();
    __syncthreads();                     // this tile's layer-1 operand is complete; the previous tile is finished
    if (tid == 0) {
      fence_after_sync();
      issue_layer<kK1, kN1>(sm, kB1, tmem, &sm.bar);
      // the obs tile has been consumed (the A chunks were built before the barrier) and the other mask slot was
      // released when the previous tile ended: fetch the next tile behind the three layers
      prefetch(tile + gridDim.x, slot ^ 1);
    }
    mbar_wait(&sm.bar, phase); phase ^= 1;
    fence_after_sync();

    // ---- hidden layers: this tower's 32 accumulators -> +bias -> tanh -> its 8 chunks of the next A operand
#pragma unroll
    for (int layer = 0; layer < 2; ++layer) {
      const float* bias = sm.bias + (layer == 0 ? 0 : kN1) + 32 * half;
      float v[32];
      tmem_ld32(tcol, v);
      fence_before_sync();               // this thread's TMEM reads are done: the next MMA may overwrite the columns
#pragma unroll
      for (int kc = 0; kc < 4; ++kc) {
        float h[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) h[j] = policy_tanh(v[8 * kc + j] + bias[8 * kc + j]);
        sm.a[(4 * half + kc) * kRows + row] = pack8(h);
      }
      fence_async_proxy();
      __syncthreads();
      if (tid == 0) {
        fence_after_sync();
        if (layer == 0) issue_layer<kK2, kN2>(sm, kB2, tmem, &sm.bar);
        else issue_layer<kK3, kN3>(sm, kB3, tmem, &sm.bar);
      }
      mbar_wait(&sm.bar, phase); phase ^= 1;
      fence_after_sync();
    }

    // ---- output layer (columns 0..A-1 logits, column A value): the policy-tower threads finish the env
    if (half == 0) {
      float o[32];
      tmem_ld32(tcol, o);
      if (by_tma) mbar_wait(&sm.full_mask[slot], (it >> 1) & 1u);
      const long long i = row0 + row;
      if (row < rows) {
        const float* b3 = sm.bias + kN1 + kN2;
        const uint8_t* mrow = mtile + row * A;
        float mx = -3.0e38f;
#pragma unroll
        for (int a = 0; a < A; ++a) {
          o[a] = mrow[a] ? o[a] + b3[a] : -1e8f;        // sb3_contrib masks logits with -1e8
          mx = fmaxf(mx, o[a]);
        }
        float sum = 0.f;
#pragma unroll
        for (int a = 0; a < A; ++a) { o[a] = __expf(o[a] - mx); sum += o[a]; }   // o[a]: unnormalised probability
        int act = 0;
        if (deterministic) {
          float best = -1.f;
#pragma unroll
          for (int a = 0; a < A; ++a) if (o[a] > best) { best = o[a]; act = a; }
        } else {
          const unsigned long long g = (unsigned long long)(gid0 + i);
          const U4 r4 = philox4x32_10((uint32_t)g, (uint32_t)(g >> 32), 0xAC70u, t, key0, key1);
          const float u = ((float)(r4.x >> 8) + 0.5f) * (1.0f / 16777216.0f) * sum;   // uniform in (0, sum)
          float c = 0.f;
          int last = 0;
          bool found = false;
#pragma unroll
          for (int a = 0; a < A; ++a) {
            c += o[a];
            if (o[a] > 0.f) last = a;
            if (!found && u < c && o[a] > 0.f) { act = a; found = true; }
          }
          if (!found) act = last;
        }
        float pa = 0.f;
#pragma unroll
        for (int a = 0; a < A; ++a) if (a == act) pa = o[a];
        actions[i] = act;
        logp_out[i] = __logf(pa) - __logf(sum);
        value_out[i] = o[A] + b3[A];
      }
    }
    else {
      // value-tower warps: layer 3 has consumed the A buffer, so build the NEXT tile's layer-1 operand now
      const long long next = tile + gridDim.x;
      if (next < ntiles) {
        if (tile_by_tma(next)) mbar_wait(&sm.full_obs, (it + 1) & 1u);
        else {
          plain_load(next, slot ^ 1, row, kRows);
          asm volatile("bar.sync 1, %0;" :: "n"(kRows) : "memory");   // the 128 value-tower threads only
        }
        build_a1(0, 4);
      }
    }
    // (the barrier at the top of the next iteration ends this tile)
  }

  fence_before_sync();
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "r"(kTmemCols) : "memory");
  }
}

template <int D, int A>
static cudaError_t launch_policy_act_da(const DevConfig& c, const float* obs, const uint8_t* mask, const float* packed,
                                        uint64_t seed, uint32_t t, const uint32_t* t_dev, int deterministic, int64_t* actions, float* logp,
                                        float* value, int sm_count, cudaStream_t st, long long ld_obs = D, long long ld_mask = A) {
  static_assert(sizeof(PolicySmem<D, A>) <= 55 * 1024, "four CTAs per SM");
  const size_t smem = sizeof(PolicySmem<D, A>);   // opt-in size set once per device by prepare_policy_kernels (msort_create)
  const long long ntiles = (c.n + kRows - 1) / kRows;
  const unsigned grid = (unsigned)std::min<long long>(ntiles, 4ll * sm_count);
  // TMA bulk copies need 16-byte aligned tile addresses (tile sizes are multiples of 16 bytes)
  const int use_tma = mask && ld_obs == D && ld_mask == A &&
                      ((reinterpret_cast<uintptr_t>(obs) | reinterpret_cast<uintptr_t>(mask)) & 15u) == 0;
  policy_act_kernel<D, A><<<grid, kThreads, smem, st>>>(obs, mask, packed, c.n, c.gid0, (unsigned)(seed & 0xffffffffu),
                                                        (unsigned)(seed >> 32), t, t_dev, deterministic, use_tma, ld_obs, ld_mask,
                                                        (long long*)actions, logp, value);
  return cudaGetLastError();
}

// once per handle, on the handle's device (msort_create): the kernels' dynamic shared memory exceeds the 48 KB default
cudaError_t prepare_policy_kernels() {
  cudaError_t e = cudaFuncSetAttribute(policy_act_kernel<29, 22>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(PolicySmem<29, 22>));
  if (e == cudaSuccess) e = cudaFuncSetAttribute(policy_act_kernel<16, 11>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(PolicySmem<16, 11>));
  if (e == cudaSuccess) e = cudaFuncSetAttribute(policy_act_kernel<13, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(PolicySmem<13, 2>));
  return e;
}

cudaError_t launch_policy_act(const DevConfig& c, const float* obs, const uint8_t* mask, const float* packed,
                              int D, int A, uint64_t seed, uint32_t t, const uint32_t* t_dev, int deterministic, int64_t* actions,
                              float* logp, float* value, int sm_count, cudaStream_t st) {
  if (D == 29 && A == 22) return launch_policy_act_da<29, 22>(c, obs, mask, packed, seed, t, t_dev, deterministic, actions, logp, value, sm_count, st);
  if (D == 16 && A == 11) return launch_policy_act_da<16, 11>(c, obs, mask, packed, seed, t, t_dev, deterministic, actions, logp, value, sm_count, st);
  if (D == 13 && A == 2) return launch_policy_act_da<13, 2>(c, obs, mask, packed, seed, t, t_dev, deterministic, actions, logp, value, sm_count, st);
  return cudaErrorInvalidValue;
}

// msort_policy_eval: the same kernels on caller-given rows (any env kind's handle; strided observations, optional mask)
cudaError_t launch_policy_eval(const DevConfig& c, int D, int A, long long rows, const float* obs, long long ld_obs, const uint8_t* mask,
                               long long ld_mask, const float* packed, uint64_t seed, uint32_t t, int deterministic, int64_t* actions,
                               float* logp, float* value, int sm_count, cudaStream_t st) {
  DevConfig d = c;
  d.n = rows;
  if (rows <= 0) return cudaSuccess;
  if (D == 29 && A == 22) return launch_policy_act_da<29, 22>(d, obs, mask, packed, seed, t, nullptr, deterministic, actions, logp, value, sm_count, st, ld_obs, ld_mask);
  if (D == 16 && A == 11) return launch_policy_act_da<16, 11>(d, obs, mask, packed, seed, t, nullptr, deterministic, actions, logp, value, sm_count, st, ld_obs, ld_mask);
  if (D == 13 && A == 2) return launch_policy_act_da<13, 2>(d, obs, mask, packed, seed, t, nullptr, deterministic, actions, logp, value, sm_count, st, ld_obs, ld_mask);
  return cudaErrorInvalidValue;
}

}  // namespace msort
