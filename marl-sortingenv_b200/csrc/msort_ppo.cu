// msort_ppo.cu — the UPDATE half of the GPU-resident MaskablePPO loop as hand-written kernels:
// generalised advantage estimation, fused actor-critic forward + loss + backward per minibatch, and a fused
// global-norm-clip + Adam step.  ref: sb3_contrib.MaskablePPO as the reference configures it (training.py:115-131:
// net_arch=dict(pi=[32,32], vf=[32,32]), tanh, ent_coef=0.05; SB3 defaults n_epochs=10, clip 0.2, gamma 0.99,
// lambda 0.95, lr 3e-4, vf_coef 0.5, max_grad_norm 0.5, Adam eps 1e-5, advantages normalised per minibatch).
//
// The networks are tiny (4 791 parameters at D=29, A=22) and a minibatch is thousands of rows, so the work is
// row-parallel fp32 on the CUDA cores (bit-comparable with the PyTorch fp32 reference the tests check against):
//   * one CTA = 128 rows = 128 threads; both towers' weights live in shared memory in input-major order, so a thread
//     advances 4 output units per LDS.128 (all threads read the same address: a broadcast) — 1 load per 4 FMAs;
//   * per-row activations / pre-activation gradients go to shared memory in rows of 36 floats (16-byte aligned for the
//     weight-gradient phase, conflict-free for a thread reading its own row with LDS.128); five such buffers (the layer-1
//     gradient rows reuse the layer-3 buffer after its weight gradients are taken) + the weights = 103 KB: two CTAs per SM;
//   * weight gradients dW[o][i] = sum_r dz[r][o] * act[r][i]: thread (o, i-block of 8) loops over the tile's rows
//     (dz: one conflict-free load; act: two broadcast LDS.128), then ONE atomicAdd per parameter per CTA.
// Parameter vector (flat fp32, owned by the caller):
//   pi: W1[32][D] b1[32] W2[32][32] b2[32] W3[A][32] b3[A]   |   vf: W1[32][D] b1[32] W2[32][32] b2[32] W3[1][32] b3[1]
#include <cuda_runtime.h>
#include <stdint.h>

#include <algorithm>

#include "../../include/msort.h"
#include "msort_launch.h"

namespace msort {

namespace {

#ifndef MSORT_EXP_PPO_NOATOMIC
#define MSORT_EXP_PPO_NOATOMIC 0
#endif
constexpr int kRows = 128;     // rows per CTA == threads per CTA
constexpr int kH = 32;         // hidden width of every tower layer
constexpr int kLd = 36;        // row stride of the per-row shared buffers

struct PpoLayout {             // offsets into the flat parameter vector
  int pi_w1, pi_b1, pi_w2, pi_b2, pi_w3, pi_b3, vf_w1, vf_b1, vf_w2, vf_b2, vf_w3, vf_b3, total;
};
__host__ __device__ inline PpoLayout ppo_layout(int D, int A) {
  PpoLayout L;
  int o = 0;
  L.pi_w1 = o; o += kH * D; L.pi_b1 = o; o += kH; L.pi_w2 = o; o += kH * kH; L.pi_b2 = o; o += kH;
  L.pi_w3 = o; o += A * kH; L.pi_b3 = o; o += A;
  L.vf_w1 = o; o += kH * D; L.vf_b1 = o; o += kH; L.vf_w2 = o; o += kH * kH; L.vf_b2 = o; o += kH;
  L.vf_w3 = o; o += kH; L.vf_b3 = o; o += 1;
  L.total = o;
  return L;
}

struct PpoArgs {
  const float* params;
  float* grads;                // GRAD: accumulated with atomics (caller zeroes / the Adam kernel zeroes after use)
  const float* obs; const uint8_t* mask; const long long* actions;
  const float* old_logp; const float* adv; const float* ret;
  const long long* idx;        // minibatch row indices (nullptr: rows first .. first + count)
  long long first, count;
  const double* adv_sums;      // {sum a, sum a^2} of the minibatch's advantages (adv_sums_kernel), or nullptr (no normalisation)
  const float* image;          // GRAD: both towers' weights already transposed / padded in the kernel's shared-memory order
                               // (ppo_image_kernel, 2 * PpoSmem::kImage floats) or nullptr (each CTA gathers them from `params`)
  float clip, vf_coef, ent_coef, inv_count;
  float* stats;                // GRAD: += {policy-gradient loss, value loss, entropy, clipped fraction, rows}
  float* logp_out; float* value_out;   // !GRAD: per-row outputs at the row's own index
};

__device__ __forceinline__ float tanh_exact(float x) {   // 1 - 2 / (exp(2x) + 1): ~1e-7 absolute, like torch's fp32 tanh
  const float e = __expf(2.0f * x);
  return 1.0f - __fdividef(2.0f, e + 1.0f);
}

// y[0..O) = b + W x: W^T in shared memory ([I][OP] floats, OP = O rounded up to 4, padding zero), x = this thread's row of a
// per-row shared buffer.  The input loop is NOT fully unrolled on purpose: unrolled, ptxas front-loads the layer's 256
// weight loads and spills kilobytes; per input it is 1 LDS (x_i) + OP/4 broadcast LDS.128 + OP FMAs.
template <int I, int O, int OP>
__device__ __forceinline__ void dense_fwd(const float* __restrict__ wt, const float* __restrict__ b, const float* __restrict__ xrow, float (&y)[OP]) {
#pragma unroll
  for (int o = 0; o < OP; ++o) y[o] = o < O ? b[o] : 0.f;
#pragma unroll 2
  for (int i = 0; i < I; ++i) {
    const float xi = xrow[i];
#pragma unroll
    for (int o4 = 0; o4 < OP / 4; ++o4) {
      const float4 w = *reinterpret_cast<const float4*>(wt + i * OP + 4 * o4);
      y[4 * o4] = fmaf(w.x, xi, y[4 * o4]); y[4 * o4 + 1] = fmaf(w.y, xi, y[4 * o4 + 1]);
      y[4 * o4 + 2] = fmaf(w.z, xi, y[4 * o4 + 2]); y[4 * o4 + 3] = fmaf(w.w, xi, y[4 * o4 + 3]);
    }
  }
}

// Backward through one dense + tanh pair: dzprev[i] = (sum_o W[o][i] dz[o]) * (1 - h[i]^2), with h and dzprev this thread's rows
// of per-row shared buffers (same W^T layout: row i holds the weights of input i towards every output).
template <int I, int OP>
__device__ __forceinline__ void dense_bwd_tanh(const float* __restrict__ wt, const float (&dz)[OP], const float* __restrict__ hrow,
                                               float* __restrict__ dzprev_row) {
#pragma unroll 2
  for (int i = 0; i < I; ++i) {
    float a = 0.f;
#pragma unroll
    for (int o4 = 0; o4 < OP / 4; ++o4) {
      const float4 w = *reinterpret_cast<const float4*>(wt + i * OP + 4 * o4);
      a = fmaf(w.x, dz[4 * o4], a); a = fmaf(w.y, dz[4 * o4 + 1], a); a = fmaf(w.z, dz[4 * o4 + 2], a); a = fmaf(w.w, dz[4 * o4 + 3], a);
    }
    const float h = hrow[i];
    dzprev_row[i] = a * (1.f - h * h);
  }
}

template <int N>
__device__ __forceinline__ void row_store(float* __restrict__ buf, int r, const float (&v)[N]) {   // N a multiple of 4, N <= kLd
#pragma unroll
  for (int k = 0; k < N / 4; ++k) *reinterpret_cast<float4*>(buf + r * kLd + 4 * k) = make_float4(v[4 * k], v[4 * k + 1], v[4 * k + 2], v[4 * k + 3]);
}
template <int N>
__device__ __forceinline__ void row_load(const float* __restrict__ buf, int r, float (&v)[N]) {
#pragma unroll
  for (int k = 0; k < N / 4; ++k) {
    const float4 q = *reinterpret_cast<const float4*>(buf + r * kLd + 4 * k);
    v[4 * k] = q.x; v[4 * k + 1] = q.y; v[4 * k + 2] = q.z; v[4 * k + 3] = q.w;
  }
}

// dW[o][i] (+ db[o]) of one layer for the CTA's rows: thread t -> output o = t & 31, inputs [8 * (t >> 5), +8).
// dz / act are the per-row shared buffers (rows beyond `rows` hold zeros in dz).  The CTA's sums leave through a small shared
// staging tile ([input][33] floats: written with lanes = outputs, read with lanes = consecutive parameters, both conflict-free)
// so that a warp's atomicAdds hit 32 CONSECUTIVE parameters — 4 sector requests at L2 instead of 32 (the scattered form cost
// 14-34 % of the kernel).  Contains two CTA barriers: every thread calls it.
constexpr int kGs = 33;        // pitch of the staging tile
template <int I, int O>
__device__ __forceinline__ void weight_grad(const float* __restrict__ dz, const float* __restrict__ act, int rows, float* __restrict__ gw,
                                            float* __restrict__ gb, float* __restrict__ gstage, int tid) {
  const int o = tid & 31, i0 = 8 * (tid >> 5);
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f}, accb = 0.f;
  for (int r = 0; r < rows; ++r) {
    const float d = dz[r * kLd + o];
    const float4 a0 = *reinterpret_cast<const float4*>(act + r * kLd + i0), a1 = *reinterpret_cast<const float4*>(act + r * kLd + i0 + 4);
    acc[0] = fmaf(d, a0.x, acc[0]); acc[1] = fmaf(d, a0.y, acc[1]); acc[2] = fmaf(d, a0.z, acc[2]); acc[3] = fmaf(d, a0.w, acc[3]);
    acc[4] = fmaf(d, a1.x, acc[4]); acc[5] = fmaf(d, a1.y, acc[5]); acc[6] = fmaf(d, a1.z, acc[6]); acc[7] = fmaf(d, a1.w, acc[7]);
    accb += d;
  }
#if MSORT_EXP_PPO_NOATOMIC      // timing experiment: no global traffic for the weight gradients at all (only a NaN would be stored)
  if (o < O) {
#pragma unroll
    for (int k = 0; k < 8; ++k)
      if (i0 + k < I && acc[k] != acc[k]) gw[o * I + i0 + k] = acc[k];
    if (i0 == 0 && accb != accb) gb[o] = accb;
  }
#else
#pragma unroll
  for (int k = 0; k < 8; ++k) gstage[(i0 + k) * kGs + o] = acc[k];      // [input][output]
  if (i0 == 0) gstage[32 * kGs + o] = accb;                              // row 32: the bias sums
  __syncthreads();
  for (int e = tid; e < O * I; e += kRows) {                             // parameter e = o * I + i, consecutive across the lanes
    const int po = e / I, pi = e - po * I;
    atomicAdd(&gw[e], gstage[pi * kGs + po]);
  }
  if (tid < O) atomicAdd(&gb[tid], gstage[32 * kGs + tid]);
  __syncthreads();                                                       // the staging tile is free for the next layer
#endif
}

// W[O][I] (global, row-major) -> W^T[I][OP] in shared memory, padding zero
template <int I, int O, int OP>
__device__ __forceinline__ void stage_wt(const float* __restrict__ w, float* __restrict__ wt, int tid) {
  for (int e = tid; e < I * OP; e += kRows) {
    const int i = e / OP, o = e % OP;
    wt[e] = o < O ? w[o * I + i] : 0.f;
  }
}

template <int D, int A>
struct PpoSmem {
  static constexpr int DP = (D + 3) / 4 * 4, AP = (A + 3) / 4 * 4;
  float w1[D * kH], w2[kH * kH], w3[kH * AP];     // the current tower's W^T
  float b1[kH], b2[kH], b3[AP];
  static constexpr int kImage = D * kH + kH * kH + kH * AP + kH + kH + AP;   // floats of the six arrays above (contiguous, a multiple of 4)
  float x[kRows * kLd], h1[kRows * kLd], h2[kRows * kLd];        // per-row activations
  float dz2[kRows * kLd], dz3[kRows * kLd];                      // per-row pre-activation gradients; dz1 takes dz3's place once the
                                                                 // layer-3 weight gradients have consumed it: 5 row buffers + weights =
                                                                 // 103 KB per CTA = TWO CTAs per SM (six buffers: one)
  float gstage[33 * kGs];                                        // weight-gradient staging tile (+ one row of bias sums)
  float red[8];
};

// One tower's weights in PpoSmem order (W1^T [D][32] | W2^T [32][32] | W3^T [32][OP] | b1 | b2 | b3, zero padded), both towers
// back to back: written once per optimizer step, so that each of the hundreds of CTAs of a minibatch fills its shared memory
// with coalesced 16-byte copies instead of a strided gather (the gather was 19 % of the gradient kernel's time).
template <int D, int A>
__global__ void __launch_bounds__(256)
ppo_image_kernel(const float* __restrict__ params, float* __restrict__ image) {
  using S = PpoSmem<D, A>;
  constexpr int AP = S::AP, N = S::kImage;
  const PpoLayout L = ppo_layout(D, A);
  const int e0 = blockIdx.x * 256 + threadIdx.x;
  if (e0 == 0) { double* sums = reinterpret_cast<double*>(image) - 2; sums[0] = 0.0; sums[1] = 0.0; }   // the advantage sums in front of the image
  if (e0 >= 2 * N) return;
  const int tower = e0 / N;
  int e = e0 - tower * N;
  const int w1 = tower ? L.vf_w1 : L.pi_w1, b1 = tower ? L.vf_b1 : L.pi_b1, w2 = tower ? L.vf_w2 : L.pi_w2, b2 = tower ? L.vf_b2 : L.pi_b2;
  const int w3 = tower ? L.vf_w3 : L.pi_w3, b3 = tower ? L.vf_b3 : L.pi_b3;
  const int O3 = tower ? 1 : A, OP3 = tower ? 4 : AP;          // the value head uses the first [32][4] floats of the W3 slot
  float v = 0.f;
  if (e < D * kH) { const int i = e / kH, o = e % kH; v = params[w1 + o * D + i]; }
  else if ((e -= D * kH) < kH * kH) { const int i = e / kH, o = e % kH; v = params[w2 + o * kH + i]; }
  else if ((e -= kH * kH) < kH * AP) { if (e < kH * OP3) { const int i = e / OP3, o = e % OP3; v = o < O3 ? params[w3 + o * kH + i] : 0.f; } }
  else if ((e -= kH * AP) < kH) v = params[b1 + e];
  else if ((e -= kH) < kH) v = params[b2 + e];
  else { e -= kH; v = e < O3 ? params[b3 + e] : 0.f; }
  image[e0] = v;
}

// this CTA's copy of one tower's image into the six weight arrays (16-byte vectors)
template <class S>
__device__ __forceinline__ void stage_image(S& sm, const float* __restrict__ image, int tid) {
  static_assert(S::kImage % 4 == 0, "whole 16-byte vectors");
  const float4* src = reinterpret_cast<const float4*>(image);
  float4* dst = reinterpret_cast<float4*>(sm.w1);
  for (int e = tid; e < S::kImage / 4; e += kRows) dst[e] = src[e];
}

// GRAD = false: forward only, writes log-prob of the row's action and the value.
template <int D, int A, bool GRAD>
__global__ void __launch_bounds__(kRows, 2)
ppo_kernel(const __grid_constant__ PpoArgs a) {
  using S = PpoSmem<D, A>;
  constexpr int DP = S::DP, AP = S::AP;
  static_assert(DP <= 32 && AP <= 32 && DP <= kLd, "tower input / output widths");
  extern __shared__ __align__(16) unsigned char smem_raw[];
  S& sm = *reinterpret_cast<S*>(smem_raw);
  const PpoLayout L = ppo_layout(D, A);
  const int tid = threadIdx.x;
  const long long j = (long long)blockIdx.x * kRows + tid;
  const bool live = j < a.count;
  const int rows = (int)min((long long)kRows, a.count - (long long)blockIdx.x * kRows);
  const long long row = live ? (a.idx ? a.idx[a.first + j] : a.first + j) : 0;

  // ---- policy tower weights + this row's observation
  if (GRAD && a.image) {
    stage_image(sm, a.image, tid);
  } else {
    stage_wt<D, kH, kH>(a.params + L.pi_w1, sm.w1, tid);
    stage_wt<kH, kH, kH>(a.params + L.pi_w2, sm.w2, tid);
    stage_wt<kH, A, AP>(a.params + L.pi_w3, sm.w3, tid);
    if (tid < kH) { sm.b1[tid] = a.params[L.pi_b1 + tid]; sm.b2[tid] = a.params[L.pi_b2 + tid]; }
    if (tid < AP) sm.b3[tid] = tid < A ? a.params[L.pi_b3 + tid] : 0.f;
  }
  float* const xrow = sm.x + tid * kLd;
  float* const h1row = sm.h1 + tid * kLd;
  float* const h2row = sm.h2 + tid * kLd;
#pragma unroll
  for (int k = 0; k < DP; ++k) xrow[k] = (live && k < D) ? a.obs[row * D + k] : 0.f;
  __syncthreads();

  // ---- policy forward (activations go to this thread's shared rows: the backward pass and the weight gradients read them)
  float lg[AP];
  {
    float h[kH];
    dense_fwd<D, kH, kH>(sm.w1, sm.b1, xrow, h);
#pragma unroll
    for (int k = 0; k < kH; ++k) h[k] = tanh_exact(h[k]);
    row_store<kH>(sm.h1, tid, h);
    dense_fwd<kH, kH, kH>(sm.w2, sm.b2, h1row, h);
#pragma unroll
    for (int k = 0; k < kH; ++k) h[k] = tanh_exact(h[k]);
    row_store<kH>(sm.h2, tid, h);
    dense_fwd<kH, A, AP>(sm.w3, sm.b3, h2row, lg);
  }

  // ---- masked log-softmax, entropy, clipped surrogate (sb3_contrib masks logits with -1e8)
  const int act = live ? (int)a.actions[row] : 0;
  float mx = -3.0e38f;
  uint32_t valid = 0;
#pragma unroll
  for (int k = 0; k < A; ++k) {
    const bool v = live ? a.mask[row * A + k] != 0 : k == 0;
    valid |= (uint32_t)v << k;
    lg[k] = v ? lg[k] : -1e8f;
    mx = fmaxf(mx, lg[k]);
  }
  float sum = 0.f;
#pragma unroll
  for (int k = 0; k < A; ++k) { lg[k] = __expf(lg[k] - mx); sum += lg[k]; }     // lg[k]: unnormalised probability
  const float inv = 1.0f / sum, lsum = __logf(sum);
  float logp_a = 0.f, ent = 0.f;
#pragma unroll
  for (int k = 0; k < A; ++k) {
    const float p = lg[k] * inv;
    const float lp = __logf(fmaxf(lg[k], 1e-37f)) - lsum;                       // log p_k (finite for masked entries too)
    if ((valid >> k) & 1u) ent -= p * lp;
    if (k == act) logp_a = lp;
    lg[k] = p;
  }

  if (!GRAD) {
    if (live && a.logp_out) a.logp_out[row] = logp_a;
  } else {
    float advn = live ? a.adv[row] : 0.f;
    if (a.adv_sums) {        // (a - mean) / (std + 1e-8), unbiased std like torch.Tensor.std
      const double sa = a.adv_sums[0], mean = sa / (double)a.count;
      const double var = a.count > 1 ? fmax(0.0, (a.adv_sums[1] - sa * mean) / (double)(a.count - 1)) : 0.0;
      advn = (advn - (float)mean) * (float)(1.0 / (sqrt(var) + 1e-8));
    }
    const float ratio = __expf(logp_a - (live ? a.old_logp[row] : 0.f));
    const bool clipped = (advn > 0.f && ratio > 1.f + a.clip) || (advn < 0.f && ratio < 1.f - a.clip);
    const float surr = fminf(advn * ratio, advn * fminf(fmaxf(ratio, 1.f - a.clip), 1.f + a.clip));
    const float g_logp = (live && !clipped) ? -advn * ratio * a.inv_count : 0.f;   // d(-min(...)) / d logp, mean over the minibatch
    const float g_ent = live ? a.ent_coef * a.inv_count : 0.f;                     // d(-ent_coef * H) / dz_k = ent_coef * p_k (log p_k + H)
    float dz3[kH];
#pragma unroll
    for (int k = 0; k < kH; ++k) {
      float g = 0.f;
      if (k < A && ((valid >> k) & 1u)) {
        const float p = lg[k < AP ? k : 0], lp = __logf(fmaxf(p, 1e-37f));
        g = g_logp * ((k == act ? 1.f : 0.f) - p) + g_ent * p * (lp + ent);
      }
      dz3[k] = g;
    }
    row_store<kH>(sm.dz3, tid, dz3);
    // backward through the two tanh layers
    dense_bwd_tanh<kH, AP>(sm.w3, reinterpret_cast<const float(&)[AP]>(dz3), h2row, sm.dz2 + tid * kLd);
    // loss statistics: warp sums -> one atomic per warp
    float s0 = live ? -surr : 0.f, s2 = live ? ent : 0.f, s3 = (live && clipped) ? 1.f : 0.f;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      s0 += __shfl_xor_sync(0xffffffffu, s0, o); s2 += __shfl_xor_sync(0xffffffffu, s2, o); s3 += __shfl_xor_sync(0xffffffffu, s3, o);
    }
    if ((tid & 31) == 0 && a.stats) { atomicAdd(&a.stats[0], s0); atomicAdd(&a.stats[2], s2); atomicAdd(&a.stats[3], s3); }
    __syncthreads();
    weight_grad<kH, A>(sm.dz3, sm.h2, rows, a.grads + L.pi_w3, a.grads + L.pi_b3, sm.gstage, tid);
    __syncthreads();                                                  // dz3 has been consumed: its rows take dz1
    float dz2[kH];
    row_load<kH>(sm.dz2, tid, dz2);
    dense_bwd_tanh<kH, kH>(sm.w2, dz2, h1row, sm.dz3 + tid * kLd);
    __syncthreads();
    weight_grad<kH, kH>(sm.dz2, sm.h1, rows, a.grads + L.pi_w2, a.grads + L.pi_b2, sm.gstage, tid);
    weight_grad<D, kH>(sm.dz3, sm.x, rows, a.grads + L.pi_w1, a.grads + L.pi_b1, sm.gstage, tid);
  }
  __syncthreads();

  // ---- value tower (same buffers)
  if (GRAD && a.image) {
    stage_image(sm, a.image + S::kImage, tid);
  } else {
    stage_wt<D, kH, kH>(a.params + L.vf_w1, sm.w1, tid);
    stage_wt<kH, kH, kH>(a.params + L.vf_w2, sm.w2, tid);
    stage_wt<kH, 1, 4>(a.params + L.vf_w3, sm.w3, tid);
    if (tid < kH) { sm.b1[tid] = a.params[L.vf_b1 + tid]; sm.b2[tid] = a.params[L.vf_b2 + tid]; }
    if (tid < 4) sm.b3[tid] = tid == 0 ? a.params[L.vf_b3] : 0.f;
  }
  __syncthreads();
  float v4[4];
  {
    float h[kH];
    dense_fwd<D, kH, kH>(sm.w1, sm.b1, xrow, h);
#pragma unroll
    for (int k = 0; k < kH; ++k) h[k] = tanh_exact(h[k]);
    row_store<kH>(sm.h1, tid, h);
    dense_fwd<kH, kH, kH>(sm.w2, sm.b2, h1row, h);
#pragma unroll
    for (int k = 0; k < kH; ++k) h[k] = tanh_exact(h[k]);
    row_store<kH>(sm.h2, tid, h);
    dense_fwd<kH, 1, 4>(sm.w3, sm.b3, h2row, v4);
  }
  const float v = v4[0];
  if (!GRAD) {
    if (live && a.value_out) a.value_out[row] = v;
    return;
  }
  const float err = live ? v - a.ret[row] : 0.f;
  const float dz3v[4] = {live ? 2.f * a.vf_coef * err * a.inv_count : 0.f, 0.f, 0.f, 0.f};   // d(vf_coef * mse) / dv
#pragma unroll
  for (int k = 0; k < kH / 4; ++k)
    *reinterpret_cast<float4*>(sm.dz3 + tid * kLd + 4 * k) = make_float4(k == 0 ? dz3v[0] : 0.f, 0.f, 0.f, 0.f);
  dense_bwd_tanh<kH, 4>(sm.w3, dz3v, h2row, sm.dz2 + tid * kLd);
  float s1 = err * err, s4 = live ? 1.f : 0.f;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { s1 += __shfl_xor_sync(0xffffffffu, s1, o); s4 += __shfl_xor_sync(0xffffffffu, s4, o); }
  if ((tid & 31) == 0 && a.stats) { atomicAdd(&a.stats[1], s1); atomicAdd(&a.stats[4], s4); }
  __syncthreads();
  weight_grad<kH, 1>(sm.dz3, sm.h2, rows, a.grads + L.vf_w3, a.grads + L.vf_b3, sm.gstage, tid);
  __syncthreads();
  float dz2[kH];
  row_load<kH>(sm.dz2, tid, dz2);
  dense_bwd_tanh<kH, kH>(sm.w2, dz2, h1row, sm.dz3 + tid * kLd);
  __syncthreads();
  weight_grad<kH, kH>(sm.dz2, sm.h1, rows, a.grads + L.vf_w2, a.grads + L.vf_b2, sm.gstage, tid);
  weight_grad<D, kH>(sm.dz3, sm.x, rows, a.grads + L.vf_w1, a.grads + L.vf_b1, sm.gstage, tid);
}

// sum and sum of squares of adv[idx[first .. first + count)] added into out[0..1] (float64; zeroed by ppo_image_kernel, which runs
// just before on the same stream): a grid of CTAs with a grid-stride loop — one CTA alone needed ~0.25 us per 1 024 rows
__global__ void __launch_bounds__(1024)
adv_sums_kernel(const float* __restrict__ adv, const long long* __restrict__ idx, long long first, long long count, double* __restrict__ out) {
  __shared__ double sh[2][32];
  double s = 0.0, q = 0.0;
  for (long long j = (long long)blockIdx.x * 1024 + threadIdx.x; j < count; j += (long long)gridDim.x * 1024) {
    const double v = (double)adv[idx ? idx[first + j] : first + j];
    s += v; q += v * v;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { s += __shfl_xor_sync(0xffffffffu, s, o); q += __shfl_xor_sync(0xffffffffu, q, o); }
  if ((threadIdx.x & 31) == 0) { sh[0][threadIdx.x >> 5] = s; sh[1][threadIdx.x >> 5] = q; }
  __syncthreads();
  if (threadIdx.x < 32) {
    s = sh[0][threadIdx.x]; q = sh[1][threadIdx.x];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { s += __shfl_xor_sync(0xffffffffu, s, o); q += __shfl_xor_sync(0xffffffffu, q, o); }
    if (threadIdx.x == 0) { atomicAdd(&out[0], s); atomicAdd(&out[1], q); }
  }
}

// One CTA: global-norm clip (torch.nn.utils.clip_grad_norm_) + Adam (torch.optim.Adam, no weight decay) + zero the gradients.
__global__ void __launch_bounds__(1024)
adam_kernel(float* __restrict__ params, float* __restrict__ grads, float* __restrict__ m, float* __restrict__ v, int* __restrict__ step,
            int n, float lr, float beta1, float beta2, float eps, float max_norm) {
  __shared__ float sh[32];
  __shared__ float s_scale;
  float q = 0.f;
  for (int k = threadIdx.x; k < n; k += 1024) q += grads[k] * grads[k];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = q;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int w = 0; w < 32; ++w) t += sh[w];
    const float norm = sqrtf(t);
    s_scale = max_norm > 0.f ? fminf(1.f, max_norm / (norm + 1e-6f)) : 1.f;
    *step += 1;
  }
  __syncthreads();
  const float scale = s_scale;
  const int t = *step;
  const float c1 = 1.f - powf(beta1, (float)t), c2 = 1.f - powf(beta2, (float)t);
  const float step_size = lr / c1, rc2 = rsqrtf(c2);
  for (int k = threadIdx.x; k < n; k += 1024) {
    const float g = grads[k] * scale;
    const float mk = beta1 * m[k] + (1.f - beta1) * g;
    const float vk = beta2 * v[k] + (1.f - beta2) * g * g;
    m[k] = mk; v[k] = vk;
    params[k] -= step_size * mk / (sqrtf(vk) * rc2 + eps);
    grads[k] = 0.f;
  }
}

// GAE(lambda) per env, backwards over the rollout (`terminated` ends the episode; SB3 RolloutBuffer.compute_returns_and_advantage)
__global__ void __launch_bounds__(256)
gae_kernel(int T, long long n, const float* __restrict__ rew, const float* __restrict__ val, const uint8_t* __restrict__ done,
           const float* __restrict__ last_val, float gamma, float lam, float* __restrict__ adv, float* __restrict__ ret) {
  const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
  if (i >= n) return;
  float gae = 0.f, next_v = last_val[i];
  for (int t = T - 1; t >= 0; --t) {
    const long long o = (long long)t * n + i;
    const float nonterm = done[o] ? 0.f : 1.f, v = val[o];
    const float delta = rew[o] + gamma * next_v * nonterm - v;
    gae = delta + gamma * lam * nonterm * gae;
    adv[o] = gae; ret[o] = gae + v;
    next_v = v;
  }
}

template <int D, int A, bool GRAD>
cudaError_t launch_ppo(const PpoArgs& a, cudaStream_t st) {
  static bool prepared[64] = {};      // opt-in shared-memory size, once per device and instantiation (idempotent)
  const int smem = (int)sizeof(PpoSmem<D, A>);
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64 || !prepared[dev]) {
    cudaError_t e = cudaFuncSetAttribute(ppo_kernel<D, A, GRAD>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return e;
    if (dev >= 0 && dev < 64) prepared[dev] = true;
  }
  if (a.count <= 0) return cudaSuccess;
  ppo_kernel<D, A, GRAD><<<(unsigned)((a.count + kRows - 1) / kRows), kRows, smem, st>>>(a);
  return cudaGetLastError();
}

template <bool GRAD>
cudaError_t launch_ppo_da(int D, int A, const PpoArgs& a, cudaStream_t st) {
  if (D == 29 && A == 22) return launch_ppo<29, 22, GRAD>(a, st);
  if (D == 16 && A == 11) return launch_ppo<16, 11, GRAD>(a, st);
  if (D == 13 && A == 2) return launch_ppo<13, 2, GRAD>(a, st);
  return cudaErrorInvalidValue;
}

PpoArgs make_args(const msort_ppo_batch_t& b, const float* params) {
  PpoArgs a{};
  a.params = params; a.obs = b.obs; a.mask = b.mask; a.actions = (const long long*)b.actions;
  a.old_logp = b.old_logp; a.adv = b.adv; a.ret = b.ret;
  return a;
}

}  // namespace

int ppo_param_count(int D, int A) { return ppo_layout(D, A).total; }

// scratch of msort_ppo_gradient / _update (floats): [0..1] advantage statistics, [4 ..) the two towers' weight image
static int ppo_image_floats(int D, int A) {
  if (D == 29 && A == 22) return 2 * PpoSmem<29, 22>::kImage;
  if (D == 16 && A == 11) return 2 * PpoSmem<16, 11>::kImage;
  if (D == 13 && A == 2) return 2 * PpoSmem<13, 2>::kImage;
  return 0;
}
int ppo_scratch_floats(int D, int A) { const int n = ppo_image_floats(D, A); return n ? 4 + n : 0; }

cudaError_t ppo_forward(const msort_ppo_batch_t& b, const float* params, float* logp_out, float* value_out, cudaStream_t st) {
  PpoArgs a = make_args(b, params);
  a.first = 0; a.count = b.num_rows; a.logp_out = logp_out; a.value_out = value_out;
  return launch_ppo_da<false>(b.obs_dim, b.num_actions, a, st);
}

cudaError_t ppo_gradient(const msort_ppo_batch_t& b, const msort_ppo_hparams_t& hp, const float* params, float* grads,
                         const int64_t* idx, long long first, long long count, float* adv_stats, float* stats, cudaStream_t st) {
  float* image = adv_stats + 4;              // scratch: two float64 advantage sums (16 bytes), then the weight image
  {
    const int D = b.obs_dim, A = b.num_actions, n = ppo_image_floats(D, A);
    const unsigned g = (unsigned)((n + 255) / 256);
    if (D == 29 && A == 22) ppo_image_kernel<29, 22><<<g, 256, 0, st>>>(params, image);
    else if (D == 16 && A == 11) ppo_image_kernel<16, 11><<<g, 256, 0, st>>>(params, image);
    else if (D == 13 && A == 2) ppo_image_kernel<13, 2><<<g, 256, 0, st>>>(params, image);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
  }
  if (hp.normalize_advantage) {              // after the image kernel: it zeroes the sums
    const unsigned g = (unsigned)std::min<long long>(64, (count + 4095) / 4096);
    adv_sums_kernel<<<g, 1024, 0, st>>>(b.adv, (const long long*)idx, first, count, reinterpret_cast<double*>(adv_stats));
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
  }
  PpoArgs a = make_args(b, params);
  a.image = image;
  a.grads = grads; a.idx = (const long long*)idx; a.first = first; a.count = count;
  a.adv_sums = hp.normalize_advantage ? reinterpret_cast<const double*>(adv_stats) : nullptr;
  a.clip = hp.clip_range; a.vf_coef = hp.vf_coef; a.ent_coef = hp.ent_coef; a.inv_count = 1.0f / (float)count; a.stats = stats;
  return launch_ppo_da<true>(b.obs_dim, b.num_actions, a, st);
}

cudaError_t ppo_adam(float* params, float* grads, float* m, float* v, int* step, int n, const msort_ppo_hparams_t& hp, cudaStream_t st) {
  adam_kernel<<<1, 1024, 0, st>>>(params, grads, m, v, step, n, hp.learning_rate, hp.beta1, hp.beta2, hp.adam_eps, hp.max_grad_norm);
  return cudaGetLastError();
}

cudaError_t ppo_gae(int T, long long n, const float* rew, const float* val, const uint8_t* done, const float* last_val, float gamma,
                    float lam, float* adv, float* ret, cudaStream_t st) {
  if (n <= 0 || T <= 0) return cudaSuccess;
  gae_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(T, n, rew, val, done, last_val, gamma, lam, adv, ret);
  return cudaGetLastError();
}

}  // namespace msort
