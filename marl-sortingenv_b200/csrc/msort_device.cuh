// msort_device.cuh — device-side plant model shared by all kernels (sm_100a).
//
// State layout in HBM (DESIGN.md "Data layout"): structure-of-arrays in 16-byte planes,
//   plane p, env i  ->  ((uint4*)state)[p * n_pad + i]
// so a warp reads/writes 512 contiguous bytes per plane with one LDG.128/STG.128 per lane.
//   hot planes (touched every step)
//     P0  u8 input[4] | u8 belt[4] | u8 sorting[4] | timer1, timer2, mat1|mat2<<4, flags
//     P1  i32 true[A..D]          P2  i32 false[A..D]
//     P3  i32 E, n1, n2, last_press_amount
//     P4  u32 step | q1,q2,gen_counter,0 | u32 episode | i32 replay_cursor
//     P5  f64 episode_return | 8 B spare
//     P6  f64 acc_belt[A,B]       P7  f64 acc_belt[C,D]   (REPLAY mode only: in PHILOX mode the
//         accuracies are a pure function of (env, episode, step, last mode) and are recomputed)
//   cold planes (one per material, touched only when a press finishes that material)
//     P8+m  u32 bale_n[m], bale_sum[m], last_size[m] | last_q[m]<<24, spare
//
// Arithmetic policy.  Everything that can reach INTEGER state is computed so
// that it equals the reference's float64 result exactly:
//   * rint(t*acc), the accuracy noise: float64 with explicit round-to-nearest intrinsics
//     (never contracted to FMA; numpy does not fuse);
//   * 2-decimal purity / bale quality rint(true/total*100): exact integer rounding, which
//     provably equals the float64 pipeline except on exact .5 ties, whose float64 outcome
//     depends only on the tie value and comes from a host-built table (purity_k);
//   * fill-ratio and bale-remainder comparisons: host-precomputed integer thresholds.
// Observations and rewards are emitted as float32 within 1e-5 relative of the reference
// (they never feed back into state), so they use float32 / reciprocal arithmetic.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/msort.h"

namespace msort {

constexpr int kHotPlanes = 8;
constexpr int kColdPlanes = 5;
constexpr int kPlanes = kHotPlanes + kColdPlanes;
#ifndef MSORT_TILE
#define MSORT_TILE 128
#endif
constexpr int kTile = MSORT_TILE;  // envs per CTA tile == threads per CTA (32, 64 or 128)

// ---------------------------------------------------------------- mbarrier / TMA bulk-copy helpers (step + policy kernels)
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count) : "memory");
}

// one poll; the thread may stay suspended inside the instruction for up to ~`hint_ns` (suspend-time hint), so a
// waiting warp costs a handful of issue slots instead of a spin loop's worth
#ifndef MSORT_MBAR_HINT_NS
#define MSORT_MBAR_HINT_NS 4000
#endif
__device__ __forceinline__ uint32_t mbar_try_wait(uint32_t bar_saddr, uint32_t parity, uint32_t hint_ns = MSORT_MBAR_HINT_NS) {
  uint32_t done;
#if MSORT_MBAR_HINT_NS == 0   // experiment: non-blocking poll
  asm volatile(
      "{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(done) : "r"(bar_saddr), "r"(parity) : "memory");
  return done;
#endif
  asm volatile(
      "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\tselp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(done) : "r"(bar_saddr), "r"(parity), "r"(hint_ns) : "memory");
  return done;
}

// Wait for a phase of an mbarrier.  Fast polls first; a wait that outlasts them (time-slicing, a debugger, a
// profiler replay) backs off with nanosleep and is bounded by WALL CLOCK — 20 s of %globaltimer — so that a
// mis-programmed copy / MMA still fails loudly instead of hanging the GPU, while a legitimately slow wait survives.
static __device__ __noinline__ void mbar_wait_slow(uint32_t a, uint32_t parity) {
  unsigned long long t0, t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
  for (;;) {
    if (mbar_try_wait(a, parity)) return;
    __nanosleep(256);
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    if (t - t0 > 20000000000ull) __trap();
  }
}

__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t a = smem_u32(bar);
#pragma unroll 1
  for (uint32_t spin = 0; spin < 2048u; ++spin)
    if (mbar_try_wait(a, parity)) return;
  mbar_wait_slow(a, parity);
}

// TMA bulk copy global -> shared, completion counted in bytes on an mbarrier
__device__ __forceinline__ void bulk_load(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               :: "r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(bar)), "r"(bytes) : "memory");
}

// Philox draw blocks (shared with oracle/msort_oracle.c)
constexpr uint32_t kBlkNoise = 0, kBlkPress = 1, kBlkReset = 2, kBlkInput = 3, kBlkRedis = 16;

// Kernel-parameter copy of msort_config_t plus host-precomputed constants.
struct DevConfig {
  long long n, n_pad, gid0;
  int kind, max_steps;
  int sm_count;                    // multiprocessors of the handle's device
  unsigned flags;
  int rng_mode;
  int layout;                      // LAYOUT_* of the state blob (fixed at create)
  unsigned rk[20];  // Philox round keys: rk[2r] = key0 + r*W0, rk[2r+1] = key1 + r*W1
  int batch, spp;
  unsigned pat[2];  // packed u8x4 counts of pattern 1 / 2
  int pat_remainder;
  int stage_cap, cap, S;
  int press_time[2];
  int lvl_cat, lvl_sev, lvl_mild;  // smallest level with level/cap > 1.0 / 0.95 / 0.90 (f64)
  int rem_new_bale;                // smallest remainder with rem > S*threshold (f64)
  int qthr100[4];                  // 100*quality_threshold when that is an integer (fast_pdiff)
  int fast_pdiff;                  // 1: obs purity diff == (k - qthr100)/100 within tolerance for all k
  int fast;                        // 1: the host proved (digest_config) that a boosted accuracy always clips to
                                   //    exactly 1.0, an unboosted one never clips, and a batch fits 7 bits, so the
                                   //    step kernel may run its FAST instantiation (same results, fewer instructions)
  unsigned tie_up[4];              // bit k: the float64 pipeline rounds the exact tie (2k+1)/200 up to k+1 (purity_k)
  int small_lv;                    // 1: every container level is provably <= 8192 (compact layout, batch*max_steps
                                   //    <= 8192): the float32 quotient in purity_k then rounds to the right integer
  int one_block;                   // 1: no station can ever have more than 12 mis-sorted units (one Philox block of draws)
  unsigned S_magic;                // floor(2^32/S)+1: n/S == umulhi(n, S_magic) for n, S < 2^16
  double pen_sev0, pen_mild0;      // min(0, severe / mild overflow penalty) (FAST press reward)
  float obs_belt_tab[3][5];        // FAST: belt part of the observation for belt == pattern 1 / pattern 2 / empty
  float obs_sort_tab[3][4];        // FAST: sorting-stage part of the observation, same three cases
  float inv_cap, inv_stage, inv_pt[2];
  double base_acc[4], boost, noise_low, noise_range;
  double qthr[4];
  double qthr_empty[4];            // purity an EMPTY container reports: Python round(threshold, 2) (env_super.py:788-789)
  float pdiff_empty[4];            // ... and its purity-difference observation round(qthr_empty - threshold, 2) (:222-225); 0 for whole percents
  double theta4, c_sort;           // sum of 4 thetas; (scaling/4)/temperature
  double c_state, c_eff;           // max_state/(5*cap); 4/S
  double pen_cat, pen_sev, pen_mild, bef, ovf_pen;
  const double* sort_lut;          // kSortLut float64 sorting rewards indexed by the purity sum (see sort_reward)
};

// Env_2's embedded 13->32->32->2 policy on the tensor cores (step kernel, TCMLP instantiation):
// packed operand buffer built by pack_policy_tc() (msort_kernels.cu).  fp16 weight tiles in the K-major core-matrix
// order [k/8][n][k%8], each weight as a THREE-term fp16 split (two for the scaled last layer), then the fp32 biases
// the epilogues add.  Offsets in fp16 elements / 32-bit words.
constexpr int kTcB1 = 0;                          // layer 1: 3 term tiles [K=16 x N=32] (rows 13..15 of term 0 = bias split)
constexpr int kTcB2 = kTcB1 + 3 * 16 * 32;        // layer 2: 3 term tiles [K=32 x N=32]
constexpr int kTcHalves = kTcB2 + 3 * 32 * 32;
constexpr int kTcBias2 = kTcHalves / 2;           // 32 floats: layer-2 bias (with the folded constants)
constexpr int kTcW3 = kTcBias2 + 32;              // 32 float pairs (-2 W3[0][j], -2 W3[1][j]): the output layer runs as fp32 FFMA2
constexpr int kTcBias3 = kTcW3 + 64;              // 2 floats (+2 padding)
constexpr int kTcWords = kTcBias3 + 4;            // 2404 words = 9 616 B (a multiple of 16)
static_assert(kTcWords % 4 == 0, "whole 16-byte vectors");

constexpr int kSortLut = 401;      // purity sum in hundredths: 4 containers x (0..100)

// ---------------------------------------------------------------- exact f64 helpers
__device__ __forceinline__ double dadd(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double dsub(double a, double b) { return __dsub_rn(a, b); }
__device__ __forceinline__ double dmul(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double ddiv(double a, double b) { return __ddiv_rn(a, b); }
// numpy-scalar round(x, 2) == rint(x*100)/100
__device__ __forceinline__ double round2(double x) { return ddiv(rint(dmul(x, 100.0)), 100.0); }
__device__ __forceinline__ double clipd(double x, double lo, double hi) {
  return x < lo ? lo : (x > hi ? hi : x);
}
__device__ __forceinline__ float clipf(float x, float lo, float hi) { return fminf(fmaxf(x, lo), hi); }

// ---------------------------------------------------------------- Philox4x32-10
struct U4 { uint32_t x, y, z, w; };

// 10 rounds, round keys precomputed on the host (uniform kernel parameters).
__device__ __forceinline__ U4 philox_rk(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                        const unsigned (&rk)[20]) {
#ifndef MSORT_PHILOX_ROUNDS
#define MSORT_PHILOX_ROUNDS 10   // anything else is a timing experiment, never a product build
#endif
#pragma unroll
  for (int r = 0; r < MSORT_PHILOX_ROUNDS; ++r) {
    unsigned long long p0 = (unsigned long long)0xD2511F53u * c0;
    unsigned long long p1 = (unsigned long long)0xCD9E8D57u * c2;
    c0 = (uint32_t)(p1 >> 32) ^ c1 ^ rk[2 * r];
    c1 = (uint32_t)p1;
    c2 = (uint32_t)(p0 >> 32) ^ c3 ^ rk[2 * r + 1];
    c3 = (uint32_t)p0;
  }
  return U4{c0, c1, c2, c3};
}

// plain form (explicit key) for kernels off the hot path
__device__ __forceinline__ U4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                            uint32_t k0, uint32_t k1) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    uint32_t h0 = __umulhi(0xD2511F53u, c0), l0 = 0xD2511F53u * c0;
    uint32_t h1 = __umulhi(0xCD9E8D57u, c2), l1 = 0xCD9E8D57u * c2;
    c0 = h1 ^ c1 ^ k0; c1 = l1; c2 = h0 ^ c3 ^ k1; c3 = l0;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  return U4{c0, c1, c2, c3};
}

// counter = {gid_lo, (gid_hi & 0xffff) | block<<16, episode, step}, key = seed
__device__ __forceinline__ U4 env_draw(const DevConfig& c, uint32_t gid_lo, uint32_t gid_hi16, uint32_t block,
                                       uint32_t episode, uint32_t step) {
  return philox_rk(gid_lo, gid_hi16 | (block << 16), episode, step, c.rk);
}

// One redistribution draw from a 64-bit lane (xh:xl) of a Philox block: r = floor(lane * tot / 2^64) is
// uniform on [0, tot) up to tot/2^64; the low 64 bits of the product are the lane for the next draw
// (they are uniform on a progression of 2^64/tot points, so draw j of a lane is biased by <= tot^j / 2^64:
// six draws per lane with tot <= 100 stay below 6e-8).  Shared with oracle/msort_oracle.c.
__device__ __forceinline__ uint32_t draw64(uint32_t& xl, uint32_t& xh, uint32_t tot) {
  const unsigned long long p0 = (unsigned long long)xl * tot;
  const unsigned long long p1 = (unsigned long long)xh * tot + (p0 >> 32);
  xl = (uint32_t)p0; xh = (uint32_t)p1;
  return (uint32_t)(p1 >> 32);
}

// ---------------------------------------------------------------- env registers
struct Env {
  uint32_t in4, belt4, sort4;  // packed u8x4 stage counts (A | B<<8 | C<<16 | D<<24)
  int tr[4], fl[4], e;         // containers
  int timer[2], mat[2], pn[2], pq[2];
  int started, last_amt;
  int gfirst, gidx, gcount;    // generator: gfirst 0 -> pattern 1 first, 1 -> pattern 2 first
  int mode;
  uint32_t step, episode;
  int cursor;
  double ep_ret;
  double acc[4];
};

__device__ __forceinline__ int b4(uint32_t v, int m) { return (int)((v >> (8 * m)) & 0xffu); }
__device__ __forceinline__ int sum4(uint32_t v) { return (int)__dp4a(v, 0x01010101u, 0u); }

__device__ __forceinline__ double u2d(uint32_t lo, uint32_t hi) { return __hiloint2double((int)hi, (int)lo); }

__device__ __forceinline__ void unpack_env(const uint4& p0, const uint4& p1, const uint4& p2, const uint4& p3,
                                           const uint4& p4, const uint4& p5, const uint4& p6, const uint4& p7,
                                           Env& s) {
  s.in4 = p0.x; s.belt4 = p0.y; s.sort4 = p0.z;
  s.timer[0] = p0.w & 0xff; s.timer[1] = (p0.w >> 8) & 0xff;
  s.mat[0] = (p0.w >> 16) & 0xf; s.mat[1] = (p0.w >> 20) & 0xf;
  uint32_t fl = p0.w >> 24;
  s.gfirst = fl & 1; s.gidx = (fl >> 1) & 1; s.started = (fl >> 2) & 1; s.mode = (fl >> 3) & 1;
  s.tr[0] = (int)p1.x; s.tr[1] = (int)p1.y; s.tr[2] = (int)p1.z; s.tr[3] = (int)p1.w;
  s.fl[0] = (int)p2.x; s.fl[1] = (int)p2.y; s.fl[2] = (int)p2.z; s.fl[3] = (int)p2.w;
  s.e = (int)p3.x; s.pn[0] = (int)p3.y; s.pn[1] = (int)p3.z; s.last_amt = (int)p3.w;
  s.step = p4.x; s.pq[0] = p4.y & 0xff; s.pq[1] = (p4.y >> 8) & 0xff; s.gcount = (p4.y >> 16) & 0xff;
  s.episode = p4.z; s.cursor = (int)p4.w;
  s.ep_ret = u2d(p5.x, p5.y);
  s.acc[0] = u2d(p6.x, p6.y); s.acc[1] = u2d(p6.z, p6.w);
  s.acc[2] = u2d(p7.x, p7.y); s.acc[3] = u2d(p7.z, p7.w);
}

// update_accuracy (env_super.py:492-509) for one material: clip(base [+ boost] + (low + range*u), 0, 1)
__device__ __forceinline__ double accuracy_of(const DevConfig& c, int m, int mode, double u) {
  double base = c.base_acc[m];
  if ((m & 1) == mode) base = dadd(base, c.boost);
  return clipd(dadd(base, dadd(c.noise_low, dmul(c.noise_range, u))), 0.0, 1.0);
}

// PHILOX mode: accuracy_belt after step (episode, step) under `mode`
__device__ __forceinline__ void philox_accuracy(const DevConfig& c, uint32_t gid_lo, uint32_t gid_hi16, uint32_t episode,
                                                uint32_t step, int mode, double acc[4]) {
  const U4 r4 = env_draw(c, gid_lo, gid_hi16, kBlkNoise, episode, step);
  acc[0] = accuracy_of(c, 0, mode, (double)r4.x * 2.3283064365386963e-10);
  acc[1] = accuracy_of(c, 1, mode, (double)r4.y * 2.3283064365386963e-10);
  acc[2] = accuracy_of(c, 2, mode, (double)r4.z * 2.3283064365386963e-10);
  acc[3] = accuracy_of(c, 3, mode, (double)r4.w * 2.3283064365386963e-10);
}

// FAST form (DevConfig::fast): only the two stations the mode does NOT boost carry noise — a = station 0
// (mode 1) or 1 (mode 0), b = station 2 (mode 1) or 3 (mode 0); the boosted ones are exactly 1.0 and
// the clip is the identity for the others.  Same words, same float64 operations as accuracy_of.
__device__ __forceinline__ void philox_accuracy2(const DevConfig& c, uint32_t gid_lo, uint32_t gid_hi16, uint32_t episode,
                                                 uint32_t step, int mode, double& a, double& b) {
  const U4 r4 = env_draw(c, gid_lo, gid_hi16, kBlkNoise, episode, step);
  const uint32_t ua = mode ? r4.x : r4.y, ub = mode ? r4.z : r4.w;
  const double ba = c.base_acc[1 - mode], bb = c.base_acc[3 - mode];   // mode 1 -> stations 0 and 2, mode 0 -> 1 and 3
  a = dadd(ba, dadd(c.noise_low, dmul(c.noise_range, (double)ua * 2.3283064365386963e-10)));
  b = dadd(bb, dadd(c.noise_low, dmul(c.noise_range, (double)ub * 2.3283064365386963e-10)));
}

// ---------------------------------------------------------------- state layouts
// LAYOUT_REPLAY : planes P0..P7 (float64 accuracies stored: they come from recorded streams)
// LAYOUT_WIDE   : planes P0..P5 (PHILOX: accuracies recomputed from the counter)
// LAYOUT_COMPACT: PHILOX with auto-reset and 100*max_steps <= 65535, so that every container
//                 level / press amount is provably < 2^16 (at most 100 units enter per step):
//     C0 = P0
//     C1 = u16 true[A..D], u16 false[A..D]
//     C2 = u16 E, n1, n2, last_press_amount | f64 episode_return
//     C3 = P4
//   64 B/env instead of 96 B — 64 B less DRAM traffic per env-step.
enum { LAYOUT_REPLAY = 0, LAYOUT_WIDE = 1, LAYOUT_COMPACT = 2 };

template <int LAYOUT>
__device__ __forceinline__ void load_planes(const uint4* __restrict__ st, long long n_pad, long long i, Env& s) {
  if (LAYOUT == LAYOUT_COMPACT) {
    const uint4 c0 = st[i], c1 = st[n_pad + i], c2 = st[2 * n_pad + i], c3 = st[3 * n_pad + i];
    const uint4 p1 = make_uint4(c1.x & 0xffffu, c1.x >> 16, c1.y & 0xffffu, c1.y >> 16);
    const uint4 p2 = make_uint4(c1.z & 0xffffu, c1.z >> 16, c1.w & 0xffffu, c1.w >> 16);
    const uint4 p3 = make_uint4(c2.x & 0xffffu, c2.x >> 16, c2.y & 0xffffu, c2.y >> 16);
    const uint4 p5 = make_uint4(c2.z, c2.w, 0u, 0u), z = make_uint4(0u, 0u, 0u, 0u);
    unpack_env(c0, p1, p2, p3, c3, p5, z, z, s);
  } else {
    const uint4 p0 = st[i], p1 = st[n_pad + i], p2 = st[2 * n_pad + i], p3 = st[3 * n_pad + i];
    const uint4 p4 = st[4 * n_pad + i], p5 = st[5 * n_pad + i];
    uint4 p6 = make_uint4(0u, 0u, 0u, 0u), p7 = p6;
    if (LAYOUT == LAYOUT_REPLAY) { p6 = st[6 * n_pad + i]; p7 = st[7 * n_pad + i]; }
    unpack_env(p0, p1, p2, p3, p4, p5, p6, p7, s);
  }
}

__device__ __forceinline__ uint2 d2u(double d) {
  return make_uint2((uint32_t)__double2loint(d), (uint32_t)__double2hiint(d));
}

template <int LAYOUT>
__device__ __forceinline__ void store_planes(uint4* __restrict__ st, long long n_pad, long long i, const Env& s) {
  const uint32_t fl = (uint32_t)(s.gfirst | (s.gidx << 1) | (s.started << 2) | ((s.mode & 1) << 3));
  const uint32_t w = (uint32_t)s.timer[0] | ((uint32_t)s.timer[1] << 8) | ((uint32_t)s.mat[0] << 16) |
                     ((uint32_t)s.mat[1] << 20) | (fl << 24);
  const uint4 p0 = make_uint4(s.in4, s.belt4, s.sort4, w);
  const uint4 p4 = make_uint4(s.step, (uint32_t)s.pq[0] | ((uint32_t)s.pq[1] << 8) | ((uint32_t)s.gcount << 16),
                              s.episode, (uint32_t)s.cursor);
  const uint2 r = d2u(s.ep_ret);
  if (LAYOUT == LAYOUT_COMPACT) {
    st[i] = p0;
    st[n_pad + i] = make_uint4((uint32_t)s.tr[0] | ((uint32_t)s.tr[1] << 16), (uint32_t)s.tr[2] | ((uint32_t)s.tr[3] << 16),
                               (uint32_t)s.fl[0] | ((uint32_t)s.fl[1] << 16), (uint32_t)s.fl[2] | ((uint32_t)s.fl[3] << 16));
    st[2 * n_pad + i] = make_uint4((uint32_t)s.e | ((uint32_t)s.pn[0] << 16), (uint32_t)s.pn[1] | ((uint32_t)s.last_amt << 16),
                                   r.x, r.y);
    st[3 * n_pad + i] = p4;
  } else {
    st[i] = p0;
    st[n_pad + i] = make_uint4((uint32_t)s.tr[0], (uint32_t)s.tr[1], (uint32_t)s.tr[2], (uint32_t)s.tr[3]);
    st[2 * n_pad + i] = make_uint4((uint32_t)s.fl[0], (uint32_t)s.fl[1], (uint32_t)s.fl[2], (uint32_t)s.fl[3]);
    st[3 * n_pad + i] = make_uint4((uint32_t)s.e, (uint32_t)s.pn[0], (uint32_t)s.pn[1], (uint32_t)s.last_amt);
    st[4 * n_pad + i] = p4;
    st[5 * n_pad + i] = make_uint4(r.x, r.y, 0u, 0u);
    if (LAYOUT == LAYOUT_REPLAY) {
      const uint2 a0 = d2u(s.acc[0]), a1 = d2u(s.acc[1]), a2 = d2u(s.acc[2]), a3 = d2u(s.acc[3]);
      st[6 * n_pad + i] = make_uint4(a0.x, a0.y, a1.x, a1.y);
      st[7 * n_pad + i] = make_uint4(a2.x, a2.y, a3.x, a3.y);
    }
  }
}

// Runtime-dispatched load/store for the kernels off the hot path.  In PHILOX layouts accuracy_belt is
// recomputed from the previous step's counter and mode (baseline right after a reset, env_super.py:395).
__device__ __forceinline__ void load_env(const DevConfig& c, const uint4* __restrict__ st, long long i, Env& s) {
  if (c.layout == LAYOUT_COMPACT) load_planes<LAYOUT_COMPACT>(st, c.n_pad, i, s);
  else if (c.layout == LAYOUT_WIDE) load_planes<LAYOUT_WIDE>(st, c.n_pad, i, s);
  else load_planes<LAYOUT_REPLAY>(st, c.n_pad, i, s);
  if (c.layout != LAYOUT_REPLAY) {
    if (s.step == 0) {
#pragma unroll
      for (int m = 0; m < 4; ++m) s.acc[m] = c.base_acc[m];
    } else {
      const unsigned long long g = (unsigned long long)(c.gid0 + i);
      philox_accuracy(c, (uint32_t)g, (uint32_t)(g >> 32) & 0xffffu, s.episode, s.step - 1, s.mode, s.acc);
    }
  }
}

__device__ __forceinline__ void store_env(const DevConfig& c, uint4* __restrict__ st, long long i, const Env& s) {
  if (c.layout == LAYOUT_COMPACT) store_planes<LAYOUT_COMPACT>(st, c.n_pad, i, s);
  else if (c.layout == LAYOUT_WIDE) store_planes<LAYOUT_WIDE>(st, c.n_pad, i, s);
  else store_planes<LAYOUT_REPLAY>(st, c.n_pad, i, s);
}

__device__ __forceinline__ int level_of(const Env& s, int m) {
  int l = s.e;
#pragma unroll
  for (int k = 0; k < 4; ++k) if (m == k) l = s.tr[k] + s.fl[k];
  return l;
}

// ref: press_action_masks env_super.py:869-885 (11 information bits)
__device__ __forceinline__ uint32_t press_mask_bits(const DevConfig& c, const Env& s) {
  uint32_t ready = 0;
#pragma unroll
  for (int m = 0; m < 4; ++m) ready |= (uint32_t)(s.tr[m] + s.fl[m] >= c.S) << m;
  ready |= (uint32_t)(s.e >= c.S) << 4;
  uint32_t b = 1u;
  if (s.timer[0] == 0) b |= ready << 1;
  if (s.timer[1] == 0) b |= ready << 6;
  return b;
}

// ref: validate_press_action env_super.py:811-836
__device__ __forceinline__ bool press_action_valid(const DevConfig& c, const Env& s, int pa) {
  return (press_mask_bits(c, s) >> pa) & 1u;
}

// ---------------------------------------------------------------- 2-decimal purity, exactly
// k = rint(fl(fl(tr/tot)*100)) as the reference computes round(true/total, 2)*100
// (env_super.py:754,789).  For tot < 2^22 the float64 pipeline is within 2.3e-14 of the exact
// value 100*tr/tot while a non-tie is at least 1/(2*tot) away from a rounding boundary, so
// exact integer rounding gives the same k; on an exact .5 tie the float64 pipeline itself
// decides (its result depends on how tr/tot rounds).  The integer rounding is found with one
// float32 approximate quotient and an exact integer correction.
static __device__ __noinline__ int purity_k_f64(int tr, int tot) {
  return __double2int_rn(dmul(ddiv((double)tr, (double)tot), 100.0));
}

// On an exact tie 100*tr/tot = k0 + 1/2 the quotient tr/tot equals (2*k0+1)/200 whatever tr and tot
// are, so fl(tr/tot), fl(.*100) and the final rint depend on k0 alone: the host evaluates the
// float64 pipeline once per k0 = 0..99 and hands the outcomes over as the bit table c.tie_up.
// SMALL: the caller knows c.small_lv holds (compile-time copy of the flag).  LARGE16: the caller knows it does
// NOT hold but tot < 2^16 (compact layout): the correction step is taken without asking.
template <bool SMALL = false, bool LARGE16 = false>
__device__ __forceinline__ int purity_k(const DevConfig& c, int tr, int tot) {  // tot > 0, 0 <= tr <= tot
  if (SMALL || LARGE16 || tot < (1 << 17)) {   // 100*tr and tot are exact in float32
    const int a2 = 200 * tr, t2 = 2 * tot;
    int k = __float2int_rn(__fdividef((float)(100 * tr), (float)tot));  // within 1 of the exact rounding
    int d = a2 - k * t2;                       // exact: twice the signed distance to k, in units of 1/tot
    // __fdividef is within 2 ulp (2.4e-5 at a quotient of 100) while a non-tie sits at least 1/(2*tot)
    // from a rounding boundary: for tot <= 8192 (c.small_lv) k is already the exact rounding
    if (!SMALL && (LARGE16 || !c.small_lv)) {
      if (d > tot) { k += 1; d -= t2; } else if (d < -tot) { k -= 1; d += t2; }
    }
    if (abs(d) == tot) {                       // .5 tie: k0 = floor of the exact value (0..99), outcome from the table
      const int k0 = k - (d < 0 ? 1 : 0);
      k = k0 + (int)((c.tie_up[k0 >> 5] >> (k0 & 31)) & 1u);
    }
    return k;                                  // not a tie: exact rounding == float64 pipeline
  }
  return purity_k_f64(tr, tot);
}

// Sorting reward tanh(((sum_m(p_m - theta))/4*scaling)/T) (calculate_sorting_reward env_super.py:963-1003).
// With thresholds that are whole percents (fast_pdiff) the argument depends only on the integer
// purity sum kt = sum_m (k_m, or 100*qthr_m for an empty container), so the reward is a float64
// table built on the host with the same formula; otherwise the float64 formula is evaluated here.
static __device__ __noinline__ double sort_reward_f64(const DevConfig& c, int k0, int k1, int k2, int k3) {
  int ksum = 0;
  double extra = 0.0;
  if (k0 >= 0) ksum += k0; else extra += c.qthr_empty[0];
  if (k1 >= 0) ksum += k1; else extra += c.qthr_empty[1];
  if (k2 >= 0) ksum += k2; else extra += c.qthr_empty[2];
  if (k3 >= 0) ksum += k3; else extra += c.qthr_empty[3];
  return tanh(((double)ksum * 0.01 + (extra - c.theta4)) * c.c_sort);
}

// Observation value of one purity difference: round(purity - threshold, 2)
// (compute_purity_differences env_super.py:212-227) for a non-empty container with purity k/100.
static __device__ __noinline__ float pdiff_f64(int k, double qthr) {
  return (float)round2(dsub(ddiv((double)k, 100.0), qthr));
}

// ---------------------------------------------------------------- observations (float32 out)
// The observation row is assembled piecewise so the step kernel can write each group as soon as
// its inputs are final (keeps them out of registers).
// ref: get_sort_obs env_super.py:306-325 = [belt_occ, belt proportions x4, accuracy_belt x4, purity diffs x4]
__device__ __forceinline__ void obs_belt(const Env& s, float* o) {  // o[0..4]; compute_belt_proportions :199-210
  const int bt = sum4(s.belt4);
  const float inv_bt = bt > 0 ? __frcp_rn((float)bt) : 0.f;
  o[0] = fminf((float)bt * 0.01f, 1.f);
#pragma unroll
  for (int m = 0; m < 4; ++m) o[1 + m] = fminf((float)b4(s.belt4, m) * inv_bt, 1.f);
}

__device__ __forceinline__ void obs_acc(const double acc[4], float* o) {  // o[5..8]; already clipped to [0,1]
#pragma unroll
  for (int m = 0; m < 4; ++m) o[5 + m] = (float)acc[m];
}

// kq[m] = purity k (0..100) of container m, or -1 when the container is empty (purity == round(threshold, 2))
__device__ __forceinline__ void obs_pdiff(const DevConfig& c, const int kq[4], float* o) {  // o[9..12]
#pragma unroll
  for (int m = 0; m < 4; ++m) {
    float d = c.pdiff_empty[m];
    if (kq[m] >= 0) d = c.fast_pdiff ? (float)(kq[m] - c.qthr100[m]) * 0.01f : pdiff_f64(kq[m], c.qthr[m]);
    o[9 + m] = clipf(d, -1.f, 1.f);
  }
}

__device__ __forceinline__ void sort_obs(const DevConfig& c, const Env& s, const int kq[4], float* o) {
  obs_belt(s, o);
  obs_acc(s.acc, o);
  obs_pdiff(c, kq, o);
}

// ref: get_press_obs env_super.py:327-359 = [levels x5, levels x5 again, sorting/stage x4, timers x2]
__device__ __forceinline__ void obs_levels_timers(const DevConfig& c, const Env& s, float* o) {  // o[0..9], o[14..15]
#pragma unroll
  for (int m = 0; m < 5; ++m) {
    const int l = m < 4 ? s.tr[m] + s.fl[m] : s.e;
    const float v = fminf((float)l * c.inv_cap, 1.f);
    o[m] = v; o[5 + m] = v;
  }
  o[14] = fminf((float)s.timer[0] * c.inv_pt[0], 1.f);
  o[15] = fminf((float)s.timer[1] * c.inv_pt[1], 1.f);
}

__device__ __forceinline__ void obs_sorting(const DevConfig& c, const Env& s, float* o) {  // o[10..13]
#pragma unroll
  for (int m = 0; m < 4; ++m) o[10 + m] = fminf((float)b4(s.sort4, m) * c.inv_stage, 1.f);
}

__device__ __forceinline__ void press_obs(const DevConfig& c, const Env& s, float* o) {
  obs_levels_timers(c, s, o);
  obs_sorting(c, s, o);
}

__device__ __forceinline__ void purity_ks(const DevConfig& c, const Env& s, int kq[4]) {
#pragma unroll
  for (int m = 0; m < 4; ++m) {
    int tot = s.tr[m] + s.fl[m];
    kq[m] = tot > 0 ? purity_k(c, s.tr[m], tot) : -1;
  }
}

template <int KIND>
__device__ __forceinline__ void env_obs(const DevConfig& c, const Env& s, float* o) {
  if (KIND != MSORT_ENV_PRESS) {
    int kq[4];
    purity_ks(c, s, kq);
    sort_obs(c, s, kq, o);
  }
  if (KIND == MSORT_ENV_PRESS) press_obs(c, s, o);
  if (KIND == MSORT_ENV_MONO) press_obs(c, s, o + 13);
}

// fresh plant (ref: Env_Super.reset env_super.py:365-420); keeps what the caller sets after
__device__ __forceinline__ void reset_env(const DevConfig& c, Env& s) {
  s.in4 = s.belt4 = s.sort4 = 0;
#pragma unroll
  for (int m = 0; m < 4; ++m) { s.tr[m] = 0; s.fl[m] = 0; s.acc[m] = c.base_acc[m]; }
  s.e = 0;
  s.timer[0] = s.timer[1] = s.mat[0] = s.mat[1] = s.pn[0] = s.pn[1] = s.pq[0] = s.pq[1] = 0;
  s.started = 0; s.last_amt = 0; s.gidx = 0; s.gcount = 0; s.mode = 0; s.step = 0; s.ep_ret = 0.0;
}

__device__ __forceinline__ void zero_cold(uint4* __restrict__ st, long long n_pad, long long i) {
#pragma unroll
  for (int m = 0; m < kColdPlanes; ++m) st[(kHotPlanes + m) * n_pad + i] = make_uint4(0u, 0u, 0u, 0u);
}

// ref: press_bale env_super.py:661-687 — updates the material's cold plane in place
static __device__ __noinline__ int press_bale(const DevConfig& c, uint4* __restrict__ st, long long n_pad,
                                       long long i, int m, int n, int qk) {
  uint4* cell = &st[(kHotPlanes + m) * n_pad + i];
  uint4 v = *cell;
  uint32_t cnt = v.x, sum = v.y, last_size = v.z & 0xffffffu, last_q = v.z >> 24;
  double q = ddiv((double)qk, 100.0);
  uint32_t q100 = (uint32_t)__double2int_rz(dmul(q, 100.0));  // int(q*100): truncation (:663)
  int S = c.S, full = n / S, rem = n - full * S, made = 0;
  if (full > 0) { cnt += full; sum += (uint32_t)(full * S); last_size = (uint32_t)S; last_q = q100; made += full; }
  if (rem > 0) {
    if (rem >= c.rem_new_bale || cnt == 0) { cnt += 1; last_size = (uint32_t)rem; last_q = q100; made += 1; }
    else last_size += (uint32_t)rem;
    sum += (uint32_t)rem;
  }
  *cell = make_uint4(cnt, sum, (last_size & 0xffffffu) | (last_q << 24), 0u);
  return made;
}

// tanh(x) = 1 - 2/(exp(2x)+1) with the hardware ex2/rcp approximations: absolute error ~1e-7
// (float32 epsilon), i.e. at the level of the fp32 reference's own rounding — the argmax of the
// policy can only differ from the reference inside its numerical tie band.
__device__ __forceinline__ float tanh_fast(float x) {
  const float e = __expf(2.0f * x);
  return 1.0f - __fdividef(2.0f, e + 1.0f);
}

// fp32 MLP 13->32->32->2 with tanh, fully unrolled (ref: sort_agent.predict env_2_press.py:106-109; arch
// training.py:115).  `w` is the kernel-parameter copy of the weights in the PAIRED layout pack_policy_pairs()
// builds (msort_kernels.cu): the weights of output neurons 2p and 2p+1 for the same input sit side by side, so
// one Blackwell packed FMA (`fma.rn.f32x2`, SASS FFMA2: two independent IEEE fp32 FMAs per lane and issue
// slot — the 3-register FFMA only reaches half of the fp32 pipe) advances both neurons.  Its weight pair is a
// uniform-register operand straight from the constant bank (LDCU.128 feeds two FFMA2), its input the scalar
// register broadcast to both halves.  Every neuron still sums bias + w_0 x_0 + w_1 x_1 + ... in the same
// order with the same roundings as the scalar loop, so the logits are bit-identical to it.
// acc + w * (x, x) on both halves; one asm statement per FFMA2 (the pair (x, x) is folded into the instruction's
// scalar-broadcast operand form by ptxas).  Inline PTX rather than the `__ffma2_rn` builtin on purpose: with the
// builtin the compiler interleaves all sixteen accumulator chains and spills (measured 199 us vs 154 us per
// 1 048 576 envs); the opaque statements keep the chains in program order.
__device__ __forceinline__ unsigned long long ffma2_bcast(unsigned long long w, float x, unsigned long long acc) {
  unsigned long long d;
  asm("{\n\t.reg .b64 xx;\n\tmov.b64 xx, {%2, %2};\n\tfma.rn.f32x2 %0, %1, xx, %3;\n\t}" : "=l"(d) : "l"(w), "f"(x), "l"(acc));
  return d;
}
__device__ __forceinline__ void f2unpack(unsigned long long v, float& lo, float& hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}

// PACKED = the FFMA2 form (used by the HOT instantiation); otherwise the same sums with scalar FFMA — identical
// roundings, identical logits — which keeps the seven colder Env_2 instantiations cheap to compile (cicc spends
// ~20 s per kernel on the ~800 asm statements of the packed form).
template <bool PACKED>
__device__ __forceinline__ int mlp_sort_mode(const float (&w)[(MSORT_POLICY_WEIGHTS + 3) / 4 * 4], const float* x) {
  constexpr int W1 = 0, b1 = 416, W2 = 448, b2 = 1472, W3 = 1504, b3 = 1568;
  float h1[32];
  if (PACKED) {
    const unsigned long long* const w2 = reinterpret_cast<const unsigned long long*>(w);   // pair p = floats 2p, 2p+1
#pragma unroll
    for (int jp = 0; jp < 16; ++jp) {
      unsigned long long a = w2[b1 / 2 + jp];
#pragma unroll
      for (int k = 0; k < 13; ++k) a = ffma2_bcast(w2[W1 / 2 + jp * 13 + k], x[k], a);
      float lo, hi;
      f2unpack(a, lo, hi);
      h1[2 * jp] = tanh_fast(lo); h1[2 * jp + 1] = tanh_fast(hi);
    }
    unsigned long long l = w2[b3 / 2];   // (logit 0, logit 1)
#pragma unroll
    for (int jp = 0; jp < 16; ++jp) {
      unsigned long long a = w2[b2 / 2 + jp];
#pragma unroll
      for (int k = 0; k < 32; ++k) a = ffma2_bcast(w2[W2 / 2 + jp * 32 + k], h1[k], a);
      float lo, hi;
      f2unpack(a, lo, hi);
      l = ffma2_bcast(w2[W3 / 2 + 2 * jp], tanh_fast(lo), l);
      l = ffma2_bcast(w2[W3 / 2 + 2 * jp + 1], tanh_fast(hi), l);
    }
    float l0, l1;
    f2unpack(l, l0, l1);
    return l1 > l0 ? 1 : 0;
  }
#pragma unroll
  for (int j = 0; j < 32; ++j) {       // neuron j = half (j & 1) of pair j / 2
    float a = w[b1 + j];
#pragma unroll
    for (int k = 0; k < 13; ++k) a = fmaf(w[W1 + ((j >> 1) * 13 + k) * 2 + (j & 1)], x[k], a);
    h1[j] = tanh_fast(a);
  }
  float l0 = w[b3], l1 = w[b3 + 1];
#pragma unroll
  for (int j = 0; j < 32; ++j) {
    float a = w[b2 + j];
#pragma unroll
    for (int k = 0; k < 32; ++k) a = fmaf(w[W2 + ((j >> 1) * 32 + k) * 2 + (j & 1)], h1[k], a);
    const float h = tanh_fast(a);
    l0 = fmaf(w[W3 + 2 * j], h, l0);
    l1 = fmaf(w[W3 + 2 * j + 1], h, l1);
  }
  return l1 > l0 ? 1 : 0;
}

}  // namespace msort
