// msort_kernels.cu — hand-written sm_100a kernels: fused step (K1), reset (K2), observe,
// state export/import (K6), state statistics (K5).  One thread = one env; one CTA = one tile
// of 128 consecutive envs.  Row-major obs[N,D] / mask[N,A] are transposed through shared
// memory so every global store is a coalesced 4-byte-per-lane (128 B per warp) store.
#include <cuda_runtime.h>

#include "msort_device.cuh"
#include "msort_launch.h"

namespace msort {

template <int KIND> struct Dims;
template <> struct Dims<MSORT_ENV_SORT> { static constexpr int D = 13, DP = 13, A = 2; };
template <> struct Dims<MSORT_ENV_PRESS> { static constexpr int D = 16, DP = 17, A = 11; };
template <> struct Dims<MSORT_ENV_MONO> { static constexpr int D = 29, DP = 29, A = 22; };

// ---------------------------------------------------------------- tile output helpers
// Copy the CTA's obs tile (smem, row stride DP) to global rows [row0, row0+rows) of obs[N,D].
template <int D, int DP>
__device__ __forceinline__ void flush_obs_tile(const float* __restrict__ tile, float* __restrict__ obs,
                                               long long row0, int rows) {
  float* dst = obs + row0 * D;
  const int total = rows * D;
  for (int e = threadIdx.x; e < total; e += kTile) {
    int r = e / D, k = e - r * D;
    dst[e] = tile[r * DP + k];
  }
}

// Expand per-env 11-bit press masks (smem) into mask[N,A] bytes, 4 bytes per lane per store.
template <int A>
__device__ __forceinline__ void flush_mask_tile(const uint16_t* __restrict__ bits, uint8_t* __restrict__ mask,
                                                long long row0, int rows) {
  uint8_t* dst = mask + row0 * A;  // row0 is a multiple of 128 -> 4-byte aligned for every A
  const int total = rows * A;
  const int words = total >> 2;
  for (int w = threadIdx.x; w < words; w += kTile) {
    uint32_t v = 0;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      int e = 4 * w + q;
      int r = e / A, k = e - r * A;
      uint32_t bit = A == 2 ? 1u : (bits[r] >> (k >= 11 ? k - 11 : k)) & 1u;
      v |= bit << (8 * q);
    }
    reinterpret_cast<uint32_t*>(dst)[w] = v;
  }
  for (int e = 4 * words + threadIdx.x; e < total; e += kTile) {
    int r = e / A, k = e - r * A;
    dst[e] = A == 2 ? 1 : (uint8_t)((bits[r] >> (k >= 11 ? k - 11 : k)) & 1u);
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

constexpr int kNumAcc = 10;  // stats slots accumulated by the step kernel

template <int KIND, int RNG>
__global__ void __launch_bounds__(kTile)
step_kernel(const __grid_constant__ DevConfig c, const __grid_constant__ StepArgs a) {
  constexpr int D = Dims<KIND>::D, DP = Dims<KIND>::DP, A = Dims<KIND>::A;
  __shared__ float s_obs[kTile * DP];
  __shared__ uint16_t s_bits[kTile];
  __shared__ float s_policy[KIND == MSORT_ENV_PRESS ? MSORT_POLICY_WEIGHTS : 1];
  __shared__ double s_acc[kNumAcc][kTile / 32];

  const bool masking = c.flags & MSORT_F_ACTION_MASKING;
  const bool auto_reset = c.flags & MSORT_F_AUTO_RESET;
  const bool use_mlp = KIND == MSORT_ENV_PRESS && (c.flags & MSORT_F_SORT_POLICY_MLP) &&
                       !(RNG == MSORT_RNG_REPLAY && a.sort_mode_in);
  if (use_mlp) {
    for (int k = threadIdx.x; k < MSORT_POLICY_WEIGHTS; k += kTile) s_policy[k] = c.policy[k];
    __syncthreads();
  }

  const long long row0 = (long long)blockIdx.x * kTile;
  const long long i = row0 + threadIdx.x;
  const bool live = i < c.n;
  const int rows = (int)min((long long)kTile, c.n - row0);
  const long long gid = c.gid0 + i;

  double acc_stat[kNumAcc];
#pragma unroll
  for (int k = 0; k < kNumAcc; ++k) acc_stat[k] = 0.0;

  if (live) {
    Env s;
    load_env(a.state, c.n_pad, i, s);
    long long act = a.actions[i];
    const uint32_t ep = s.episode, stp = s.step;

    if (act < 0) { act = 0; acc_stat[8] += 1.0; }
    if (act >= A) { act = A - 1; acc_stat[8] += 1.0; }

    // 1: material flow (update_environment env_super.py:440-442)
    s.sort4 = s.belt4; s.belt4 = s.in4;
    // 2: seasonal generator (input_generator.py:37-64); counts only
    if (RNG == MSORT_RNG_REPLAY && a.input_counts) {
      s.in4 = a.input_counts[i];
    } else {
      if (s.gcount >= c.spp) { s.gidx ^= 1; s.gcount = 0; }
      s.in4 = c.pat[s.gidx ^ s.gfirst];
      if (c.pat_remainder > 0) {
        U4 r4 = {0, 0, 0, 0};
        for (int k = 0; k < c.pat_remainder; ++k) {
          if ((k & 3) == 0) r4 = env_draw(c, gid, kBlkInput + 0x100u * (uint32_t)(k >> 2), ep, stp);
          s.in4 += 1u << (8 * (u4_get(r4, k & 3) & 3u));
        }
      }
      s.gcount += 1;
    }
    double acc_sorter[4];
#pragma unroll
    for (int m = 0; m < 4; ++m) acc_sorter[m] = s.acc[m];  // env_super.py:457

    // 3: decode the action
    int mode = 0, pa = 0;
    bool skip_press = false, invalid = false;
    if (KIND == MSORT_ENV_SORT) {
      mode = (int)act;
    } else if (KIND == MSORT_ENV_MONO) {
      mode = (int)act / 11; pa = (int)act - 11 * mode;
      if (!masking && !press_action_valid(c, s, pa)) { pa = 0; skip_press = true; invalid = true; }
    } else {
      pa = (int)act;
      if (RNG == MSORT_RNG_REPLAY && a.sort_mode_in) {
        mode = a.sort_mode_in[i] & 1;
      } else if (use_mlp) {
        float so[13];
        double pur[4];
        container_purity(c, s, pur);
        sort_obs(c, s, pur, so);
        mode = mlp_sort_mode(s_policy, so);
      } else {  // sorting_rules env_super.py:469-482 (float64 proportions, as the reference)
        int bt = sum4(s.belt4);
        double p[4];
#pragma unroll
        for (int m = 0; m < 4; ++m) p[m] = bt > 0 ? ddiv((double)b4(s.belt4, m), (double)bt) : 0.0;
        mode = dadd(p[0], p[2]) > dadd(p[1], p[3]) ? 0 : 1;
      }
    }
    s.mode = mode;

    // 4: update_accuracy env_super.py:492-509
    {
      double u[4];
      if (RNG == MSORT_RNG_REPLAY) {
        const double2* nz = reinterpret_cast<const double2*>(a.noise_u) + 2 * i;
        double2 n0 = nz[0], n1 = nz[1];
        u[0] = n0.x; u[1] = n0.y; u[2] = n1.x; u[3] = n1.y;
      } else {
        U4 r4 = env_draw(c, gid, kBlkNoise, ep, stp);
        u[0] = (double)r4.x * 2.3283064365386963e-10; u[1] = (double)r4.y * 2.3283064365386963e-10;
        u[2] = (double)r4.z * 2.3283064365386963e-10; u[3] = (double)r4.w * 2.3283064365386963e-10;
      }
#pragma unroll
      for (int m = 0; m < 4; ++m) {
        double base = c.base_acc[m];
        if ((m & 1) == mode) base = dadd(base, c.boost);
        double nzv = dadd(c.noise_low, dmul(c.noise_range, u[m]));
        s.acc[m] = clipd(dadd(base, nzv), 0.0, 1.0);
      }
    }

    // 5: sort_material env_super.py:511-609 — one flattened loop over all redistribution draws
    {
      uint32_t L = s.sort4, T4 = 0, F4 = 0;
      int m = 0, rem = 0, k = 0;
      U4 r4 = {0, 0, 0, 0};
      while (true) {
        while (rem == 0 && m < 4) {
          int t = b4(L, m);
          double am = m == 0 ? acc_sorter[0] : (m == 1 ? acc_sorter[1] : (m == 2 ? acc_sorter[2] : acc_sorter[3]));
          int tv = __double2int_rn(dmul((double)t, am));  // int(round(t*acc)) half-to-even (:539)
          int f = t - tv;
          T4 |= (uint32_t)tv << (8 * m); F4 |= (uint32_t)f << (8 * m);
          L = (L & ~(0xffu << (8 * m))) | ((uint32_t)f << (8 * m));
          rem = f; ++m;
        }
        if (rem == 0) break;
        int tot = sum4(L);
        if (tot == 0) { rem = 0; continue; }  // :557-559
        int j;
        if (RNG == MSORT_RNG_REPLAY) {
          if ((long long)s.cursor >= a.redis_len) {
            acc_stat[9] += 1.0;
            j = 0; while (b4(L, j) == 0) ++j;  // defined fallback; reported as an under-run
          } else {
            double uu = a.redis_u[i * a.redis_len + s.cursor];
            s.cursor += 1;
            double cdf[4], cs = 0.0;  // numpy Generator.choice(4, p=): cumsum, /= last, searchsorted right
#pragma unroll
            for (int q = 0; q < 4; ++q) { cs = dadd(cs, ddiv((double)b4(L, q), (double)tot)); cdf[q] = cs; }
            j = 0;
#pragma unroll
            for (int q = 0; q < 3; ++q) j += ddiv(cdf[q], cs) <= uu ? 1 : 0;
          }
        } else {
          if ((k & 3) == 0) r4 = env_draw(c, gid, kBlkRedis + (uint32_t)(k >> 2), ep, stp);
          uint32_t x = u4_get(r4, k & 3);
          uint32_t r = __umulhi(x, (uint32_t)tot);
          uint32_t pre = L * 0x01010101u;  // byte q = L0+..+Lq (tot <= 255: no carries)
          j = (r >= (pre & 0xffu)) + (r >= ((pre >> 8) & 0xffu)) + (r >= ((pre >> 16) & 0xffu));
        }
        ++k;
        L -= 1u << (8 * j);
        --rem;
      }
      s.e += sum4(L);
#pragma unroll
      for (int q = 0; q < 4; ++q) { s.tr[q] += b4(T4, q); s.fl[q] += b4(F4, q); }
    }

    // 6: Env_1 samples its own press action under the mask (env_super.py:291-300)
    if (KIND == MSORT_ENV_SORT) {
      if (RNG == MSORT_RNG_REPLAY) pa = a.press_choice[i];
      else {
        uint32_t vb = press_mask_bits(c, s);
        U4 r4 = env_draw(c, gid, kBlkPress, ep, stp);
        int pick = (int)__umulhi(r4.x, (uint32_t)__popc(vb));
        for (int q = 0; q < pick; ++q) vb &= vb - 1;  // drop the `pick` lowest valid actions
        pa = __ffs(vb) - 1;
      }
      if (pa > 10) pa = 0;
    }
    if (KIND == MSORT_ENV_PRESS && !masking && !press_action_valid(c, s, pa)) { pa = 0; invalid = true; }

    // 7: press_action_rules env_super.py:626-640
    int bales_made = 0;
    if (!skip_press) {
#pragma unroll
      for (int p = 0; p < 2; ++p) {  // check_press_status :642-659
        if (s.timer[p] > 0) {
          s.timer[p] -= 1;
          if (s.timer[p] == 0) {
            bales_made += press_bale(c, a.state, c.n_pad, i, s.mat[p], s.pn[p], s.pq[p]);
            s.mat[p] = 0; s.pn[p] = 0; s.pq[p] = 0;
          }
        }
      }
      if (pa != 0) {  // use_press :722-769
        int p = pa <= 5 ? 0 : 1, m = pa - 1 - 5 * p;
        if (s.timer[p] == 0) {
          int amt = level_of(s, m);
          s.started = 1; s.last_amt = amt;
          int qk = 0;
          if (m < 4 && amt > 0) {
            int tv = s.tr[0];
#pragma unroll
            for (int q = 1; q < 4; ++q) if (m == q) tv = s.tr[q];
            qk = __double2int_rn(dmul(ddiv((double)tv, (double)amt), 100.0));  // round(true/total, 2) (:754)
          }
#pragma unroll
          for (int q = 0; q < 4; ++q) if (m == q) { s.tr[q] = 0; s.fl[q] = 0; }
          if (m == 4) s.e = 0;
          s.timer[p] = c.press_time[p]; s.mat[p] = m; s.pn[p] = amt; s.pq[p] = qk;
        }
      }
    }

    // 8: overflow termination (detect_overflow :900-905)
    int overflow_mat = -1;
    if (c.flags & MSORT_F_CHECK_OVERFLOW) {
#pragma unroll
      for (int m = 4; m >= 0; --m) if (level_of(s, m) > c.cap) overflow_mat = m;
    }
    const bool overflow = overflow_mat >= 0;

    // 9: rewards
    double reward;
    bool terminated;
    double pur[4];
    if (KIND != MSORT_ENV_PRESS) container_purity(c, s, pur);
    if (overflow) {
      reward = c.ovf_pen;
      s.step += 1;
      terminated = true;
    } else {
      double r_sort = 0.0, r_press = 0.0;
      if (KIND != MSORT_ENV_PRESS) {  // calculate_sorting_reward :963-1003
        double total = 0.0;
#pragma unroll
        for (int m = 0; m < 4; ++m) total = dadd(total, dsub(pur[m], c.theta));
        double sb = dmul(ddiv(total, 4.0), c.scaling);
        r_sort = tanh(ddiv(sb, c.temperature));
      }
      if (KIND != MSORT_ENV_SORT) {  // calculate_press_reward :1006-1080
        int mx = s.e, tl = s.e;
        bool sev = s.e >= c.lvl_sev, mild = s.e >= c.lvl_mild && s.e < c.lvl_sev;
#pragma unroll
        for (int m = 0; m < 4; ++m) {
          int l = s.tr[m] + s.fl[m];
          mx = max(mx, l); tl += l;
          sev |= l >= c.lvl_sev; mild |= l >= c.lvl_mild && l < c.lvl_sev;
        }
        double max_pen = 0.0;  // min(0, severe if any fill>0.95, mild if any fill in (0.90,0.95]) (:1024-1027)
        if (sev && c.pen_sev < max_pen) max_pen = c.pen_sev;
        if (mild && c.pen_mild < max_pen) max_pen = c.pen_mild;
        if (mx >= c.lvl_cat) r_press = c.pen_cat;            // :1022-1023
        else if (max_pen < 0.0) r_press = max_pen;           // :1029-1030 (started flag NOT cleared)
        else {
          double state_reward = dmul(ddiv((double)tl, (double)(5 * c.cap)), c.max_state);
          double action_reward = 0.0;
          if (s.started) {
            int S = c.S, amount = s.last_amt, nb = amount / S, rm = amount - nb * S;
            int d = min(rm, S - rm);
            double eff = dmul(dsub(1.0, dmul(4.0, ddiv((double)d, (double)S))), c.bef);
            double peak = nb == 0 ? 0.0 : (nb == 1 ? 1.0 / 3.0 : (nb == 2 ? 2.0 / 3.0 : 1.0));
            action_reward = dadd(eff, dsub(peak, c.bef));
            s.started = 0; s.last_amt = 0;
          }
          r_press = clipd(dadd(state_reward, action_reward), -1.0, 1.0);
        }
      }
      reward = KIND == MSORT_ENV_SORT ? r_sort : (KIND == MSORT_ENV_PRESS ? r_press : dadd(r_sort, r_press));
      s.step += 1;
      terminated = s.step >= (uint32_t)c.max_steps;
    }
    s.ep_ret = dadd(s.ep_ret, reward);

    // 10: observation (private row of the shared tile), outputs, auto-reset
    float* orow = &s_obs[threadIdx.x * DP];
    if (KIND != MSORT_ENV_PRESS) sort_obs(c, s, pur, orow);
    if (KIND == MSORT_ENV_PRESS) press_obs(c, s, orow);
    if (KIND == MSORT_ENV_MONO) press_obs(c, s, orow + 13);

    a.reward[i] = (float)reward;
    a.terminated[i] = terminated ? 1 : 0;
    if (a.info_action) a.info_action[i] = act;
    if (a.info_overflow) a.info_overflow[i] = overflow ? 1 : 0;
    if (a.info_overflow_mat) a.info_overflow_mat[i] = (int8_t)overflow_mat;
    if (a.info_sort_mode) a.info_sort_mode[i] = (uint8_t)mode;
    if (a.info_press_action) a.info_press_action[i] = (uint8_t)pa;
    if (a.info_invalid) a.info_invalid[i] = invalid ? 1 : 0;
    acc_stat[3] = 1.0; acc_stat[4] = reward; acc_stat[5] = overflow ? 1.0 : 0.0;
    acc_stat[6] = (double)bales_made; acc_stat[7] = invalid ? 1.0 : 0.0;
    if (terminated) {
      acc_stat[0] = 1.0; acc_stat[1] = s.ep_ret; acc_stat[2] = (double)s.step;
      if (a.episode_return) a.episode_return[i] = s.ep_ret;
      if (a.episode_length) a.episode_length[i] = (int)s.step;
      if (auto_reset) {
        if (a.terminal_obs) {
          float* t = a.terminal_obs + i * D;
#pragma unroll
          for (int k = 0; k < D; ++k) t[k] = orow[k];
        }
        // unseeded reset (env_super.py:365-420): streams run on, generator re-seeded
        reset_env(c, s);
        s.episode = ep + 1;
        s.gfirst = (int)(env_draw(c, gid, kBlkReset, s.episode, 0u).x & 1u);
        zero_cold(a.state, c.n_pad, i);
        env_obs<KIND>(c, s, orow);
      }
    }
    s_bits[threadIdx.x] = (uint16_t)press_mask_bits(c, s);
    store_env(a.state, c.n_pad, i, s);
  }

  if (a.stats) {  // warp shuffle -> shared -> one atomic per CTA per statistic
#pragma unroll
    for (int k = 0; k < kNumAcc; ++k) {
      double v = acc_stat[k];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      if ((threadIdx.x & 31) == 0) s_acc[k][threadIdx.x >> 5] = v;
    }
  }
  __syncthreads();
  if (a.stats && threadIdx.x < kNumAcc) {
    double v = 0.0;
#pragma unroll
    for (int w = 0; w < kTile / 32; ++w) v += s_acc[threadIdx.x][w];
    if (v != 0.0) atomicAdd(&a.stats[threadIdx.x], v);
  }
  flush_obs_tile<D, DP>(s_obs, a.obs, row0, rows);
  if (a.mask) flush_mask_tile<A>(s_bits, a.mask, row0, rows);
}

// ---------------------------------------------------------------- K2: reset
template <int KIND>
__global__ void __launch_bounds__(kTile)
reset_kernel(const __grid_constant__ DevConfig c, uint4* __restrict__ state, const uint8_t* __restrict__ which,
             const uint8_t* __restrict__ first_pattern, float* __restrict__ obs, uint8_t* __restrict__ mask,
             uint32_t reset_flags) {
  constexpr int D = Dims<KIND>::D, A = Dims<KIND>::A;
  const long long i = (long long)blockIdx.x * kTile + threadIdx.x;
  if (i >= c.n) return;
  if (which && !which[i]) return;
  Env s;
  s.episode = 0; s.cursor = 0;
  if (reset_flags & MSORT_RESET_KEEP_STREAMS) {  // reset(seed=None): streams run on (env_super.py:377)
    load_env(state, c.n_pad, i, s);
    s.episode += 1;
  }
  reset_env(c, s);
  int fp = first_pattern ? first_pattern[i] : 0;
  if (fp == 1 || fp == 2) s.gfirst = fp - 1;
  else s.gfirst = (int)(env_draw(c, c.gid0 + i, kBlkReset, s.episode, 0u).x & 1u);
  zero_cold(state, c.n_pad, i);
  store_env(state, c.n_pad, i, s);
  if (obs) {
    float o[D];
    env_obs<KIND>(c, s, o);
#pragma unroll
    for (int k = 0; k < D; ++k) obs[i * D + k] = o[k];
  }
  if (mask) {
    uint32_t b = press_mask_bits(c, s);
#pragma unroll
    for (int k = 0; k < A; ++k) mask[i * A + k] = A == 2 ? 1 : (uint8_t)((b >> (k >= 11 ? k - 11 : k)) & 1u);
  }
}

// ---------------------------------------------------------------- observe (no transition)
template <int KIND>
__global__ void __launch_bounds__(kTile)
observe_kernel(const __grid_constant__ DevConfig c, const uint4* __restrict__ state, float* __restrict__ obs,
               uint8_t* __restrict__ mask) {
  constexpr int D = Dims<KIND>::D, DP = Dims<KIND>::DP, A = Dims<KIND>::A;
  __shared__ float s_obs[kTile * DP];
  __shared__ uint16_t s_bits[kTile];
  const long long row0 = (long long)blockIdx.x * kTile;
  const long long i = row0 + threadIdx.x;
  const int rows = (int)min((long long)kTile, c.n - row0);
  if (i < c.n) {
    Env s;
    load_env(state, c.n_pad, i, s);
    env_obs<KIND>(c, s, &s_obs[threadIdx.x * DP]);
    s_bits[threadIdx.x] = (uint16_t)press_mask_bits(c, s);
  }
  __syncthreads();
  if (obs) flush_obs_tile<D, DP>(s_obs, obs, row0, rows);
  if (mask) flush_mask_tile<A>(s_bits, mask, row0, rows);
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

// ---------------------------------------------------------------- K6: export / import
__global__ void __launch_bounds__(kTile)
export_kernel(const __grid_constant__ DevConfig c, const uint4* __restrict__ state, msort_env_state_t* __restrict__ out) {
  const long long i = (long long)blockIdx.x * kTile + threadIdx.x;
  if (i >= c.n) return;
  Env s;
  load_env(state, c.n_pad, i, s);
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
  out[i] = o;
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
  store_env(state, c.n_pad, i, s);
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
    load_env(state, c.n_pad, i, s);
    v[0] += 1.0;
    double lvl = (double)s.e, pm = 0.0;
    double pur[4];
    container_purity(c, s, pur);
#pragma unroll
    for (int m = 0; m < 4; ++m) { lvl += (double)(s.tr[m] + s.fl[m]); pm += pur[m]; }
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
static inline unsigned tiles(long long n) { return (unsigned)((n + kTile - 1) / kTile); }

template <int KIND>
static cudaError_t launch_step_kind(const DevConfig& c, const StepArgs& a, int rng, cudaStream_t st) {
  if (rng == MSORT_RNG_REPLAY) step_kernel<KIND, MSORT_RNG_REPLAY><<<tiles(c.n), kTile, 0, st>>>(c, a);
  else step_kernel<KIND, MSORT_RNG_PHILOX><<<tiles(c.n), kTile, 0, st>>>(c, a);
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
  a.terminal_obs = f ? f->terminal_obs : nullptr;
  a.episode_return = f ? f->episode_return : nullptr;
  a.episode_length = f ? f->episode_length : nullptr;
  a.stats = f ? f->stats : nullptr;
  const msort_replay_t* r = l.replay;
  a.noise_u = r ? r->noise_u : nullptr; a.redis_u = r ? r->redis_u : nullptr;
  a.redis_len = r ? r->redis_len : 0; a.input_counts = r ? r->input_counts : nullptr;
  a.press_choice = r ? r->press_choice : nullptr; a.sort_mode_in = r ? r->sort_mode : nullptr;
  switch (c.kind) {
    case MSORT_ENV_SORT: return launch_step_kind<MSORT_ENV_SORT>(c, a, rng, st);
    case MSORT_ENV_PRESS: return launch_step_kind<MSORT_ENV_PRESS>(c, a, rng, st);
    default: return launch_step_kind<MSORT_ENV_MONO>(c, a, rng, st);
  }
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

cudaError_t launch_observe(const DevConfig& c, const void* state, float* obs, uint8_t* mask, cudaStream_t st) {
  switch (c.kind) {
    case MSORT_ENV_SORT: observe_kernel<MSORT_ENV_SORT><<<tiles(c.n), kTile, 0, st>>>(c, (const uint4*)state, obs, mask); break;
    case MSORT_ENV_PRESS: observe_kernel<MSORT_ENV_PRESS><<<tiles(c.n), kTile, 0, st>>>(c, (const uint4*)state, obs, mask); break;
    default: observe_kernel<MSORT_ENV_MONO><<<tiles(c.n), kTile, 0, st>>>(c, (const uint4*)state, obs, mask); break;
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

cudaError_t launch_export(const DevConfig& c, const void* state, msort_env_state_t* out, cudaStream_t st) {
  export_kernel<<<tiles(c.n), kTile, 0, st>>>(c, (const uint4*)state, out);
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
