/*
 * msort_oracle.c — TEST INFRASTRUCTURE.  CPU restatement (plain C, scalar, float64 in the
 * reference's own operation order) of the step()/reset() dynamics of MARL-SortingEnv.
 *
 * This is the checker, not the product: only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may load it.  The product path
 * (marl-sortingenv_b200/) never links, imports or calls anything in oracle/.
 *
 * Parity pin: the reference ships no tests or golden vectors for this path
 * (SURVEY.md §4), so this restatement is pinned against OUTPUTS OF THE REFERENCE ITSELF,
 * recorded in the build container by oracle/ref_record.py from the unmodified classes in
 * /root/reference and committed under tests/golden/ (generator: tests/golden/make_golden.py).
 * tests/test_oracle_vs_golden.py replays every fixture through this file and requires
 * bit-identical integer state / masks / flags and float64-identical rewards (<=1e-12 rel).
 *
 * Third-party arithmetic on the path (numpy, not vendored in the reference): PCG64
 * streams are NOT re-implemented here — REPLAY mode consumes the uniforms numpy
 * produced; `Generator.choice(n, p=)` is restated from numpy's published
 * algorithm (cdf = cumsum(p); cdf /= cdf[-1]; searchsorted(u, side='right')), and
 * `round(np.float64, 2)` as rint(x*100)/100.
 *
 * Every block cites the reference lines it follows ("ref:" = path under /root/reference).
 * Build: see oracle/Makefile (gcc -O2 -ffp-contract=off: no FMA contraction, ever).
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>

#include "../include/msort.h"

/* ---------------------------------------------------------------- Philox4x32-10 */
/* Published algorithm (Salmon et al., SC'11, Random123).  Counter scheme shared with the
 * device kernels (DESIGN.md "Philox counter layout"):
 *   ctr = { gid_lo32, (gid_hi & 0xffff) | block<<16, episode, step },  key = seed. */
static void philox4x32_10(const uint32_t ctr_in[4], const uint32_t key_in[2], uint32_t out[4]) {
  uint32_t c0 = ctr_in[0], c1 = ctr_in[1], c2 = ctr_in[2], c3 = ctr_in[3];
  uint32_t k0 = key_in[0], k1 = key_in[1];
  for (int r = 0; r < 10; ++r) {
    uint64_t p0 = (uint64_t)0xD2511F53u * c0;
    uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
    uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
    uint32_t n1 = (uint32_t)p1;
    uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
    uint32_t n3 = (uint32_t)p0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

void mso_philox4x32_10(const uint32_t* ctr, const uint32_t* key, uint32_t* out) {
  philox4x32_10(ctr, key, out);
}

enum { BLK_NOISE = 0, BLK_PRESS = 1, BLK_RESET = 2, BLK_INPUT = 3, BLK_REDIS = 16 };

static void env_draw(const msort_config_t* cfg, int64_t gid, uint32_t block, uint32_t episode,
                     uint32_t step, uint32_t out[4]) {
  uint32_t ctr[4], key[2];
  uint64_t g = (uint64_t)gid;
  ctr[0] = (uint32_t)g;
  ctr[1] = (uint32_t)((g >> 32) & 0xffffu) | (block << 16);
  ctr[2] = episode;
  ctr[3] = step;
  key[0] = (uint32_t)cfg->seed;
  key[1] = (uint32_t)(cfg->seed >> 32);
  philox4x32_10(ctr, key, out);
}

/* ---------------------------------------------------------------- helpers */
static double round2(double x) { return rint(x * 100.0) / 100.0; } /* numpy-scalar round(x, 2) */
/* Python-float round(x, 2): correctly rounded on the exact decimal value of the double (float.__round__ goes
 * through dtoa), which is what glibc's "%.2f" prints.  The reference applies it where both operands are plain
 * Python floats: an EMPTY container's purity round(threshold, 2) (env_super.py:788-789) and that container's
 * purity difference round(purity - threshold, 2) (env_super.py:222-225). */
static double py_round2(double x) {
  char buf[64];
  snprintf(buf, sizeof buf, "%.2f", x);
  return strtod(buf, NULL);
}
static double clipd(double x, double lo, double hi) { return x < lo ? lo : (x > hi ? hi : x); }
static float clipf(float x, float lo, float hi) { return x < lo ? lo : (x > hi ? hi : x); }

static int kind_obs_dim(int kind) {
  return kind == MSORT_ENV_SORT ? 13 : (kind == MSORT_ENV_PRESS ? 16 : 29);
}
static int kind_num_actions(int kind) {
  return kind == MSORT_ENV_SORT ? 2 : (kind == MSORT_ENV_PRESS ? 11 : 22);
}

static int level_of(const msort_env_state_t* s, int m) {
  return m < 4 ? s->cont_true[m] + s->cont_false[m] : s->cont_e;
}

/* ref: get_container_purity, env_super.py:771-791 */
static void container_purity(const msort_config_t* cfg, const msort_env_state_t* s, double pur[4]) {
  for (int m = 0; m < 4; ++m) {
    int tot = s->cont_true[m] + s->cont_false[m];
    if (tot > 0) pur[m] = round2((double)s->cont_true[m] / (double)tot);
    else pur[m] = py_round2(cfg->quality_threshold[m]); /* Python float: round(0.9, 2) == 0.9, round(0.905, 2) == 0.91 */
  }
}

/* ref: compute_belt_proportions, env_super.py:199-210 */
static void belt_proportions(const msort_env_state_t* s, double prop[4]) {
  int tot = s->belt[0] + s->belt[1] + s->belt[2] + s->belt[3];
  for (int m = 0; m < 4; ++m) prop[m] = tot > 0 ? (double)s->belt[m] / (double)tot : 0.0;
}

/* ref: get_sort_obs, env_super.py:306-325 ; compute_purity_differences :212-227 */
static void sort_obs(const msort_config_t* cfg, const msort_env_state_t* s, float* o) {
  double prop[4], pur[4];
  belt_proportions(s, prop);
  container_purity(cfg, s, pur);
  int bt = s->belt[0] + s->belt[1] + s->belt[2] + s->belt[3];
  double v[13];
  v[0] = (double)bt / 100.0; /* belt_occupancy = previous round(sum(input)/100, 2): env_super.py:442,456 */
  for (int m = 0; m < 4; ++m) {
    v[1 + m] = prop[m];
    v[5 + m] = s->acc_belt[m];
    int tot = s->cont_true[m] + s->cont_false[m];
    double diff = pur[m] - cfg->quality_threshold[m];
    v[9 + m] = tot > 0 ? round2(diff) : py_round2(diff); /* empty: both Python floats (0.0 for whole-percent thresholds) */
  }
  for (int i = 0; i < 13; ++i) o[i] = clipf((float)v[i], -1.0f, 1.0f);
}

/* ref: get_press_obs, env_super.py:327-359 */
static void press_obs(const msort_config_t* cfg, const msort_env_state_t* s, float* o) {
  double v[16];
  for (int m = 0; m < 5; ++m) {
    v[m] = (double)level_of(s, m) / (double)cfg->container_capacity;
    v[5 + m] = v[m];
  }
  for (int m = 0; m < 4; ++m) v[10 + m] = (double)s->sorting[m] / (double)cfg->stage_capacity;
  for (int p = 0; p < 2; ++p) v[14 + p] = (double)s->press_timer[p] / (double)cfg->press_time[p];
  for (int i = 0; i < 16; ++i) o[i] = clipf((float)v[i], 0.0f, 1.0f);
}

static void write_obs(const msort_config_t* cfg, const msort_env_state_t* s, float* o) {
  if (cfg->env_kind == MSORT_ENV_SORT) sort_obs(cfg, s, o);
  else if (cfg->env_kind == MSORT_ENV_PRESS) press_obs(cfg, s, o);
  else { sort_obs(cfg, s, o); press_obs(cfg, s, o + 13); } /* env_monolith.py:98-104 */
}

/* ref: press_action_masks env_super.py:869-885 ; monolith_action_masks :887-898 ;
 * Env_1.action_masks env_1_sort.py:74-76 */
static unsigned press_mask_bits(const msort_config_t* cfg, const msort_env_state_t* s) {
  unsigned m = 1u;
  for (int i = 0; i < 5; ++i) {
    if (level_of(s, i) >= cfg->bale_size) {
      if (s->press_timer[0] == 0) m |= 1u << (1 + i);
      if (s->press_timer[1] == 0) m |= 1u << (6 + i);
    }
  }
  return m;
}

static void write_mask(const msort_config_t* cfg, const msort_env_state_t* s, uint8_t* mk) {
  if (cfg->env_kind == MSORT_ENV_SORT) { mk[0] = 1; mk[1] = 1; return; }
  unsigned b = press_mask_bits(cfg, s);
  for (int i = 0; i < 11; ++i) mk[i] = (b >> i) & 1u;
  if (cfg->env_kind == MSORT_ENV_MONO) for (int i = 0; i < 11; ++i) mk[11 + i] = mk[i];
}

/* ref: validate_press_action, env_super.py:811-836 */
static int press_action_valid(const msort_config_t* cfg, const msort_env_state_t* s, int pa) {
  if (pa == 0) return 1;
  int p = pa <= 5 ? 0 : 1, m = (pa - 1) % 5;
  if (s->press_timer[p] > 0) return 0;
  if (level_of(s, m) < cfg->bale_size) return 0;
  return 1;
}

/* ref: press_bale, env_super.py:661-687 */
static int press_bale(const msort_config_t* cfg, msort_env_state_t* s, int m, int n, int qk) {
  double q = (double)qk / 100.0;      /* == rint(x*100)/100 stored by use_press */
  int q100 = (int)(q * 100.0);        /* int(q*100): truncation (:663) */
  int S = cfg->bale_size, made = 0;
  int full = n / S, rem = n % S;
  if (full > 0) {
    s->bale_n[m] += full; s->bale_sum[m] += full * S;
    s->bale_last_size[m] = S; s->bale_last_q[m] = q100; made += full;
  }
  if (rem > 0) {
    if ((double)rem > (double)S * cfg->bale_remainder_threshold || s->bale_n[m] == 0) {
      s->bale_n[m] += 1; s->bale_last_size[m] = rem; s->bale_last_q[m] = q100; made += 1;
    } else {
      s->bale_last_size[m] += rem;    /* merged into the last bale, quality kept (:679-681) */
    }
    s->bale_sum[m] += rem;
  }
  return made;
}

/* fp32 MLP 13->32->32->2, tanh (ref: training.py:115 net_arch, SB3 MlpPolicy; call site
 * env_2_press.py:106-109).  weights: [W1(32x13) b1(32) W2(32x32) b2(32) W3(2x32) b3(2)]. */
static int mlp_sort_mode(const float* w, const float* x, float* margin) {
  const float *W1 = w, *b1 = w + 416, *W2 = w + 448, *b2 = w + 1472, *W3 = w + 1504, *b3 = w + 1568;
  float h1[32], h2[32], lg[2];
  for (int j = 0; j < 32; ++j) {
    float a = b1[j];
    for (int k = 0; k < 13; ++k) a += W1[j * 13 + k] * x[k];
    h1[j] = tanhf(a);
  }
  for (int j = 0; j < 32; ++j) {
    float a = b2[j];
    for (int k = 0; k < 32; ++k) a += W2[j * 32 + k] * h1[k];
    h2[j] = tanhf(a);
  }
  for (int j = 0; j < 2; ++j) {
    float a = b3[j];
    for (int k = 0; k < 32; ++k) a += W3[j * 32 + k] * h2[k];
    lg[j] = a;
  }
  if (margin) *margin = fabsf(lg[0] - lg[1]);
  return lg[1] > lg[0] ? 1 : 0; /* argmax, first index wins ties */
}

/* ---------------------------------------------------------------- reset */
/* ref: Env_Super.reset, env_super.py:365-420 */
static void reset_one(const msort_config_t* cfg, msort_env_state_t* s, int64_t gid, int first_pattern,
                      int fresh) {
  int32_t cursor = fresh ? 0 : s->replay_cursor;       /* streams run on across unseeded resets (:377) */
  int32_t episode = fresh ? 0 : s->episode + 1;
  memset(s, 0, sizeof(*s));
  s->replay_cursor = cursor;
  s->episode = episode;
  for (int m = 0; m < 4; ++m) s->acc_belt[m] = cfg->baseline_accuracy[m]; /* :395 */
  if (first_pattern == 1 || first_pattern == 2) s->gen_first = first_pattern;
  else {
    uint32_t r[4];
    env_draw(cfg, gid, BLK_RESET, (uint32_t)episode, 0u, r);
    s->gen_first = 1 + (int)(r[0] & 1u); /* rng.permutation([1,2])[0], input_generator.py:30 */
  }
}

int mso_reset(const msort_config_t* cfg, msort_env_state_t* st, int64_t n, const uint8_t* which,
              const uint8_t* first_pattern, float* obs, uint8_t* mask) {
  int D = kind_obs_dim(cfg->env_kind), A = kind_num_actions(cfg->env_kind);
  for (int64_t i = 0; i < n; ++i) {
    if (which && !which[i]) continue;
    reset_one(cfg, &st[i], cfg->global_env_offset + i, first_pattern ? first_pattern[i] : 0, 1);
    if (obs) write_obs(cfg, &st[i], obs + i * D);
    if (mask) write_mask(cfg, &st[i], mask + i * A);
  }
  return 0;
}

/* ---------------------------------------------------------------- step */
typedef struct mso_step_out {
  float* reward32;   /* nullable */
  double* reward64;  /* nullable */
  float* mlp_margin; /* nullable: |logit0-logit1| of the embedded policy */
  /* Recording of a PHILOX step as REPLAY inputs (all nullable).  What the step drew is written in the form the
   * reference's numpy generators would have had to produce for the SAME outcome, so the trajectory can be replayed
   * through the REPLAY instantiations and through the unmodified reference (oracle/ref_drive.py):
   *  rec_noise_u [n,4]   the four uniforms of update_accuracy (env_super.py:508)
   *  rec_redis_u [n,cap] one uniform per `rng.choice(4, p=leftover/total)` call of sort_material (:563), in call order:
   *                      a draw that took unit r of the pool of `tot` leftovers becomes u = (r + 1/2) / tot — the middle
   *                      of that unit's cell of numpy's cdf, whose cells are laid out station 0..3 exactly like the
   *                      pool here (processed stations = the lump first, then the unprocessed ones); the last
   *                      station's draws, which cannot change the outcome, are recorded as 1/2
   *  rec_n_draws [n]     how many were recorded (== total false units of the step)
   *  rec_press_choice [n] Env_1's internally sampled press action (:291-300)
   *  rec_input_counts [n] the generator's batch, packed A | B<<8 | C<<16 | D<<24 */
  double* rec_noise_u;
  double* rec_redis_u;
  int32_t* rec_n_draws;
  uint8_t* rec_press_choice;
  uint32_t* rec_input_counts;
  int32_t rec_cap;
} mso_step_out_t;

typedef struct step_acc { double v[MSORT_NUM_STATS]; } step_acc_t;

static void step_one(const msort_config_t* cfg, msort_env_state_t* s, int64_t i, int64_t a_in,
                     float* obs, double* reward_out, uint8_t* term_out, uint8_t* mask,
                     const msort_info_out_t* info, const msort_replay_t* rp, const float* policy,
                     const mso_step_out_t* out, step_acc_t* acc) {
  float* const mlp_margin = out ? out->mlp_margin : NULL;
  int rec_n = 0;   /* recorded choice() uniforms of this step */
  const int kind = cfg->env_kind;
  const int D = kind_obs_dim(kind), A = kind_num_actions(kind);
  const int64_t gid = cfg->global_env_offset + i;
  const int masking = (cfg->flags & MSORT_F_ACTION_MASKING) != 0;
  const int replay = cfg->rng_mode == MSORT_RNG_REPLAY;
  const uint32_t ep = (uint32_t)s->episode, stp = (uint32_t)s->step;
  uint32_t r4[4];

  /* action range: the reference's Discrete(A) contract; out-of-range is clamped and counted */
  int64_t a = a_in;
  if (a < 0) { a = 0; acc->v[8] += 1; }
  if (a >= A) { a = A - 1; acc->v[8] += 1; }

  /* 0: rng_input.integers(60,81) is drawn and discarded (env_super.py:911-922, :433,445) — elided */
  /* 1: material flow, update_environment env_super.py:440-442 */
  for (int m = 0; m < 4; ++m) { s->sorting[m] = s->belt[m]; s->belt[m] = s->input[m]; }
  /* 2: generator, input_generator.py:37-64 (counts only; the shuffle does not change them) */
  if (replay && rp && rp->input_counts) {
    uint32_t pk = rp->input_counts[i];
    for (int m = 0; m < 4; ++m) s->input[m] = (int32_t)((pk >> (8 * m)) & 0xffu);
  } else {
    if (s->gen_counter >= cfg->steps_per_pattern) { s->gen_idx ^= 1; s->gen_counter = 0; } /* :42-43,32-35 */
    int pat = s->gen_idx == 0 ? s->gen_first : 3 - s->gen_first;                            /* :45 */
    int tot = 0;
    for (int m = 0; m < 4; ++m) { s->input[m] = cfg->pattern_counts[pat - 1][m]; tot += s->input[m]; }
    int remainder = cfg->input_batch_size - tot;                                            /* :52-55 */
    if (remainder > 0) {
      for (int k = 0; k < remainder; ++k) {
        if ((k & 3) == 0) env_draw(cfg, gid, BLK_INPUT + 0x100u * (uint32_t)(k >> 2), ep, stp, r4);
        s->input[r4[k & 3] & 3u] += 1;
      }
    }
    s->gen_counter += 1;                                                                    /* :63 */
  }
  if (out && out->rec_input_counts)
    out->rec_input_counts[i] = (uint32_t)s->input[0] | ((uint32_t)s->input[1] << 8) | ((uint32_t)s->input[2] << 16) | ((uint32_t)s->input[3] << 24);
  double acc_sorter[4];
  for (int m = 0; m < 4; ++m) acc_sorter[m] = s->acc_belt[m]; /* env_super.py:457 */

  /* 3: decode the action */
  int mode = 0, pa = 0, skip_press = 0, invalid = 0;
  if (kind == MSORT_ENV_SORT) {
    mode = (int)a;                                       /* env_1_sort.py:114 */
  } else if (kind == MSORT_ENV_MONO) {
    mode = (int)(a / 11); pa = (int)(a % 11);            /* env_monolith.py:127-129 */
    if (!masking && !press_action_valid(cfg, s, pa)) {   /* :132-138 — levels BEFORE this step's sort */
      pa = 0; skip_press = 1; invalid = 1;               /* :237-243: press_action_rules is not called */
    }
  } else {
    pa = (int)a;                                         /* env_2_press.py:125 */
    if (replay && rp && rp->sort_mode) mode = rp->sort_mode[i];
    else if (cfg->flags & MSORT_F_SORT_POLICY_MLP) {     /* env_2_press.py:106-109 */
      float so[13];
      sort_obs(cfg, s, so);
      mode = mlp_sort_mode(policy, so, mlp_margin ? &mlp_margin[i] : NULL);
    } else {                                             /* sorting_rules, env_super.py:469-482 */
      double prop[4];
      belt_proportions(s, prop);
      mode = (prop[0] + prop[2] > prop[1] + prop[3]) ? 0 : 1;
    }
  }
  s->sensor_mode = mode;                                 /* set_multisensor_mode :484-490 */

  /* 4: update_accuracy env_super.py:492-509 */
  {
    double u[4];
    if (replay) for (int m = 0; m < 4; ++m) u[m] = rp->noise_u[i * 4 + m];
    else {
      env_draw(cfg, gid, BLK_NOISE, ep, stp, r4);
      for (int m = 0; m < 4; ++m) u[m] = (double)r4[m] * (1.0 / 4294967296.0);
    }
    if (out && out->rec_noise_u) for (int m = 0; m < 4; ++m) out->rec_noise_u[i * 4 + m] = u[m];
    double low = -cfg->noise, range = cfg->noise - low;  /* numpy uniform: low + (high-low)*u */
    for (int m = 0; m < 4; ++m) {
      double base = cfg->baseline_accuracy[m];
      if ((mode == 0 && (m == 0 || m == 2)) || (mode == 1 && (m == 1 || m == 3))) base = base + cfg->boost;
      double nz = low + range * u[m];
      s->acc_belt[m] = clipd(base + nz, 0.0, 1.0);
    }
  }

  /* 5: sort_material env_super.py:511-609 */
  uint32_t sorted_true4 = 0;   /* true_arr (:539) packed one byte per station; reported for the logged mean purity (:605) */
  if (replay) {
    /* REPLAY: the reference's loop nest, one recorded numpy uniform per `choice` call. */
    int L[4], T[4] = {0, 0, 0, 0}, F[4] = {0, 0, 0, 0};
    for (int m = 0; m < 4; ++m) L[m] = s->sorting[m];
    for (int m = 0; m < 4; ++m) {
      int t = L[m];                                               /* :535 (already reduced) */
      int tr = (int)rint((double)t * acc_sorter[m]);              /* :539 half-to-even */
      int f = t - tr;
      T[m] = tr; F[m] = f; L[m] = f;                              /* :542-546 */
      for (int k = 0; k < f; ++k) {                               /* :553-571 */
        int tot = L[0] + L[1] + L[2] + L[3];
        if (tot == 0) break;                                      /* :557-559 (unreachable: tot >= f-k) */
        int j;
        if (s->replay_cursor >= rp->redis_len) {
          acc->v[9] += 1;                       /* stream exhausted: defined fallback, reported */
          j = 0; while (L[j] == 0) ++j;
        } else {
          double uu = rp->redis_u[i * rp->redis_len + s->replay_cursor];
          s->replay_cursor += 1;
          double cdf[4], c = 0.0;                                 /* numpy Generator.choice(4, p=) */
          for (int q = 0; q < 4; ++q) { c = c + (double)L[q] / (double)tot; cdf[q] = c; }
          for (int q = 0; q < 4; ++q) cdf[q] = cdf[q] / c;        /* cdf /= cdf[-1] */
          j = 0;
          while (j < 4 && cdf[j] <= uu) ++j;                      /* searchsorted(u, 'right') */
          if (j > 3) j = 3;
        }
        L[j] -= 1;                                                /* :566-568 */
      }
    }
    s->cont_e += L[0] + L[1] + L[2] + L[3];                       /* :579,597 */
    for (int q = 0; q < 4; ++q) { s->cont_true[q] += T[q]; s->cont_false[q] += F[q]; } /* :600-602 */
    for (int q = 0; q < 4; ++q) sorted_true4 |= (uint32_t)T[q] << (8 * q);
  } else {
    /* PHILOX: the same random process in the form the device kernel evaluates it (DESIGN.md §4
     * "Sorting").  Each draw removes one unit chosen uniformly from the pool of leftovers
     * (p_j = leftover_j / total, :562-563).  Units removed from stations that were already
     * processed (incl. the current station's own false units) matter only through the SUM of those
     * leftovers (later selection probabilities and finally E), so they are kept as one `lump`;
     * stations still to be processed are tracked individually.  The last station's draws reduce
     * only that sum, by exactly their number, so they are not simulated.
     * Draw i of station S uses block BLK_REDIS + 64*S + i/12, whose words form two 64-bit lanes (w1:w0)
     * and (w3:w2): draws i%12 = 0..5 come from the first, 6..11 from the second, each as
     * r = hi64(lane*tot), lane = lo64(lane*tot). */
    int L[4];
    for (int m = 0; m < 4; ++m) L[m] = s->sorting[m];
    int tot = L[0] + L[1] + L[2] + L[3], lump = 0;
    for (int S = 0; S < 4; ++S) {
      int t = L[S];
      int tr = (int)rint((double)t * acc_sorter[S]);              /* :539 half-to-even */
      int f = t - tr;
      s->cont_true[S] += tr; s->cont_false[S] += f;               /* :600-602 */
      sorted_true4 |= (uint32_t)tr << (8 * S);
      tot -= tr;
      if (S == 3) {                                               /* its f draws leave lump unchanged */
        if (out && out->rec_redis_u)
          for (int k = 0; k < f; ++k, ++rec_n) if (rec_n < out->rec_cap) out->rec_redis_u[i * out->rec_cap + rec_n] = 0.5;
        break;
      }
      lump += f;                                                  /* leftover[S] = false_val joins the lump */
      uint64_t lane = 0;
      for (int k = 0; k < f; ++k) {
        if (k % 12 == 0) env_draw(cfg, gid, BLK_REDIS + 64u * (uint32_t)S + (uint32_t)(k / 12), ep, stp, r4);
        if (k % 6 == 0) lane = (k % 12 == 0) ? (((uint64_t)r4[1] << 32) | r4[0]) : (((uint64_t)r4[3] << 32) | r4[2]);
        unsigned __int128 prod = (unsigned __int128)lane * (uint64_t)(uint32_t)tot;
        int r = (int)(uint32_t)(prod >> 64);
        lane = (uint64_t)prod;
        if (out && out->rec_redis_u) {
          if (rec_n < out->rec_cap) out->rec_redis_u[i * out->rec_cap + rec_n] = ((double)r + 0.5) / (double)tot;
          ++rec_n;
        }
        int c = lump;
        if (r < c) lump -= 1;
        else {
          int q = S + 1;
          for (; q < 3; ++q) { c += L[q]; if (r < c) break; }
          L[q] -= 1;
        }
        tot -= 1;
      }
    }
    s->cont_e += lump;                                            /* :579,597: sum of all leftovers */
  }

  if (out && out->rec_n_draws) out->rec_n_draws[i] = rec_n;

  /* 6: Env_1 samples its own press action under the mask (env_super.py:291-300) */
  if (kind == MSORT_ENV_SORT) {
    if (replay) pa = rp->press_choice[i];
    else {
      unsigned vb = press_mask_bits(cfg, s);
      int nv = __builtin_popcount(vb);
      env_draw(cfg, gid, BLK_PRESS, ep, stp, r4);
      int pick = (int)(((uint64_t)r4[0] * (uint64_t)nv) >> 32);
      pa = 0;
      for (int b = 0; b < 11; ++b) if ((vb >> b) & 1u) { if (pick == 0) { pa = b; break; } --pick; }
      if (out && out->rec_press_choice) out->rec_press_choice[i] = (uint8_t)pa;
    }
  }
  /* Env_2 without masking sanitises HERE, on post-sort levels (env_2_press.py:127-131) */
  if (kind == MSORT_ENV_PRESS && !masking && !press_action_valid(cfg, s, pa)) { pa = 0; invalid = 1; }

  /* 7: press_action_rules env_super.py:626-640 */
  int bales_made = 0;
  if (!skip_press) {
    for (int p = 0; p < 2; ++p) {                                 /* check_press_status :642-659 */
      if (s->press_timer[p] > 0) {
        s->press_timer[p] -= 1;
        if (s->press_timer[p] == 0) {
          bales_made += press_bale(cfg, s, s->press_mat[p], s->press_n[p], s->press_q[p]);
          s->press_mat[p] = 0; s->press_n[p] = 0; s->press_q[p] = 0;
        }
      }
    }
    if (pa != 0) {                                                /* use_press :722-769 */
      int p = pa <= 5 ? 0 : 1, m = (pa - 1) % 5;                  /* :804-809 */
      if (s->press_timer[p] > 0) {
        /* busy press: flagged + logged, no effect (:725-733) */
      } else {
        int amt = level_of(s, m);
        s->last_press_started = 1; s->last_press_amount = amt;    /* :745-746 */
        int qk = 0;
        if (m < 4 && amt > 0) qk = (int)rint((double)s->cont_true[m] / (double)amt * 100.0); /* :754 */
        if (m < 4) { s->cont_true[m] = 0; s->cont_false[m] = 0; } else s->cont_e = 0;        /* :760-762 */
        s->press_timer[p] = cfg->press_time[p]; s->press_mat[p] = m;
        s->press_n[p] = amt; s->press_q[p] = qk;                  /* :765-769 */
      }
    }
  }

  /* 8: overflow termination (only when the caller asks; env_monolith.py:265-272 etc.) */
  int overflow = 0, overflow_mat = -1;
  if (cfg->flags & MSORT_F_CHECK_OVERFLOW) {
    for (int m = 0; m < 5; ++m)                                   /* detect_overflow :900-905 */
      if (level_of(s, m) > cfg->container_capacity) { overflow = 1; overflow_mat = m; break; }
  }

  double reward, rs_term = 0.0, rp_term = 0.0;   /* the two terms as _log_step_data records them (env_super.py:933) */
  int terminated;
  if (overflow) {
    reward = cfg->overflow_termination_penalty;
    /* env_monolith.py:271 logs (reward/2, reward/2); env_1_sort.py:141, env_2_press.py:152 log (0, reward) */
    if (kind == MSORT_ENV_MONO) { rs_term = rp_term = reward / 2; } else { rp_term = reward; }
    s->step += 1;
    terminated = 1;
  } else {
    /* 9: rewards */
    double r_sort = 0.0, r_press = 0.0;
    if (kind != MSORT_ENV_PRESS) {                                /* calculate_sorting_reward :963-1003 */
      double pur[4], total = 0.0;
      container_purity(cfg, s, pur);
      for (int m = 0; m < 4; ++m) total += pur[m] - cfg->purity_theta;
      double state_based = (total / 4.0) * cfg->purity_scaling;
      r_sort = tanh(state_based / cfg->tanh_temperature);
    }
    if (kind != MSORT_ENV_SORT) {                                 /* calculate_press_reward :1006-1080 */
      double max_pen = 0.0; int catastrophic = 0, total_level = 0;
      for (int m = 0; m < 5; ++m) {
        int lvl = level_of(s, m);
        total_level += lvl;
        double fill = (double)lvl / (double)cfg->container_capacity;
        if (fill > 1.0) { catastrophic = 1; break; }
        else if (fill > 0.95) { if (cfg->overflow_penalty_severe < max_pen) max_pen = cfg->overflow_penalty_severe; }
        else if (fill > 0.90) { if (cfg->overflow_penalty_mild < max_pen) max_pen = cfg->overflow_penalty_mild; }
      }
      if (catastrophic) r_press = cfg->overflow_penalty_catastrophic;          /* :1022-1023 */
      else if (max_pen < 0) r_press = max_pen;                                  /* :1029-1030 — flag NOT cleared */
      else {
        total_level = 0;
        for (int m = 0; m < 5; ++m) total_level += level_of(s, m);
        double state_reward = ((double)total_level / (double)(5 * cfg->container_capacity)) * cfg->max_state_reward;
        double action_reward = 0.0;
        if (s->last_press_started) {                                            /* :1054-1075 */
          int S = cfg->bale_size, amount = s->last_press_amount;
          int nb = amount / S, rem = amount % S;
          int d = rem < S - rem ? rem : S - rem;
          double eff = (1.0 - 4.0 * ((double)d / (double)S)) * cfg->bale_efficiency_factor;
          static const double peaks[4] = {0.0, 1.0 / 3.0, 2.0 / 3.0, 1.0};
          double bonus = peaks[nb < 3 ? nb : 3] - cfg->bale_efficiency_factor;
          action_reward = eff + bonus;
          s->last_press_started = 0; s->last_press_amount = 0;
        }
        r_press = clipd(state_reward + action_reward, -1.0, 1.0);
      }
    }
    reward = kind == MSORT_ENV_SORT ? r_sort : (kind == MSORT_ENV_PRESS ? r_press : r_sort + r_press);
    rs_term = r_sort; rp_term = r_press;
    s->step += 1;                                                 /* e.g. env_monolith.py:279-280 */
    terminated = s->step >= cfg->max_steps;
  }

  /* 10: observation, bookkeeping, auto-reset */
  s->ep_return += reward;
  write_obs(cfg, s, obs + i * D);
  if (reward_out) *reward_out = reward;
  term_out[i] = (uint8_t)terminated;
  acc->v[3] += 1; acc->v[4] += reward; acc->v[5] += overflow; acc->v[6] += bales_made; acc->v[7] += invalid;
  if (info) {
    if (info->action) info->action[i] = a;
    if (info->overflow) info->overflow[i] = (uint8_t)overflow;
    if (info->overflow_material) info->overflow_material[i] = (int8_t)overflow_mat;
    if (info->sort_mode) info->sort_mode[i] = (uint8_t)mode;
    if (info->press_action) info->press_action[i] = (uint8_t)pa;
    if (info->invalid_action) info->invalid_action[i] = (uint8_t)invalid;
    if (info->sorted_true) info->sorted_true[i] = sorted_true4;
    if (info->reward_sort) info->reward_sort[i] = (float)rs_term;
    if (info->reward_press) info->reward_press[i] = (float)rp_term;
  }
  if (terminated) {
    acc->v[0] += 1; acc->v[1] += s->ep_return; acc->v[2] += s->step;
    if (info) {
      if (info->episode_return) info->episode_return[i] = s->ep_return;
      if (info->episode_length) info->episode_length[i] = s->step;
    }
    if (cfg->flags & MSORT_F_AUTO_RESET) {
      if (info && info->terminal_obs) memcpy(info->terminal_obs + i * D, obs + i * D, sizeof(float) * D);
      reset_one(cfg, s, gid, 0, 0);
      write_obs(cfg, s, obs + i * D);
    }
  }
  if (mask) write_mask(cfg, s, mask + i * A);
}

typedef struct step_job {
  const msort_config_t* cfg; msort_env_state_t* st; int64_t lo, hi; const int64_t* actions;
  float* obs; uint8_t* terminated; uint8_t* mask; double* r64; float* r32; float* mm;
  const msort_info_out_t* info; const msort_replay_t* rp; const float* policy; const mso_step_out_t* out; step_acc_t acc;
} step_job_t;

static void* step_worker(void* arg) {
  step_job_t* j = (step_job_t*)arg;
  memset(&j->acc, 0, sizeof(j->acc));
  for (int64_t i = j->lo; i < j->hi; ++i) {
    double r = 0.0;
    step_one(j->cfg, &j->st[i], i, j->actions[i], j->obs, &r, j->terminated, j->mask, j->info, j->rp,
             j->policy, j->out, &j->acc);
    if (j->r64) j->r64[i] = r;
    if (j->r32) j->r32[i] = (float)r;
  }
  return NULL;
}

#define MSO_MAX_THREADS 256

int mso_step(const msort_config_t* cfg, msort_env_state_t* st, int64_t n, const int64_t* actions,
             float* obs, uint8_t* terminated, uint8_t* mask, const mso_step_out_t* out,
             const msort_info_out_t* info, const msort_replay_t* rp, const float* policy, int nthreads) {
  if (cfg->rng_mode == MSORT_RNG_REPLAY && (!rp || !rp->noise_u || !rp->redis_u)) return MSORT_E_REPLAY;
  if (cfg->rng_mode == MSORT_RNG_REPLAY && cfg->env_kind == MSORT_ENV_SORT && !rp->press_choice) return MSORT_E_REPLAY;
  if ((cfg->flags & MSORT_F_SORT_POLICY_MLP) && cfg->env_kind == MSORT_ENV_PRESS && !policy &&
      !(rp && rp->sort_mode)) return MSORT_E_INVALID;
  if (nthreads < 1) nthreads = 1;
  if (nthreads > MSO_MAX_THREADS) nthreads = MSO_MAX_THREADS;
  if ((int64_t)nthreads > n) nthreads = n > 0 ? (int)n : 1;
  step_job_t jobs[MSO_MAX_THREADS];
  pthread_t tid[MSO_MAX_THREADS];
  for (int c = 0; c < nthreads; ++c) {
    step_job_t* j = &jobs[c];
    j->cfg = cfg; j->st = st; j->lo = n * c / nthreads; j->hi = n * (c + 1) / nthreads;
    j->actions = actions; j->obs = obs; j->terminated = terminated; j->mask = mask;
    j->r64 = out ? out->reward64 : NULL; j->r32 = out ? out->reward32 : NULL;
    j->mm = out ? out->mlp_margin : NULL; j->info = info; j->rp = rp; j->policy = policy; j->out = out;
  }
  for (int c = 1; c < nthreads; ++c) pthread_create(&tid[c], NULL, step_worker, &jobs[c]);
  step_worker(&jobs[0]);
  for (int c = 1; c < nthreads; ++c) pthread_join(tid[c], NULL);
  step_acc_t total;
  memset(&total, 0, sizeof(total));
  for (int c = 0; c < nthreads; ++c)
    for (int k = 0; k < MSORT_NUM_STATS; ++k) total.v[k] += jobs[c].acc.v[k];
  if (info && info->stats) for (int k = 0; k < MSORT_NUM_STATS; ++k) info->stats[k] += total.v[k];
  return total.v[9] > 0 ? MSORT_E_REPLAY : 0;
}

/* Uniform masked-random action source for CPU-baseline timing and statistical tests: picks a
 * valid action from the CURRENT mask of each env (what a masked random policy would do). */
void mso_sample_masked_actions(const msort_config_t* cfg, const msort_env_state_t* st, int64_t n,
                               uint64_t seed, uint32_t t, int64_t* actions) {
  msort_config_t c2 = *cfg;
  c2.seed = seed;
  for (int64_t i = 0; i < n; ++i) {
    uint32_t r[4];
    env_draw(&c2, cfg->global_env_offset + i, 7u, 0u, t, r);
    if (cfg->env_kind == MSORT_ENV_SORT) { actions[i] = r[0] & 1u; continue; }
    unsigned vb = press_mask_bits(cfg, &st[i]);
    int nv = __builtin_popcount(vb);
    int pick = (int)(((uint64_t)r[0] * (uint64_t)nv) >> 32), pa = 0;
    for (int b = 0; b < 11; ++b) if ((vb >> b) & 1u) { if (pick == 0) { pa = b; break; } --pick; }
    actions[i] = cfg->env_kind == MSORT_ENV_MONO ? (int64_t)((r[1] & 1u) * 11 + pa) : pa;
  }
}

/* Rule-based action source (sorting_rules env_super.py:469-482 + check_container_level :689-720, combined as
 * Env_3.step(mode='rule_based') env_monolith.py:166-184).  after_shift: evaluate sorting_rules on the belt
 * as it is after this step's material shift (belt <- input), where Env_3.step calls it. */
void mso_rule_based_actions(const msort_config_t* cfg, const msort_env_state_t* st, int64_t n, int after_shift,
                            int64_t* actions) {
  for (int64_t i = 0; i < n; ++i) {
    const msort_env_state_t* s = &st[i];
    const int32_t* belt = after_shift ? s->input : s->belt;
    int tot = belt[0] + belt[1] + belt[2] + belt[3];
    double p[4];
    for (int m = 0; m < 4; ++m) p[m] = tot > 0 ? (double)belt[m] / (double)tot : 0.0;
    int mode = (p[0] + p[2] > p[1] + p[3]) ? 0 : 1;
    int press = 0;
    int free_press = s->press_timer[0] == 0 ? 1 : (s->press_timer[1] == 0 ? 2 : 0);
    if (free_press) {
      int best = 0, idx = -1;
      for (int m = 0; m < 4; ++m) { int l = s->cont_true[m] + s->cont_false[m]; if (l > best) { best = l; idx = m; } }
      if (s->cont_e > best) { best = s->cont_e; idx = 4; }
      if (best > 0) press = (free_press - 1) * 5 + idx + 1;
    }
    actions[i] = cfg->env_kind == MSORT_ENV_SORT ? mode : (cfg->env_kind == MSORT_ENV_PRESS ? press : 11 * mode + press);
  }
}

/* CPU-baseline driver: each of `nthreads` workers owns a contiguous slice of the envs and runs
 * T masked-random steps over it with auto-reset, like one SubprocVecEnv worker stepping its
 * envs (no per-step barrier between workers).  Returns env-steps executed.
 * (bench.py cpu_baseline / --impl reference.) */
typedef struct roll_job {
  const msort_config_t* cfg; msort_env_state_t* st; int64_t lo, hi; int T; uint64_t action_seed;
  const float* policy; double stats[MSORT_NUM_STATS]; int64_t done;
} roll_job_t;

static void* roll_worker(void* arg) {
  roll_job_t* j = (roll_job_t*)arg;
  const msort_config_t* cfg = j->cfg;
  int64_t n = j->hi - j->lo;
  int D = kind_obs_dim(cfg->env_kind), A = kind_num_actions(cfg->env_kind);
  msort_config_t cc = *cfg;
  cc.global_env_offset = cfg->global_env_offset + j->lo;
  cc.num_envs = n;
  float* obs = (float*)malloc(sizeof(float) * (size_t)(n > 0 ? n : 1) * D);
  uint8_t* term = (uint8_t*)malloc((size_t)(n > 0 ? n : 1));
  uint8_t* mask = (uint8_t*)malloc((size_t)(n > 0 ? n : 1) * A);
  int64_t* act = (int64_t*)malloc(sizeof(int64_t) * (size_t)(n > 0 ? n : 1));
  float* rew = (float*)malloc(sizeof(float) * (size_t)(n > 0 ? n : 1));
  memset(j->stats, 0, sizeof(j->stats));
  msort_info_out_t info;
  memset(&info, 0, sizeof(info));
  info.struct_size = sizeof(info);
  info.stats = j->stats;
  mso_step_out_t out;
  memset(&out, 0, sizeof(out));
  out.reward32 = rew;
  j->done = 0;
  for (int t = 0; t < j->T; ++t) {
    mso_sample_masked_actions(&cc, j->st + j->lo, n, j->action_seed, (uint32_t)t, act);
    mso_step(&cc, j->st + j->lo, n, act, obs, term, mask, &out, &info, NULL, j->policy, 1);
    j->done += n;
  }
  free(obs); free(term); free(mask); free(act); free(rew);
  return NULL;
}

int64_t mso_rollout(const msort_config_t* cfg, msort_env_state_t* st, int64_t n, int T, uint64_t action_seed,
                    const float* policy, int nthreads, double* stats16) {
  if (nthreads < 1) nthreads = 1;
  if (nthreads > MSO_MAX_THREADS) nthreads = MSO_MAX_THREADS;
  if ((int64_t)nthreads > n) nthreads = n > 0 ? (int)n : 1;
  roll_job_t jobs[MSO_MAX_THREADS];
  pthread_t tid[MSO_MAX_THREADS];
  for (int c = 0; c < nthreads; ++c) {
    jobs[c].cfg = cfg; jobs[c].st = st; jobs[c].lo = n * c / nthreads; jobs[c].hi = n * (c + 1) / nthreads;
    jobs[c].T = T; jobs[c].action_seed = action_seed; jobs[c].policy = policy;
  }
  for (int c = 1; c < nthreads; ++c) pthread_create(&tid[c], NULL, roll_worker, &jobs[c]);
  roll_worker(&jobs[0]);
  for (int c = 1; c < nthreads; ++c) pthread_join(tid[c], NULL);
  int64_t done = 0;
  if (stats16) memset(stats16, 0, sizeof(double) * MSORT_NUM_STATS);
  for (int c = 0; c < nthreads; ++c) {
    done += jobs[c].done;
    if (stats16) for (int k = 0; k < MSORT_NUM_STATS; ++k) stats16[k] += jobs[c].stats[k];
  }
  return done;
}

int mso_state_size(void) { return (int)sizeof(msort_env_state_t); }
int mso_config_size(void) { return (int)sizeof(msort_config_t); }
