/*
 * msort.h — C ABI of libmsort.so, the B200 (sm_100a) batched simulator for the
 * step()/reset() dynamics of MARL-SortingEnv's Env_1_Sorting / Env_2_Pressing /
 * Env_3_Monolith.
 *
 * Boundary.  The reference has no FFI: its boundary is the Gymnasium
 * surface of three Python classes.  Each entry point below replaces the reference
 * interface cited next to it ("ref:" paths are relative to the reference checkout).
 * A maintainer binds these with ctypes; the in-repo Python host
 * (marl-sortingenv_b200/) is exactly such a binding.
 *
 * Ownership.  Every device buffer (state, actions, obs, reward, flags, masks, replay
 * streams, stats) is allocated by the CALLER (torch tensors in the Python host) and
 * passed as a raw device pointer.  The library allocates only its small host handle.
 * It never frees caller memory and never synchronises the device, except in
 * msort_sync_check().
 *
 * Errors.  Every call returns 0 (MSORT_OK) or a negative msort_status; the text is in
 * msort_last_error() (thread-local).  No C++ exception crosses the ABI.  There is NO
 * CPU fallback: msort_create() fails unless the device reports compute capability 10.x.
 *
 * Threading.  A handle is not thread-safe; distinct handles are independent.
 */
#ifndef MSORT_H_
#define MSORT_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ABI history: 2 telemetry (msort_gather_state, reward terms) and msort_policy_act; 3 msort_observe_after_shift,
 * msort_step_variant, obs / mask must be 16-byte aligned; 4 msort_step_range, msort_policy_act_range;
 * 5 msort_set_option, MSORT_STEP_HOT_TENSOR, msort_step_host, msort_generate_streams, msort_ppo_* (scratch sized by
 * msort_ppo_scratch_floats), msort_rollout_pack / _step / _policy, msort_policy_eval, MSORT_OPT_TENSOR_POLICY 2 (split form). */
#define MSORT_ABI_VERSION 5

/* ------------------------------------------------------------------ enums */
typedef enum msort_status {
  MSORT_OK = 0,
  MSORT_E_INVALID = -1,     /* bad argument (null, misaligned, out of range, bad struct_size) */
  MSORT_E_UNSUPPORTED = -2, /* configuration the device path does not implement             */
  MSORT_E_NO_DEVICE = -3,   /* no CUDA device / not compute capability 10.x                 */
  MSORT_E_CUDA = -4,        /* CUDA runtime error (launch, sticky fault)                    */
  MSORT_E_REPLAY = -5       /* replay stream exhausted / missing in REPLAY mode             */
} msort_status;

/* ref: `name` attribute "sort" / "press" / "mono" (env_1_sort.py:26, env_2_press.py:26,
 * env_monolith.py:28) */
typedef enum msort_env_kind { MSORT_ENV_SORT = 1, MSORT_ENV_PRESS = 2, MSORT_ENV_MONO = 3 } msort_env_kind;

/* Random source of the plant.
 *  PHILOX : counter-based Philox4x32-10 keyed by cfg.seed, counter =
 *           (global env id, draw block, episode, step) — production mode; same
 *           distributions as the reference's numpy streams, different bits.
 *  REPLAY : consumes streams recorded from the reference's own numpy generators
 *           (ref: env_super.py:170-174) so both implementations can be compared
 *           bit for bit. */
typedef enum msort_rng_mode { MSORT_RNG_PHILOX = 0, MSORT_RNG_REPLAY = 1 } msort_rng_mode;

enum {
  /* ref: step(..., use_action_masking=True) — env_2_press.py:88, env_monolith.py:109 */
  MSORT_F_ACTION_MASKING = 1u << 0,
  /* ref: step(..., check_overflow=False) — env_1_sort.py:133, env_2_press.py:145, env_monolith.py:265 */
  MSORT_F_CHECK_OVERFLOW = 1u << 1,
  /* SB3 VecEnv semantics: on `terminated` keep the terminal obs aside and reset() unseeded */
  MSORT_F_AUTO_RESET = 1u << 2,
  /* Env_2 only: sort mode from the embedded MLP (ref: env_2_press.py:106-109) instead of
   * sorting_rules() (ref: env_2_press.py:112, env_super.py:469-482) */
  MSORT_F_SORT_POLICY_MLP = 1u << 3
};

#define MSORT_RESET_KEEP_STREAMS 1u

#define MSORT_OBS_DIM_SORT 13
#define MSORT_OBS_DIM_PRESS 16
#define MSORT_OBS_DIM_MONO 29
#define MSORT_NUM_ACTIONS_SORT 2
#define MSORT_NUM_ACTIONS_PRESS 11
#define MSORT_NUM_ACTIONS_MONO 22
#define MSORT_POLICY_WEIGHTS 1570 /* 13*32+32 + 32*32+32 + 32*2+2 ; ref: training.py:115 */
#define MSORT_NUM_STATS 16

/* ------------------------------------------------------------------ config */
/* All live config.yml scalars + ctor arguments.
 * ref: Env_Super.__init__ (env_super.py:25-137), config.yml:1-59. */
typedef struct msort_config {
  uint32_t struct_size;        /* = sizeof(msort_config_t) */
  int32_t env_kind;            /* msort_env_kind */
  int64_t num_envs;            /* N on this device */
  int64_t global_env_offset;   /* global id of local env 0 (multi-GPU sharding; Philox counter) */
  int32_t max_steps;           /* ctor max_steps (env_1_sort.py:19) */
  uint32_t flags;              /* MSORT_F_* */
  int32_t rng_mode;            /* msort_rng_mode */
  int32_t reserved0;
  uint64_t seed;               /* Philox key */
  /* simulation: */
  int32_t input_batch_size;    /* simulation.input_batch_size (env_super.py:33,445) ; <= 255 */
  int32_t steps_per_pattern;   /* 20: the generator is rebuilt by every reset() with its ctor
                                  default (env_super.py:375, input_generator.py:15) */
  int32_t pattern_counts[2][4];/* floor(ratio*batch) for pattern 1 / 2 (input_generator.py:17-20,49) */
  /* sorting_station: */
  double baseline_accuracy[4]; /* env_super.py:58 */
  double boost;                /* env_super.py:59 */
  double noise;                /* ctor noise_sorting or sorting_station.noise (env_super.py:71) */
  int32_t stage_capacity;      /* env_super.py:62 (obs normaliser) */
  /* pressing_station: */
  int32_t press_time[2];       /* env_super.py:97 */
  int32_t container_capacity;  /* env_super.py:99 */
  int32_t bale_size;           /* ctor balesize or bale_standard_size (env_super.py:87) */
  int32_t reserved1;
  double bale_remainder_threshold; /* env_super.py:88 */
  double quality_threshold[4];     /* env_super.py:98 */
  /* rewards: */
  double purity_theta;         /* env_super.py:119 */
  double purity_scaling;       /* 2.0, hard-coded at env_super.py:971 */
  double tanh_temperature;     /* env_super.py:124 */
  double overflow_penalty_catastrophic; /* env_super.py:1023 */
  double overflow_penalty_severe;       /* env_super.py:1025 */
  double overflow_penalty_mild;         /* env_super.py:1027 */
  double bale_efficiency_factor;        /* env_super.py:1059 */
  double max_state_reward;              /* env_super.py:132 */
  double overflow_termination_penalty;  /* env_super.py:133 */
} msort_config_t;

/* ------------------------------------------------------------------ plain state */
/* One env's full plant state as plain integers/doubles — the exchange format of
 * msort_export_state()/msort_import_state() and the native state of the CPU oracle.
 * ref: the attributes set in Env_Super.reset (env_super.py:365-420). */
typedef struct msort_env_state {
  int32_t input[4], belt[4], sorting[4]; /* current_material_{input,belt,sorting}            */
  int32_t cont_true[4], cont_false[4];   /* container_materials[m], [m+"_False"]              */
  int32_t cont_e;                        /* container_materials["E"]                          */
  int32_t press_timer[2];                /* press_state["press_i"]                            */
  int32_t press_mat[2];                  /* press_state["material_i"] as index 0..4 (0 if idle)*/
  int32_t press_n[2];                    /* press_state["n_i"]                                */
  int32_t press_q[2];                    /* rint(press_state["q_i"]*100), 0..100              */
  int32_t last_press_started;            /* _last_press_started                               */
  int32_t last_press_amount;             /* _last_press_amount                                */
  int32_t gen_first;                     /* input_generator.pattern_sequence[0] : 1 or 2      */
  int32_t gen_idx;                       /* input_generator.current_pattern_idx               */
  int32_t gen_counter;                   /* input_generator.step_counter                      */
  int32_t step;                          /* current_step                                      */
  int32_t episode;                       /* resets seen (Philox counter word)                 */
  int32_t sensor_mode;                   /* sensor_current_setting                            */
  int32_t replay_cursor;                 /* next unread element of the redistribution stream  */
  int32_t bale_n[5];                     /* len(bale_count[m])                                */
  int32_t bale_last_size[5];             /* bale_count[m][-1][0]                              */
  int32_t bale_last_q[5];                /* bale_count[m][-1][1]                              */
  int32_t bale_sum[5];                   /* sum of bale sizes of m                            */
  int32_t reserved;
  double acc_belt[4];                    /* accuracy_belt                                     */
  double ep_return;                      /* sum of rewards of the running episode (Monitor)   */
} msort_env_state_t;

/* ------------------------------------------------------------------ per-step I/O */
/* Replay streams for ONE step (REPLAY mode; all device pointers).
 * ref: SURVEY.md §8c / Appendix B. */
typedef struct msort_replay {
  uint32_t struct_size;
  uint32_t reserved;
  const double* noise_u;        /* [N,4] raw uniforms of rng_noise (env_super.py:508): noise = -n + 2n*u.  required */
  const double* redis_u;        /* [N,redis_len] uniforms of `rng` (env_super.py:563), read at the env's cursor. required */
  int64_t redis_len;
  const uint32_t* input_counts; /* [N] packed A|B<<8|C<<16|D<<24 emitted by the generator this step; NULL = generator rule */
  const uint8_t* press_choice;  /* [N] Env_1: internally sampled press action 0..10 (env_super.py:291-300). required for Env_1 */
  const uint8_t* sort_mode;     /* [N] Env_2: sort mode actually applied; NULL = MLP / sorting_rules */
} msort_replay_t;

/* Optional per-step outputs (every pointer nullable; device memory). */
typedef struct msort_info_out {
  uint32_t struct_size;
  uint32_t reserved;
  int64_t* action;              /* [N] info["action"] echo (clamped into range)                 */
  uint8_t* overflow;            /* [N] info["overflow"]                                          */
  int8_t* overflow_material;    /* [N] index 0..4 of info["overflow_material"], -1 if none      */
  uint8_t* sort_mode;           /* [N] sort mode applied this step                               */
  uint8_t* press_action;        /* [N] press action applied (after sanitising / Env_1 sampling) */
  uint8_t* invalid_action;      /* [N] 1 if the press action was rejected (log codes 111/222)   */
  float* terminal_obs;          /* [N,D] rows written only where terminated (AUTO_RESET)        */
  double* episode_return;       /* [N] written where terminated                                  */
  int32_t* episode_length;      /* [N] written where terminated                                  */
  double* stats;                /* [MSORT_NUM_STATS] accumulators, atomically added to:
                                   0 episodes finished, 1 sum episode return, 2 sum episode length,
                                   3 env-steps, 4 sum reward, 5 overflows, 6 bales pressed,
                                   7 invalid actions, 8 clamped actions, 9 replay under-runs */
  float* reward_sort;           /* [N] sorting term of the reward (reward_data['Reward'][t][0],
                                   env_super.py:933; 0 where the env kind does not compute it)   */
  float* reward_press;          /* [N] pressing term (reward_data['Reward'][t][1])               */
  uint32_t* sorted_true;        /* [N] units sorted correctly this step, one byte per station A..D
                                   (true_arr of sort_material, env_super.py:539; feeds the logged
                                   mean purity reward_data['Accuracy'], env_super.py:605-607)    */
} msort_info_out_t;

typedef struct msort_handle msort_t;

/* ------------------------------------------------------------------ entry points */
int msort_abi_version(void);
const char* msort_last_error(void);

/* Fill `cfg` with the reference's defaults (config.yml + ctor defaults of the env kind).
 * ref: config.yml:1-59; Env_X.__init__ defaults (env_1_sort.py:19-20). */
int msort_default_config(int env_kind, msort_config_t* cfg);

/* ref: Env_X.__init__ (env_1_sort.py:19-29, env_2_press.py:19-34, env_monolith.py:22-35). */
int msort_create(const msort_config_t* cfg, int device, msort_t** out);
int msort_destroy(msort_t* h);

/* Size in bytes of the opaque device state blob the caller must allocate (16-B aligned). */
size_t msort_state_bytes(const msort_t* h);
int msort_obs_dim(const msort_t* h);
int msort_num_actions(const msort_t* h);

/* ref: Env_X.reset(seed) (env_super.py:365-420 ; env_1_sort.py:81-85).
 *  which         : nullable u8[N]; reset only envs with which[i]!=0 (NULL = all)
 *  first_pattern : nullable u8[N] with 1|2 = pattern_sequence[0] of the freshly seeded
 *                  generator (input_generator.py:30); NULL = drawn from Philox
 *  obs, mask     : nullable outputs [N,D] f32 / [N,A] u8 (rows of reset envs only)
 *  reset_flags   : 0 = reset(seed=s): episode counter and replay cursor restart;
 *                  MSORT_RESET_KEEP_STREAMS = reset(seed=None): the RNG streams run on
 *                  (env_super.py:377-378) — episode counter +1, replay cursor kept */
int msort_reset(msort_t* h, void* state, const uint8_t* which, const uint8_t* first_pattern,
                float* obs, uint8_t* mask, uint32_t reset_flags, void* stream);

/* Change the per-call switches of step() (MSORT_F_ACTION_MASKING / MSORT_F_CHECK_OVERFLOW /
 * MSORT_F_AUTO_RESET / MSORT_F_SORT_POLICY_MLP) — the reference passes the first two as step()
 * keyword arguments (env_monolith.py:109). */
int msort_set_flags(msort_t* h, uint32_t flags);

/* Re-key the Philox generator of the WHOLE handle (reset(seed=s) with a new seed).  Call it only together with a
 * full msort_reset (which == NULL, reset_flags == 0 — that is what restarts episode numbering): the PHILOX layouts
 * do not store accuracy_belt but recompute it from (key, env, episode, step - 1, mode), so re-keying while some
 * envs are mid-episode changes their current accuracies and every later draw. */
int msort_set_seed(msort_t* h, uint64_t seed);

/* ref: Env_X.step(action, use_action_masking, check_overflow) (env_1_sort.py:97-154,
 * env_2_press.py:88-165, env_monolith.py:109-284) followed by Env_X.action_masks()
 * (env_super.py:869-898) on the new state.  One fused kernel launch.
 *  actions    : i64[N]    obs : f32[N,D]   reward : f32[N]   terminated : u8[N]
 *  mask       : u8[N,A] (nullable)   info / replay : nullable
 * state, obs and mask must be 16-byte aligned (whole tiles of 128 envs leave as 16-byte vectors / TMA bulk
 * copies), actions 8-byte (16-byte lets Env_2's persistent kernel fetch them by TMA), reward 4-byte. */
int msort_step(msort_t* h, void* state, const int64_t* actions, float* obs, float* reward,
               uint8_t* terminated, uint8_t* mask, const msort_info_out_t* info,
               const msort_replay_t* replay, void* stream);

/* ref: sort_agent.predict(sort_obs, deterministic=True) (env_2_press.py:106-109) —
 * uploads the 1570 fp32 weights [W1(32x13) b1 W2(32x32) b2 W3(2x32) b3] (device or host ptr). */
int msort_set_policy(msort_t* h, const float* weights, int weights_on_device, void* stream);

/* ref: Env_X.get_obs() / action_masks() on the current state, no transition. */
int msort_observe(msort_t* h, const void* state, float* obs, uint8_t* mask, void* stream);

/* The observation the agents of Env_3_Monolith.step(mode='model') are shown (env_monolith.py:114-115,
 * 186-221: get_sort_obs / get_press_obs are read after update_environment has moved input -> belt ->
 * sorting and before anything else of the step): same as msort_observe on that shifted plant.  The mask
 * is the current one (container levels and press timers do not change in the shift).  No transition. */
int msort_observe_after_shift(msort_t* h, const void* state, float* obs, uint8_t* mask, void* stream);

/* Masked-random action source: actions[i] = uniform choice among the valid entries of
 * mask[i, :] (all-zero row -> 0), Philox-keyed by (seed, t, global env id).
 * ref: Env_3 step(mode='random', use_action_masking=True): np.random.choice(flatnonzero(mask))
 * (env_monolith.py:152-158); sample_masked_press_action (env_super.py:291-300). */
int msort_sample_actions(msort_t* h, const uint8_t* mask, int64_t* actions, uint64_t seed, uint32_t t,
                         void* stream);

/* Generator stream export (K3).  The random inputs the PHILOX generator feeds the plant that do NOT depend on the plant's
 * state, for steps first_step .. first_step + num_steps - 1 of episode `episode` of every env, in the REPLAY
 * descriptor's form (all device pointers, all nullable):
 *   input_counts [T,N] u32  the batch the seasonal generator emits at that step, packed A | B<<8 | C<<16 | D<<24
 *                           (ref: SeasonalInputGenerator.generate_input, input_generator.py:37-64)
 *   noise_u [T,N,4] f64     the four uniforms of update_accuracy (ref: rng_noise.uniform, env_super.py:508)
 *   draw_words [T,N,12] u32 the first Philox block of redistribution words of stations 0, 1, 2 (ref: the stream behind
 *                           rng.choice in sort_material, env_super.py:563; how a word becomes a draw depends on the plant)
 *   first_pattern [N] u8    pattern_sequence[0] (1 | 2) of that episode (ref: input_generator.py:30)
 * With these (and the redistribution uniforms, which depend on the state and are recorded by the oracle) a PHILOX
 * trajectory can be replayed through the REPLAY instantiation and through the reference itself. */
int msort_generate_streams(msort_t* h, uint32_t episode, uint32_t first_step, uint32_t num_steps, uint32_t* input_counts,
                           double* noise_u, uint32_t* draw_words, uint8_t* first_pattern, void* stream);

/* Rule-based action source: sort mode from sorting_rules() (env_super.py:469-482), press job from
 * check_container_level() (env_super.py:689-720: first free press, fullest container with level > 0),
 * combined as Env_3.step(mode='rule_based') does (env_monolith.py:166-184).  Env_1 gets the sort mode,
 * Env_2 the press action.  after_shift != 0 evaluates sorting_rules() on the belt as it will be
 * AFTER this step's material shift (belt <- input), which is where Env_3.step calls it
 * (env_monolith.py:114-115 then :168); 0 evaluates it on the current belt (an external caller). */
int msort_rule_based_actions(msort_t* h, const void* state, int after_shift, int64_t* actions, void* stream);

/* Fused actor-critic inference + masked categorical draw for the GPU-resident MaskablePPO rollout
 * (ref: MaskablePPO with net_arch=dict(pi=[32,32], vf=[32,32]), training.py:115-131; SB3 tanh MLPs;
 * logits of invalid actions masked with -1e8 as sb3_contrib does).  One launch evaluates both towers for
 * all N envs on the tensor cores (tcgen05, fp16 operands, fp32 accumulation) and samples
 * action ~ softmax(masked logits) with one Philox uniform keyed by (seed, t, global env id), or takes the
 * argmax when `deterministic`.  All pointers are device memory.
 *  obs [N,D] f32, mask [N,A] u8 (the step() outputs);  actions [N] i64, logp [N] f32, value [N] f32
 *  packed_weights : MSORT_POLICY_ACT_WEIGHTS 32-bit words, 16-byte aligned, in the order the kernel stages them
 *    (see marl-sortingenv_b200/ppo.py pack_actor_critic): for each layer the weight matrix W[n][k]
 *    (n = output unit, k = input unit; towers concatenated / block-diagonal: layer 1 32->64,
 *    layer 2 64->64, layer 3 64->32 with rows 0..A-1 = logits, row A = value) stored as
 *    fp16 in the order [k/8][n][k%8] (two values per 32-bit word), then the three fp32 bias vectors
 *    (64, 64, 32). */
#define MSORT_POLICY_ACT_WEIGHTS 4256
int msort_policy_act(msort_t* h, const float* obs, const uint8_t* mask, const float* packed_weights,
                     uint64_t seed, uint32_t t, int deterministic, int64_t* actions, float* logp,
                     float* value, void* stream);

/* SoA device blob <-> plain msort_env_state_t[N] (device memory). */
int msort_export_state(msort_t* h, const void* state, msort_env_state_t* out, void* stream);
/* The same for a subset: out[j] = plain state of env env_ids[j], j < count (device memory).  The
 * telemetry recorder's per-step snapshot of the traced envs — the device-side stand-in for the
 * reference's per-step Python logs (reward_data / press_actions_per_timestep / bale_count,
 * env_super.py:928-946,631-637,661-687). */
int msort_gather_state(msort_t* h, const void* state, const int64_t* env_ids, int64_t count,
                       msort_env_state_t* out, void* stream);
int msort_import_state(msort_t* h, void* state, const msort_env_state_t* in, void* stream);

/* The same actor-critic kernels on caller-given rows: any of the three (obs_dim, num_actions) shapes on any handle, rows
 * `obs_row_stride` floats / `mask_row_stride` bytes apart (e.g. the 13-wide sort part and the 16-wide press part of Env_3's
 * 29-wide observation: the two agents of Env_3.step(mode='model'), env_monolith.py:186-221), mask NULL = every action
 * valid.  Dense, 16-byte aligned tensors take the TMA path, everything else plain loads.  The draw of row r is keyed by
 * (seed, t, global env id of row r of the handle). */
int msort_policy_eval(msort_t* h, int obs_dim, int num_actions, int64_t num_rows, const float* obs, int64_t obs_row_stride,
                      const uint8_t* mask, int64_t mask_row_stride, const float* packed_weights, uint64_t seed, uint32_t t,
                      int deterministic, int64_t* actions, float* logp, float* value, void* stream);

/* State-wide sums for the episode-statistics all-reduce (out16: device, f64[16], overwritten):
 * 0 N, 1 sum container level, 2..6 bales A..E, 7..11 bale size sums A..E, 12 sum mean purity,
 * 13 busy presses, 14 sum running episode return, 15 sum current step. */
int msort_reduce_stats(msort_t* h, const void* state, double* out16, void* stream);

/* The only call that synchronises: waits for `stream` and reports sticky faults. */
int msort_sync_check(msort_t* h, void* stream);

/* Number of kernels this handle has launched (bench.py's gpu_launches). */
int64_t msort_launch_count(const msort_t* h);

/* msort_step / msort_policy_act on the env range [first_env, first_env + num_envs) only (first_env a multiple of
 * 128; PHILOX mode).  Every pointer is the WHOLE-batch base, exactly as passed to msort_step / msort_policy_act;
 * the library takes the range's slice.  Trajectories do not depend on how the batch is cut (counters use global
 * env ids), so a caller can put disjoint ranges on different CUDA streams: the rollout loop runs
 * policy_act -> step of one half beside step / policy_act of the other (marl-sortingenv_b200/ppo.py), which lets
 * the latency-bound tensor-core policy kernel and the ALU-bound step kernel share the SMs.  The handle's
 * statistics accumulators (atomics) and launch counter are shared; the handle itself is still not thread-safe. */
int msort_step_range(msort_t* h, int64_t first_env, int64_t num_envs, void* state, const int64_t* actions,
                     float* obs, float* reward, uint8_t* terminated, uint8_t* mask,
                     const msort_info_out_t* info, void* stream);
int msort_policy_act_range(msort_t* h, int64_t first_env, int64_t num_envs, const float* obs, const uint8_t* mask,
                           const float* packed_weights, uint64_t seed, uint32_t t, int deterministic,
                           int64_t* actions, float* logp, float* value, void* stream);

/* ------------------------------------------------------------------ host-buffer step
 * step() for a caller whose actions and results live in HOST memory (an SB3-style loop; ref: the VecEnv contract
 * of training.py:64-69: actions in, obs / reward / done / action mask out, every step).  The batch is cut into
 * `chunks` env ranges; per range the library queues, on its own streams,
 *     H2D of the range's actions -> step kernel of the range -> D2H of the range's obs, reward and flag words,
 * so the host->device copy of range k+1, the kernel of range k and the device->host copy of range k-1 overlap
 * (PCIe is full duplex).  What travels: 1 byte per env in (Discrete(22) fits a byte; or 8 with actions_i64), and
 * 4*D + 4 + 2 bytes per env out — the action mask as its 11 information bits and `terminated` share one 16-bit
 * word (bit k = press action k valid, bit 15 = terminated; Env_3's 22-wide mask is that 11-bit mask twice,
 * env_super.py:887-898; Env_1's two actions are always valid).
 * All work is ordered after everything already queued on `stream`, and `stream` waits for it: the call returns at
 * once, the host buffers are complete when `stream` has drained (msort_sync_check).
 *  io->actions_u8 / actions_i64 : exactly one non-NULL, [N], host (pinned memory for real overlap)
 *  io->obs [N,D] f32, io->reward [N] f32, io->flags [N] u16 : host outputs
 *  io->dev_obs / dev_reward / dev_terminated / dev_mask : the DEVICE outputs of msort_step (same shapes and alignment
 *            rules); the step kernel writes them as usual and the host copies are taken from them, so the device-side
 *            view of the batch (action_masks(), obs) stays current
 *  scratch : DEVICE memory of msort_host_scratch_bytes(h) bytes, 256-byte aligned (action staging + flag words; caller-owned)
 *  info    : as for msort_step (per-step info arrays stay on the device) */
typedef struct msort_host_io {
  uint32_t struct_size;
  uint32_t chunks;              /* env ranges per step; 0 = library default (2) */
  const uint8_t* actions_u8;
  const int64_t* actions_i64;
  float* obs;
  float* reward;
  uint16_t* flags;
  float* dev_obs;
  float* dev_reward;
  uint8_t* dev_terminated;
  uint8_t* dev_mask;
} msort_host_io_t;
size_t msort_host_scratch_bytes(const msort_t* h);
int msort_step_host(msort_t* h, void* state, void* scratch, const msort_host_io_t* io,
                    const msort_info_out_t* info, void* stream);

/* ------------------------------------------------------------------ MaskablePPO update (the caller of the hot path)
 * Hand-written kernels for the UPDATE half of the GPU-resident training loop (ref: the reference trains with
 * sb3_contrib.MaskablePPO, training.py:115-143: net_arch=dict(pi=[32,32], vf=[32,32]), tanh, ent_coef=0.05, SB3
 * defaults otherwise).  Stateless: every buffer is device memory owned by the caller.
 * Parameter vector (flat fp32, msort_ppo_param_count(D, A) values; torch Linear layout [out][in]):
 *   pi: W1[32][D] b1[32] W2[32][32] b2[32] W3[A][32] b3[A]  |  vf: W1[32][D] b1[32] W2[32][32] b2[32] W3[1][32] b3[1] */
typedef struct msort_ppo_batch {
  uint32_t struct_size;
  int32_t obs_dim, num_actions;  /* (13,2) | (16,11) | (29,22) */
  int32_t reserved;
  int64_t num_rows;              /* rows of the rollout buffer (n_steps * num_envs) */
  const float* obs;              /* [rows, D] */
  const uint8_t* mask;           /* [rows, A] action masks (logits of invalid actions are masked with -1e8) */
  const int64_t* actions;        /* [rows] */
  const float* old_logp;         /* [rows] log-prob of the action under the rollout policy   (update only) */
  const float* adv;              /* [rows] advantages                                        (update only) */
  const float* ret;              /* [rows] returns = advantage + value                       (update only) */
} msort_ppo_batch_t;

typedef struct msort_ppo_hparams {
  uint32_t struct_size;
  int32_t normalize_advantage;   /* per minibatch: (a - mean) / (std + 1e-8), unbiased std */
  float clip_range, vf_coef, ent_coef;
  float learning_rate, beta1, beta2, adam_eps, max_grad_norm;   /* max_grad_norm <= 0: no clipping */
} msort_ppo_hparams_t;

int msort_ppo_param_count(int obs_dim, int num_actions);
/* floats of the `scratch` buffer msort_ppo_gradient / msort_ppo_update need (16-byte aligned device memory): the minibatch's
 * advantage statistics and both towers' weights transposed into the gradient kernel's shared-memory order */
int msort_ppo_scratch_floats(int obs_dim, int num_actions);
/* log-prob of every row's action and the value estimate under `params` (forward only; either output nullable) */
int msort_ppo_forward(const msort_ppo_batch_t* batch, const float* params, float* logp_out, float* value_out, void* stream);
/* GAE(lambda): rew / val / done [T, n] (done = terminated, u8), last_val [n] -> adv, ret [T, n] */
int msort_ppo_gae(int32_t T, int64_t n, const float* rew, const float* val, const uint8_t* done, const float* last_val,
                  float gamma, float gae_lambda, float* adv, float* ret, void* stream);
/* Gradient of the PPO loss (clipped surrogate + vf_coef * value MSE - ent_coef * entropy, means over the minibatch) for
 * rows idx[first .. first + count) (idx NULL: the rows themselves), ADDED into grads; stats (nullable, 5 floats) +=
 * {sum surrogate loss, sum squared value error, sum entropy, clipped rows, rows}; scratch: msort_ppo_scratch_floats() floats. */
int msort_ppo_gradient(const msort_ppo_batch_t* batch, const msort_ppo_hparams_t* hp, const float* params, float* grads,
                       const int64_t* idx, int64_t first, int64_t count, float* scratch, float* stats, void* stream);
/* The whole update: n_epochs passes over the buffer in minibatches of batch_size rows taken from perms[e] (n_epochs
 * permutations of 0..rows-1, device int64), each minibatch = gradient -> global-norm clip -> Adam step (torch.optim.Adam
 * semantics; `step` is the device-side step counter, grads must be zero on entry and are zero on exit). */
int msort_ppo_update(const msort_ppo_batch_t* batch, const msort_ppo_hparams_t* hp, float* params, float* grads, float* adam_m,
                     float* adam_v, int32_t* step, const int64_t* perms, int32_t n_epochs, int64_t batch_size, float* scratch,
                     float* stats, void* stream);

/* ------------------------------------------------------------------ fused rollout step (Env_3_Monolith)
 * One launch per env-step of the MaskablePPO rollout loop (ref: training.py:118-143 -> sb3 collect_rollouts: policy forward
 * on the last observation, masked categorical draw, env.step): msort_rollout_step = msort_step on `actions` AND, on the
 * observation / mask tile the step has just built in shared memory, the actor-critic forward + masked draw for the NEXT
 * step (tcgen05, fp16 operands, fp32 accumulation; the same 128 envs, the same threads).  Outputs of the policy half:
 * next_actions / next_logp / next_value [N] for the observation written to `obs` — what msort_policy_act(obs, mask, ...)
 * with draw index t would return, up to the kernels' rounding (both within ~2e-3 of an fp32 evaluation).  The first
 * action of a rollout comes from msort_policy_act.  Env_3, PHILOX mode, the HOT configuration only (action masking and
 * auto-reset on, no overflow check, mask wanted, no per-step info arrays, config passes the FAST checks): anything else
 * is MSORT_E_UNSUPPORTED and the caller runs msort_policy_act + msort_step.
 *  packed : MSORT_ROLLOUT_WEIGHTS words written by msort_rollout_pack from the flat fp32 parameter vector of
 *           msort_ppo_* (obs_dim 29, 22 actions), device memory, 16-byte aligned.
 *  seed, t, deterministic : as msort_policy_act (MSORT_OPT_DRAW_COUNTER is added to t the same way). */
#define MSORT_ROLLOUT_WEIGHTS 2688
int msort_rollout_pack(const float* params, uint32_t* packed, void* stream);
/* The policy half alone on given observations / masks ([N,29] f32, [N,22] u8): what msort_policy_act computes, with the
 * fused kernel's arithmetic and weights format (one CTA per 128 envs, 8 CTAs per SM; TMA tiles need 16-byte aligned tensors,
 * others take plain loads).  first_env / num_envs select a range (first_env a multiple of 128) of whole-batch tensors. */
int msort_rollout_policy(msort_t* h, int64_t first_env, int64_t num_envs, const float* obs, const uint8_t* mask,
                         const uint32_t* packed, uint64_t seed, uint32_t t, int deterministic, int64_t* actions, float* logp,
                         float* value, void* stream);
int msort_rollout_step(msort_t* h, void* state, const int64_t* actions, float* obs, float* reward, uint8_t* terminated,
                       uint8_t* mask, const msort_info_out_t* info, const uint32_t* packed, uint64_t seed, uint32_t t,
                       int deterministic, int64_t* next_actions, float* next_logp, float* next_value, void* stream);

/* Which instantiation of the step kernel the handle's last msort_step launched (diagnostics / tests):
 * 0 none yet, 1 REPLAY, 2 generic, 3 FAST (host-proved config facts compiled in),
 * 4 HOT (FAST + the training-loop switches compiled in), 5 HOT persistent (Env_2: TMA-staged tiles). */
#define MSORT_STEP_NONE 0
#define MSORT_STEP_REPLAY 1
#define MSORT_STEP_GENERIC 2
#define MSORT_STEP_FAST 3
#define MSORT_STEP_HOT 4
#define MSORT_STEP_HOT_PERSISTENT 5
#define MSORT_STEP_HOT_TENSOR 6 /* HOT persistent with Env_2's embedded policy on the tensor cores (tcgen05, fp16-split operands) */
#define MSORT_STEP_HOT_FUSED 7  /* Env_3 HOT + the next step's rollout policy in the same kernel (msort_rollout_step) */
#define MSORT_STEP_HOT_TENSOR_SPLIT 8 /* Env_2: tensor-core policy kernel (one mode byte per env) + policy-free HOT step kernel */
int msort_step_variant(const msort_t* h);

/* Handle options (diagnostics / experiments; defaults are the production choice).
 *  MSORT_OPT_TENSOR_POLICY (default MSORT_TENSOR_POLICY_DEFAULT): Env_2's embedded sort policy (ref: sort_agent.predict,
 *    env_2_press.py:106-109) is evaluated on the tensor cores when the HOT configuration runs and the weights fit the
 *    fp16 split: 1 = inside the step kernel (one launch per step), 2 = as its own kernel writing one mode byte per env,
 *    followed by a policy-free step kernel (two launches per step; bit-identical results); 0 = always the per-thread
 *    fp32 FFMA2 form. */
#ifndef MSORT_TENSOR_POLICY_DEFAULT
#define MSORT_TENSOR_POLICY_DEFAULT 1   /* measured on B200: 104.3 us per 1 048 576 envs inside the step kernel, 108.6 us as two kernels */
#endif
#define MSORT_OPT_TENSOR_POLICY 1
/*  MSORT_OPT_PERSIST_CTAS: resident CTAs per SM the persistent Env_2 kernels are launched with (default: what the
 *    kernel's registers and shared memory allow, asked at msort_create); get = the count of the kernel the next
 *    step would launch. */
#define MSORT_OPT_PERSIST_CTAS 2
/*  MSORT_OPT_DRAW_COUNTER: value = a DEVICE pointer to a uint32 (0 = none, the default).  msort_policy_act then keys its
 *    categorical draw with t + *pointer instead of t, so a CUDA graph of a whole rollout can be replayed with fresh
 *    draws: the caller bumps the counter between replays. */
#define MSORT_OPT_DRAW_COUNTER 3
int msort_set_option(msort_t* h, int option, int64_t value);
int msort_get_option(const msort_t* h, int option, int64_t* value);

/* Diagnostics: the tensor-core form of Env_2's embedded policy alone.  sort_obs [count,13] f32 (device) -> logits
 * [count,2] f32 (device), evaluated exactly as the HOT_TENSOR step kernel evaluates it (fp16-split operands, fp32
 * accumulation in TMEM), so a test can bound its error against the fp32 network of the reference
 * (sort_agent.predict, env_2_press.py:106-109).  MSORT_E_UNSUPPORTED when the policy does not fit the split. */
int msort_debug_policy_logits(msort_t* h, const float* sort_obs, int64_t count, float* logits, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MSORT_H_ */
