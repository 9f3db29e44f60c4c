// msort_launch.h — host-side declarations of the kernel launch wrappers (msort_kernels.cu),
// used by the C-ABI layer (msort_api.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/msort.h"

namespace msort {

struct DevConfig;

// Env_3's fused rollout kernel (step + the next step's policy): the packed actor-critic and where its outputs go
struct FusedLaunch {
  const uint32_t* packed;    // MSORT_ROLLOUT_WEIGHTS words (launch_pack_fused), device memory, 16-byte aligned
  int64_t* next_actions;
  float* next_logp;
  float* next_value;
  uint64_t seed;
  uint32_t t;                // draw index (+ *t_dev)
  const uint32_t* t_dev;
  int deterministic;
};

struct StepLaunch {
  void* state;
  const int64_t* actions;
  float* obs;
  float* reward;
  uint8_t* terminated;
  uint8_t* mask;
  const msort_info_out_t* info;
  const msort_replay_t* replay;
  int* variant;              // out (nullable): MSORT_STEP_* of the instantiation launched
  int allow_hot;             // 0: the state may hold stage contents no reset/step produces (imported): no HOT kernel
  const float* policy_host;  // Env_2 embedded policy (host copy in the paired layout, MSORT_POLICY_WEIGHTS floats) or nullptr
  const uint32_t* policy_tc; // the same policy packed for the tensor-core path (pack_policy_tc; DEVICE memory) or nullptr
  const int* persist_per_sm; // resident CTAs per SM of the persistent Env_2 kernels (query_persist_occupancy)
  const FusedLaunch* fused = nullptr;   // non-null: the fused rollout kernel (Env_3, HOT configuration only)
  uint8_t* mode_scratch = nullptr;      // Env_2 with the tensor-core policy: non-null = the split form (policy kernel -> mode bytes -> step kernel)
};

void pack_policy_pairs(const float* sb3, float* paired);   // SB3 weight order -> the step kernel's FFMA2 operand order
bool pack_policy_tc(const float* sb3, uint32_t* words);    // ... -> the tensor-core path's packed tiles (false: out of fp16 range)
int policy_tc_words();                                     // size of that buffer in 32-bit words
void query_persist_occupancy(int per_sm[4]);               // current device; index = SMALL + 2 * TCMLP
cudaError_t launch_step(const DevConfig& c, const StepLaunch& l, int rng, cudaStream_t st);
cudaError_t launch_rollout_policy(const DevConfig& c, const float* obs, const uint8_t* mask, const uint32_t* packed, uint64_t seed,
                                  uint32_t t, const uint32_t* t_dev, int deterministic, int64_t* actions, float* logp, float* value,
                                  cudaStream_t st);                                            // Env_3 (29, 22): the policy half alone
cudaError_t launch_pack_fused(const float* params, uint32_t* packed, cudaStream_t st);   // flat fp32 parameters -> MSORT_ROLLOUT_WEIGHTS words
cudaError_t launch_tc_logits(const float* obs13, const uint32_t* tcw, long long n, float* logits, int sm_count, cudaStream_t st);
cudaError_t launch_reset(const DevConfig& c, void* state, const uint8_t* which, const uint8_t* first_pattern,
                         float* obs, uint8_t* mask, uint32_t reset_flags, cudaStream_t st);
cudaError_t launch_observe(const DevConfig& c, const void* state, float* obs, uint8_t* mask, int after_shift, cudaStream_t st);
cudaError_t launch_sample(const DevConfig& c, const uint8_t* mask, int64_t* actions, uint64_t seed, uint32_t t,
                          cudaStream_t st);
cudaError_t launch_generate_streams(const DevConfig& c, uint32_t episode, uint32_t first_step, uint32_t num_steps, uint32_t* input_counts,
                                    double* noise_u, uint32_t* draw_words, uint8_t* first_pattern, cudaStream_t st);
cudaError_t launch_widen_actions(const uint8_t* in, int64_t* out, long long n, cudaStream_t st);
cudaError_t launch_pack_flags(int kind, const uint8_t* mask, const uint8_t* terminated, uint16_t* flags, long long n, cudaStream_t st);
cudaError_t launch_rule_actions(const DevConfig& c, const void* state, int after_shift, int64_t* actions, cudaStream_t st);
cudaError_t launch_policy_act(const DevConfig& c, const float* obs, const uint8_t* mask, const float* packed,
                              int D, int A, uint64_t seed, uint32_t t, const uint32_t* t_dev, int deterministic, int64_t* actions,
                              float* logp, float* value, int sm_count, cudaStream_t st);
cudaError_t launch_policy_eval(const DevConfig& c, int D, int A, long long rows, const float* obs, long long ld_obs, const uint8_t* mask,
                               long long ld_mask, const float* packed, uint64_t seed, uint32_t t, int deterministic, int64_t* actions,
                               float* logp, float* value, int sm_count, cudaStream_t st);
cudaError_t prepare_policy_kernels();                      // current device: opt-in shared-memory sizes of the policy kernels
cudaError_t launch_export(const DevConfig& c, const void* state, msort_env_state_t* out, cudaStream_t st);
cudaError_t launch_gather(const DevConfig& c, const void* state, const int64_t* env_ids, long long count,
                          msort_env_state_t* out, cudaStream_t st);
cudaError_t launch_import(const DevConfig& c, void* state, const msort_env_state_t* in, cudaStream_t st);
cudaError_t launch_stats(const DevConfig& c, const void* state, double* out16, int sm_count, cudaStream_t st);

// ---- msort_ppo.cu: the update half of the GPU-resident MaskablePPO loop
int ppo_param_count(int D, int A);
int ppo_scratch_floats(int D, int A);
cudaError_t ppo_forward(const msort_ppo_batch_t& b, const float* params, float* logp_out, float* value_out, cudaStream_t st);
cudaError_t ppo_gradient(const msort_ppo_batch_t& b, const msort_ppo_hparams_t& hp, const float* params, float* grads,
                         const int64_t* idx, long long first, long long count, float* adv_stats, float* stats, cudaStream_t st);
cudaError_t ppo_adam(float* params, float* grads, float* m, float* v, int* step, int n, const msort_ppo_hparams_t& hp, cudaStream_t st);
cudaError_t ppo_gae(int T, long long n, const float* rew, const float* val, const uint8_t* done, const float* last_val, float gamma,
                    float lam, float* adv, float* ret, cudaStream_t st);

}  // namespace msort
