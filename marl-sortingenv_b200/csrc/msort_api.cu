// msort_api.cu — the C ABI of libmsort.so (include/msort.h): argument validation, config
// digestion (host float64 arithmetic identical to the reference's), handle management and
// kernel launches.  No torch types, no exceptions across the boundary, no CPU fallback.
#include <cuda_runtime.h>
#include <algorithm>

#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>

#include "msort_device.cuh"
#include "msort_launch.h"

using namespace msort;

struct msort_handle {
  msort_config_t cfg;
  DevConfig dev;
  int device;
  int sm_count;
  double* lut_dev;    // owned device constant (tiny, allocated at create): the kSortLut-entry float64 sorting-reward table
  bool policy_set;
  int step_variant = MSORT_STEP_NONE;       // instantiation of the last step launch (msort_step_variant)
  bool imported = false;                    // msort_import_state since the last full reset: stage contents are arbitrary, so the
                                            // HOT kernel's host-proved bound of twelve draws per station does not hold
  float policy_host[MSORT_POLICY_WEIGHTS];  // host copy in the kernel's paired layout (pack_policy_pairs): travels to the step kernel as a kernel parameter
  uint32_t policy_tc_host[kTcWords];        // the same policy packed for the tensor-core path (pack_policy_tc); source of the upload
  uint32_t* policy_tc_dev = nullptr;        // owned device copy (11 KB, allocated by the first msort_set_policy)
  bool policy_tc_ok = false;                // the policy fits the fp16 split (else the FFMA2 kernel evaluates it)
  bool policy_tc_enabled = true;            // msort_set_option(MSORT_OPT_TENSOR_POLICY) != 0
  bool policy_tc_split = MSORT_TENSOR_POLICY_DEFAULT == 2;   // ... == 2: policy kernel + policy-free step kernel instead of the one fused kernel
  uint8_t* mode_scratch = nullptr;          // owned: one byte per env (n_pad), the split form's sort modes (allocated with the tensor-core policy)
  const uint32_t* draw_counter = nullptr;   // MSORT_OPT_DRAW_COUNTER: device-side offset of msort_policy_act's draw index
  int persist_per_sm[4] = {1, 1, 1, 1};     // resident CTAs per SM of the persistent Env_2 kernels on this handle's device
  // msort_step_host: the library's own streams / events (created by the first call, on the handle's device)
  static constexpr int kHostStreams = 4;
  cudaStream_t host_stream[kHostStreams] = {};
  cudaEvent_t host_start = nullptr, host_done[kHostStreams] = {};
  bool host_ready = false;
  int64_t launches;
};

static thread_local char g_err[512] = "";

static int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

static int cuda_fail(cudaError_t e, const char* where) {
  return fail(MSORT_E_CUDA, "%s: %s (%s)", where, cudaGetErrorString(e), cudaGetErrorName(e));
}

#define MSORT_TRY_CUDA(expr, where)                 \
  do {                                              \
    cudaError_t e__ = (expr);                       \
    if (e__ != cudaSuccess) return cuda_fail(e__, where); \
  } while (0)

static bool aligned(const void* p, size_t a) { return (reinterpret_cast<uintptr_t>(p) & (a - 1)) == 0; }

extern "C" int msort_abi_version(void) { return MSORT_ABI_VERSION; }
extern "C" const char* msort_last_error(void) { return g_err; }

extern "C" int msort_default_config(int env_kind, msort_config_t* cfg) {
  if (!cfg) return fail(MSORT_E_INVALID, "msort_default_config: cfg is NULL");
  if (env_kind < MSORT_ENV_SORT || env_kind > MSORT_ENV_MONO)
    return fail(MSORT_E_INVALID, "msort_default_config: unknown env kind %d", env_kind);
  memset(cfg, 0, sizeof(*cfg));
  cfg->struct_size = sizeof(*cfg);
  cfg->env_kind = env_kind;
  cfg->num_envs = 1;
  cfg->max_steps = 50;                                  // ctor default (env_1_sort.py:19)
  cfg->flags = MSORT_F_ACTION_MASKING | MSORT_F_AUTO_RESET;
  cfg->rng_mode = MSORT_RNG_PHILOX;
  cfg->input_batch_size = 100;                          // config.yml simulation.input_batch_size
  cfg->steps_per_pattern = 20;                          // input_generator.py:15 (env_super.py:375)
  const double ratios[2][4] = {{0.40, 0.15, 0.35, 0.10}, {0.15, 0.40, 0.10, 0.35}};  // input_generator.py:17-20
  for (int p = 0; p < 2; ++p)
    for (int m = 0; m < 4; ++m) cfg->pattern_counts[p][m] = (int32_t)std::floor(ratios[p][m] * 100);
  for (int m = 0; m < 4; ++m) { cfg->baseline_accuracy[m] = 0.75; cfg->quality_threshold[m] = 0.9; }
  cfg->boost = 0.5;
  cfg->noise = 0.05;
  cfg->stage_capacity = 100;
  cfg->press_time[0] = 12; cfg->press_time[1] = 15;
  cfg->container_capacity = 700;
  cfg->bale_size = 200;
  cfg->bale_remainder_threshold = 0.5;
  cfg->purity_theta = 0.80; cfg->purity_scaling = 2.0; cfg->tanh_temperature = 0.5;
  cfg->overflow_penalty_catastrophic = -1.0; cfg->overflow_penalty_severe = -0.5; cfg->overflow_penalty_mild = -0.2;
  cfg->bale_efficiency_factor = 1.0; cfg->max_state_reward = 0.5;
  cfg->overflow_termination_penalty = -10.0;
  return MSORT_OK;
}

// smallest integer level L in [0, hi] with (double)L / cap > thr, in the reference's own float64
// arithmetic (`fill_ratio = level / container_max`, env_super.py:1020-1027); hi+1 if none.
static int level_threshold(int cap, double thr, int hi) {
  for (int L = 0; L <= hi; ++L)
    if ((double)L / (double)cap > thr) return L;
  return hi + 1;
}

// Philox4x32 key schedule (key += W per round), precomputed because the key is launch-uniform.
static void set_round_keys(DevConfig& d, uint64_t seed) {
  unsigned k0 = (unsigned)(seed & 0xffffffffu), k1 = (unsigned)(seed >> 32);
  for (int r = 0; r < 10; ++r) {
    d.rk[2 * r] = k0; d.rk[2 * r + 1] = k1;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
}

// Python-float round(x, 2) (float.__round__: correctly rounded on the exact decimal value of the double) — what
// glibc's "%.2f" prints.  The reference applies it to an EMPTY container: purity = round(threshold, 2)
// (env_super.py:788-789) and purity difference round(purity - threshold, 2) (:222-225), both on plain Python floats.
static double py_round2(double x) {
  char buf[64];
  snprintf(buf, sizeof buf, "%.2f", x);
  return strtod(buf, nullptr);
}

// float64 restatement of the reference's obs purity difference, used to validate the fast path
static float pdiff_reference(int k, double qthr) {
  double d = (double)k / 100.0 - qthr;
  return (float)(std::rint(d * 100.0) / 100.0);
}

static int digest_config(const msort_config_t& c, DevConfig& d) {
  if (c.struct_size != sizeof(msort_config_t))
    return fail(MSORT_E_INVALID, "config struct_size %u != %zu", c.struct_size, sizeof(msort_config_t));
  if (c.env_kind < MSORT_ENV_SORT || c.env_kind > MSORT_ENV_MONO) return fail(MSORT_E_INVALID, "unknown env kind %d", c.env_kind);
  if (c.num_envs <= 0) return fail(MSORT_E_INVALID, "num_envs must be > 0");
  if (c.num_envs > (1ll << 31) - kTile) return fail(MSORT_E_UNSUPPORTED, "num_envs too large for one device");
  if (c.global_env_offset < 0 || c.global_env_offset + c.num_envs > (1ll << 48))
    return fail(MSORT_E_INVALID, "global env ids must fit 48 bits");
  if (c.max_steps <= 0) return fail(MSORT_E_INVALID, "max_steps must be > 0");
  if (c.rng_mode != MSORT_RNG_PHILOX && c.rng_mode != MSORT_RNG_REPLAY) return fail(MSORT_E_INVALID, "unknown rng mode");
  if (c.input_batch_size < 0 || c.input_batch_size > 255)
    return fail(MSORT_E_UNSUPPORTED, "input_batch_size %d outside [0,255] (stage counts are stored as bytes)", c.input_batch_size);
  if (c.steps_per_pattern < 1 || c.steps_per_pattern > 255) return fail(MSORT_E_UNSUPPORTED, "steps_per_pattern outside [1,255]");
  if (c.container_capacity <= 0 || c.bale_size <= 0 || c.stage_capacity <= 0) return fail(MSORT_E_INVALID, "capacities must be > 0");
  if (c.bale_size > (1 << 22)) return fail(MSORT_E_UNSUPPORTED, "bale_size too large");
  for (int p = 0; p < 2; ++p)
    if (c.press_time[p] < 1 || c.press_time[p] > 255) return fail(MSORT_E_UNSUPPORTED, "press_time outside [1,255]");
  if (!(c.noise >= 0.0)) return fail(MSORT_E_INVALID, "noise must be >= 0");
  if (c.tanh_temperature == 0.0) return fail(MSORT_E_INVALID, "tanh_temperature must be non-zero");

  memset(&d, 0, sizeof(d));
  d.n = c.num_envs;
  d.n_pad = (c.num_envs + kTile - 1) / kTile * kTile;
  d.gid0 = c.global_env_offset;
  d.kind = c.env_kind;
  d.max_steps = c.max_steps;
  d.flags = c.flags;
  d.rng_mode = c.rng_mode;
  // State layout (msort_device.cuh): compact 16-bit container counts are safe only when every level
  // is provably < 2^16: at most input_batch_size units enter per step and auto-reset bounds the episode.
  d.layout = c.rng_mode == MSORT_RNG_REPLAY ? LAYOUT_REPLAY
           : ((c.flags & MSORT_F_AUTO_RESET) && (long long)c.input_batch_size * c.max_steps <= 65535ll) ? LAYOUT_COMPACT
           : LAYOUT_WIDE;
  set_round_keys(d, c.seed);
  d.batch = c.input_batch_size;
  d.spp = c.steps_per_pattern;
  int tot[2] = {0, 0};
  for (int p = 0; p < 2; ++p) {
    d.pat[p] = 0;
    for (int m = 0; m < 4; ++m) {
      int v = c.pattern_counts[p][m];
      if (v < 0 || v > 255) return fail(MSORT_E_INVALID, "pattern count outside [0,255]");
      d.pat[p] |= (unsigned)v << (8 * m);
      tot[p] += v;
    }
    if (tot[p] > c.input_batch_size) return fail(MSORT_E_INVALID, "pattern counts exceed input_batch_size");
  }
  if (tot[0] != tot[1]) return fail(MSORT_E_UNSUPPORTED, "patterns with different totals");
  d.pat_remainder = c.input_batch_size - tot[0];  // input_generator.py:52-55
  d.stage_cap = c.stage_capacity;
  d.cap = c.container_capacity;
  d.S = c.bale_size;
  d.press_time[0] = c.press_time[0];
  d.press_time[1] = c.press_time[1];
  d.lvl_cat = level_threshold(c.container_capacity, 1.0, 2 * c.container_capacity + 2);
  d.lvl_sev = level_threshold(c.container_capacity, 0.95, 2 * c.container_capacity + 2);
  d.lvl_mild = level_threshold(c.container_capacity, 0.90, 2 * c.container_capacity + 2);
  d.rem_new_bale = c.bale_size + 1;  // `rem > S*threshold` (env_super.py:675); rem < S always
  for (int r = 0; r <= c.bale_size; ++r)
    if ((double)r > (double)c.bale_size * c.bale_remainder_threshold) { d.rem_new_bale = r; break; }
  for (int m = 0; m < 4; ++m) {
    d.base_acc[m] = c.baseline_accuracy[m]; d.qthr[m] = c.quality_threshold[m];
    d.qthr_empty[m] = py_round2(c.quality_threshold[m]);
    d.pdiff_empty[m] = (float)py_round2(d.qthr_empty[m] - c.quality_threshold[m]);
    if (d.pdiff_empty[m] < -1.f) d.pdiff_empty[m] = -1.f;
    if (d.pdiff_empty[m] > 1.f) d.pdiff_empty[m] = 1.f;
  }
  d.boost = c.boost;
  d.noise_low = -c.noise;                 // numpy uniform(low, high): low + (high-low)*u
  d.noise_range = c.noise - d.noise_low;
  d.theta4 = 4.0 * c.purity_theta;
  d.c_sort = (c.purity_scaling / 4.0) / c.tanh_temperature;
  d.c_state = c.max_state_reward / (double)(5 * c.container_capacity);
  d.c_eff = 4.0 / (double)c.bale_size;
  d.pen_cat = c.overflow_penalty_catastrophic; d.pen_sev = c.overflow_penalty_severe; d.pen_mild = c.overflow_penalty_mild;
  d.bef = c.bale_efficiency_factor; d.ovf_pen = c.overflow_termination_penalty;
  d.inv_cap = 1.0f / (float)c.container_capacity;
  d.inv_stage = 1.0f / (float)c.stage_capacity;
  d.inv_pt[0] = 1.0f / (float)c.press_time[0];
  d.inv_pt[1] = 1.0f / (float)c.press_time[1];
  // obs purity difference: (k - 100*qthr)/100 in float32 when that reproduces the float64 pipeline
  d.fast_pdiff = 1;
  for (int m = 0; m < 4; ++m) {
    d.qthr100[m] = (int)std::lrint(c.quality_threshold[m] * 100.0);
    for (int k = 0; k <= 100 && d.fast_pdiff; ++k) {
      float ref = pdiff_reference(k, c.quality_threshold[m]);
      float fast = (float)(k - d.qthr100[m]) * 0.01f;
      if (std::fabs((double)ref - (double)fast) > 1e-7 + 2e-6 * std::fabs((double)ref)) d.fast_pdiff = 0;
    }
  }
  // purity ties (purity_k in msort_device.cuh): outcome of the float64 pipeline rint(fl(fl(tr/tot)*100)) when
  // 100*tr/tot is exactly k + 1/2, i.e. tr/tot == (2k+1)/200 — a function of k alone
  for (int k = 0; k < 100; ++k) {
    volatile double q = (double)(2 * k + 1) / 200.0;
    volatile double y = q * 100.0;
    if ((int)std::rint(y) == k + 1) d.tie_up[k >> 5] |= 1u << (k & 31);
  }
  // FAST step-kernel instantiation (PHILOX mode): legal when, in the kernel's own float64 operation order,
  //  * a boosted accuracy clip((base+boost) + (low + range*u), 0, 1) is exactly 1.0 for every u in [0,1)
  //    (the noise term is monotone in u, so u = 0 is the worst case),
  //  * an unboosted accuracy base + (low + range*u) stays inside [0,1] (the clip is the identity),
  //  * and a whole batch fits 7 bits (the packed-byte class selection of the redistribution draws).
  //  * no input remainder (every stage holds one of the two patterns or nothing), whole-percent quality
  //    thresholds inside [0,1], and overflow penalties ordered severe <= mild, so that the press penalty
  //    depends on the fullest container alone.
  d.fast = c.rng_mode == MSORT_RNG_PHILOX && c.input_batch_size <= 127 && d.pat_remainder == 0 && d.fast_pdiff &&
           c.overflow_penalty_severe <= c.overflow_penalty_mild;
  for (int m = 0; m < 4; ++m) {
    if (std::fabs(c.quality_threshold[m] * 100.0 - (double)d.qthr100[m]) > 1e-9 || d.qthr100[m] < 0 || d.qthr100[m] > 100) d.fast = 0;
  }
  d.pen_sev0 = std::min(0.0, c.overflow_penalty_severe);
  d.pen_mild0 = std::min(0.0, c.overflow_penalty_mild);
  d.small_lv = d.layout == LAYOUT_COMPACT && (long long)c.input_batch_size * c.max_steps <= 8192ll;
  // at most twelve mis-sorted units per station: false(t) = t - rint(t*acc) with acc >= base + noise_low (rounding is
  // monotone), for every stage count t a station can see (0 .. largest pattern count; draws only lower it)
  d.one_block = 1;
  for (int m = 0; m < 4 && d.one_block; ++m) {
    int tmax = std::max(c.pattern_counts[0][m], c.pattern_counts[1][m]);
    volatile double amin = d.base_acc[m] + d.noise_low;
    if (amin < 0.0) amin = 0.0;
    for (int t = 0; t <= tmax; ++t) {
      volatile double prod = (double)t * amin;
      if (t - (int)std::rint(prod) > 12) { d.one_block = 0; break; }
    }
  }
  d.S_magic = (unsigned)((1ull << 32) / (unsigned long long)c.bale_size) + 1u;
  // observation tables for the three possible stage contents (same float32 operations as obs_belt / obs_sorting)
  for (int w = 0; w < 3; ++w) {
    const unsigned v = w < 2 ? d.pat[w] : 0u;
    int cnt[4], bt = 0;
    for (int m = 0; m < 4; ++m) { cnt[m] = (int)((v >> (8 * m)) & 0xffu); bt += cnt[m]; }
    volatile float inv_bt = bt > 0 ? 1.0f / (float)bt : 0.f;
    volatile float occ = (float)bt * 0.01f;
    d.obs_belt_tab[w][0] = std::fmin((float)occ, 1.f);
    for (int m = 0; m < 4; ++m) {
      volatile float pb = (float)cnt[m] * inv_bt, ps = (float)cnt[m] * d.inv_stage;
      d.obs_belt_tab[w][1 + m] = std::fmin((float)pb, 1.f);
      d.obs_sort_tab[w][m] = std::fmin((float)ps, 1.f);
    }
  }
  for (int m = 0; m < 4 && d.fast; ++m) {
    volatile double boosted = d.base_acc[m] + d.boost;
    volatile double lo_b = boosted + d.noise_low;
    volatile double lo_u = d.base_acc[m] + d.noise_low;
    volatile double span = d.noise_range * 1.0;
    volatile double nz_hi = d.noise_low + span;
    volatile double hi_u = d.base_acc[m] + nz_hi;
    if (!(lo_b >= 1.0) || !(lo_u >= 0.0) || !(hi_u <= 1.0)) d.fast = 0;
  }
  return MSORT_OK;
}

extern "C" int msort_create(const msort_config_t* cfg, int device, msort_t** out) {
  if (!cfg || !out) return fail(MSORT_E_INVALID, "msort_create: NULL argument");
  *out = nullptr;
  DevConfig d;
  int rc = digest_config(*cfg, d);
  if (rc != MSORT_OK) return rc;
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count == 0) return fail(MSORT_E_NO_DEVICE, "no CUDA device (%s); there is no CPU fallback", cudaGetErrorString(e));
  if (device < 0 || device >= count) return fail(MSORT_E_INVALID, "device %d out of range (%d devices)", device, count);
  cudaDeviceProp prop;
  MSORT_TRY_CUDA(cudaGetDeviceProperties(&prop, device), "cudaGetDeviceProperties");
  if (prop.major != 10)
    return fail(MSORT_E_NO_DEVICE, "device %d is sm_%d%d; libmsort is built for sm_100a (B200) only", device, prop.major, prop.minor);
  int prev_device = 0;
  MSORT_TRY_CUDA(cudaGetDevice(&prev_device), "cudaGetDevice");
  MSORT_TRY_CUDA(cudaSetDevice(device), "cudaSetDevice");
  msort_handle* h = new (std::nothrow) msort_handle();
  if (!h) return fail(MSORT_E_INVALID, "out of host memory");
  h->cfg = *cfg;
  h->dev = d;
  h->device = device;
  h->sm_count = prop.multiProcessorCount;
  h->dev.sm_count = prop.multiProcessorCount;
  h->lut_dev = nullptr;
  h->policy_set = false;
  h->launches = 0;
  if (d.kind == MSORT_ENV_PRESS) query_persist_occupancy(h->persist_per_sm);
  e = prepare_policy_kernels();
  if (e != cudaSuccess) { cudaSetDevice(prev_device); delete h; return cuda_fail(e, "cudaFuncSetAttribute(policy kernels)"); }   // per handle, on its own device, outside any capture
  cudaSetDevice(prev_device);  // launches run on the caller's current device, which must be `device`
  // sorting-reward table (see sort_reward_f64 in msort_device.cuh): same formula, host float64
  {
    double lut[kSortLut];
    for (int kt = 0; kt < kSortLut; ++kt) lut[kt] = std::tanh(((double)kt * 0.01 - h->dev.theta4) * h->dev.c_sort);
    cudaSetDevice(device);
    e = cudaMalloc(&h->lut_dev, sizeof(lut));
    if (e == cudaSuccess) e = cudaMemcpy(h->lut_dev, lut, sizeof(lut), cudaMemcpyHostToDevice);
    cudaSetDevice(prev_device);
    if (e != cudaSuccess) { delete h; return cuda_fail(e, "cudaMalloc(sort lut)"); }
    h->dev.sort_lut = h->lut_dev;
    // the table is exact only when every threshold is a whole percent; otherwise the kernel evaluates float64
    for (int m = 0; m < 4; ++m)
      if (std::fabs(cfg->quality_threshold[m] * 100.0 - (double)h->dev.qthr100[m]) > 1e-9) h->dev.fast_pdiff = 0;
  }
  *out = h;
  return MSORT_OK;
}

extern "C" int msort_destroy(msort_t* h) {
  if (!h) return MSORT_OK;
  if (h->lut_dev) cudaFree(h->lut_dev);
  if (h->policy_tc_dev) cudaFree(h->policy_tc_dev);
  if (h->mode_scratch) cudaFree(h->mode_scratch);
  if (h->host_ready) {
    for (int k = 0; k < msort_handle::kHostStreams; ++k) { cudaStreamDestroy(h->host_stream[k]); cudaEventDestroy(h->host_done[k]); }
    cudaEventDestroy(h->host_start);
  }
  delete h;
  return MSORT_OK;
}

extern "C" size_t msort_state_bytes(const msort_t* h) { return h ? (size_t)kPlanes * (size_t)h->dev.n_pad * 16u : 0; }
extern "C" int msort_obs_dim(const msort_t* h) {
  if (!h) return 0;
  return h->dev.kind == MSORT_ENV_SORT ? MSORT_OBS_DIM_SORT : (h->dev.kind == MSORT_ENV_PRESS ? MSORT_OBS_DIM_PRESS : MSORT_OBS_DIM_MONO);
}
extern "C" int msort_num_actions(const msort_t* h) {
  if (!h) return 0;
  return h->dev.kind == MSORT_ENV_SORT ? MSORT_NUM_ACTIONS_SORT : (h->dev.kind == MSORT_ENV_PRESS ? MSORT_NUM_ACTIONS_PRESS : MSORT_NUM_ACTIONS_MONO);
}
extern "C" int64_t msort_launch_count(const msort_t* h) { return h ? h->launches : 0; }
extern "C" int msort_step_variant(const msort_t* h) { return h ? h->step_variant : MSORT_STEP_NONE; }

extern "C" int msort_set_seed(msort_t* h, uint64_t seed) {
  if (!h) return fail(MSORT_E_INVALID, "msort_set_seed: NULL handle");
  h->cfg.seed = seed;
  set_round_keys(h->dev, seed);
  return MSORT_OK;
}

extern "C" int msort_set_option(msort_t* h, int option, int64_t value) {
  if (!h) return fail(MSORT_E_INVALID, "msort_set_option: NULL handle");
  switch (option) {
    case MSORT_OPT_TENSOR_POLICY:
      if (value < 0 || value > 2) return fail(MSORT_E_INVALID, "msort_set_option: MSORT_OPT_TENSOR_POLICY must be 0, 1 or 2");
      h->policy_tc_enabled = value != 0; h->policy_tc_split = value == 2;
      return MSORT_OK;
    case MSORT_OPT_DRAW_COUNTER:
      if (value & 3) return fail(MSORT_E_INVALID, "msort_set_option: MSORT_OPT_DRAW_COUNTER must be a 4-byte aligned device pointer");
      h->draw_counter = reinterpret_cast<const uint32_t*>((uintptr_t)value);
      return MSORT_OK;
    case MSORT_OPT_PERSIST_CTAS:
      if (value < 1 || value > 32) return fail(MSORT_E_INVALID, "msort_set_option: MSORT_OPT_PERSIST_CTAS must be in [1, 32]");
      for (int k = 0; k < 4; ++k) h->persist_per_sm[k] = (int)value;
      return MSORT_OK;
    default: return fail(MSORT_E_INVALID, "msort_set_option: unknown option %d", option);
  }
}

extern "C" int msort_get_option(const msort_t* h, int option, int64_t* value) {
  if (!h || !value) return fail(MSORT_E_INVALID, "msort_get_option: NULL argument");
  switch (option) {
    case MSORT_OPT_TENSOR_POLICY:
      *value = (h->policy_tc_enabled && h->policy_set && h->policy_tc_ok) ? ((h->policy_tc_split && h->mode_scratch) ? 2 : 1) : 0;
      return MSORT_OK;
    case MSORT_OPT_PERSIST_CTAS:
      *value = h->persist_per_sm[(h->dev.small_lv ? 1 : 0) + ((h->policy_tc_enabled && h->policy_set && h->policy_tc_ok) ? 2 : 0)];
      return MSORT_OK;
    default: return fail(MSORT_E_INVALID, "msort_get_option: unknown option %d", option);
  }
}

extern "C" int msort_set_flags(msort_t* h, uint32_t flags) {
  if (!h) return fail(MSORT_E_INVALID, "msort_set_flags: NULL handle");
  if (h->dev.layout == LAYOUT_COMPACT && !(flags & MSORT_F_AUTO_RESET))
    return fail(MSORT_E_UNSUPPORTED, "msort_set_flags: this handle was created with auto-reset and uses the compact "
                "16-bit state layout; create it without MSORT_F_AUTO_RESET to step past max_steps");
  h->cfg.flags = flags;
  h->dev.flags = flags;
  return MSORT_OK;
}

extern "C" int msort_reset(msort_t* h, void* state, const uint8_t* which, const uint8_t* first_pattern, float* obs,
                           uint8_t* mask, uint32_t reset_flags, void* stream) {
  if (!h || !state) return fail(MSORT_E_INVALID, "msort_reset: NULL handle/state");
  if (!aligned(state, 16)) return fail(MSORT_E_INVALID, "msort_reset: state must be 16-byte aligned");
  if ((obs && !aligned(obs, 16)) || (mask && !aligned(mask, 16)))
    return fail(MSORT_E_INVALID, "msort_reset: obs / mask must be 16-byte aligned (tiles leave as 16-byte vectors)");
  MSORT_TRY_CUDA(launch_reset(h->dev, state, which, first_pattern, obs, mask, reset_flags, (cudaStream_t)stream), "reset kernel");
  if (!which) h->imported = false;          // every env holds a fresh plant again
  h->launches += 1;
  return MSORT_OK;
}

// `first` / `count`: the env range [first, first + count) the launch covers (whole batch: 0 / N).  Every pointer is the
// whole-batch base; the range's slice of each buffer is taken here.
static int step_impl(msort_t* h, long long first, long long count, void* state, const int64_t* actions, float* obs, float* reward,
                     uint8_t* terminated, uint8_t* mask, const msort_info_out_t* info, const msort_replay_t* replay, void* stream,
                     const FusedLaunch* fused = nullptr) {
  if (!h || !state || !actions || !obs || !reward || !terminated)
    return fail(MSORT_E_INVALID, "msort_step: NULL required argument");
  if (!aligned(state, 16)) return fail(MSORT_E_INVALID, "msort_step: state must be 16-byte aligned");
  if (!aligned(actions, 8)) return fail(MSORT_E_INVALID, "msort_step: actions must be int64-aligned");
  if (!aligned(obs, 16)) return fail(MSORT_E_INVALID, "msort_step: obs must be 16-byte aligned (tiles leave by TMA bulk copies)");
  if (!aligned(reward, 4)) return fail(MSORT_E_INVALID, "msort_step: reward must be float-aligned");
  if (mask && !aligned(mask, 16)) return fail(MSORT_E_INVALID, "msort_step: mask must be 16-byte aligned (tiles leave by TMA bulk copies)");
  if (info && info->struct_size != sizeof(msort_info_out_t)) return fail(MSORT_E_INVALID, "msort_step: bad info struct_size");
  if (info && info->stats && !aligned(info->stats, 8)) return fail(MSORT_E_INVALID, "msort_step: stats misaligned");
  if (h->cfg.rng_mode == MSORT_RNG_REPLAY) {
    if (!replay) return fail(MSORT_E_REPLAY, "msort_step: REPLAY mode needs a replay descriptor");
    if (replay->struct_size != sizeof(msort_replay_t)) return fail(MSORT_E_INVALID, "msort_step: bad replay struct_size");
    if (!replay->noise_u || !replay->redis_u || replay->redis_len <= 0)
      return fail(MSORT_E_REPLAY, "msort_step: REPLAY mode needs noise_u and redis_u streams");
    if (!aligned(replay->noise_u, 16) || !aligned(replay->redis_u, 8))
      return fail(MSORT_E_INVALID, "msort_step: replay streams misaligned");
    if (h->dev.kind == MSORT_ENV_SORT && !replay->press_choice)
      return fail(MSORT_E_REPLAY, "msort_step: Env_1 REPLAY mode needs the press_choice stream");
    if (replay->input_counts && !aligned(replay->input_counts, 4)) return fail(MSORT_E_INVALID, "msort_step: input_counts misaligned");
  } else if (replay) {
    return fail(MSORT_E_INVALID, "msort_step: replay descriptor given but the handle is in PHILOX mode");
  }
  if (h->dev.kind == MSORT_ENV_PRESS && (h->dev.flags & MSORT_F_SORT_POLICY_MLP) && !h->policy_set &&
      !(replay && replay->sort_mode))
    return fail(MSORT_E_INVALID, "msort_step: embedded sort policy requested but msort_set_policy() was never called");
  const int D = msort_obs_dim(h), A = msort_num_actions(h);
  DevConfig d = h->dev;
  msort_info_out_t sub;
  if (first != 0 || count != h->dev.n) {            // a sub-range: same plane stride, shifted bases, global ids continue
    d.n = count; d.gid0 += first;
    state = (uint4*)state + first; actions += first; obs += first * D; reward += first; terminated += first;
    if (mask) mask += first * A;
    if (info) {
      sub = *info;
      if (sub.action) sub.action += first;
      if (sub.overflow) sub.overflow += first;
      if (sub.overflow_material) sub.overflow_material += first;
      if (sub.sort_mode) sub.sort_mode += first;
      if (sub.press_action) sub.press_action += first;
      if (sub.invalid_action) sub.invalid_action += first;
      if (sub.terminal_obs) sub.terminal_obs += first * D;
      if (sub.episode_return) sub.episode_return += first;
      if (sub.episode_length) sub.episode_length += first;
      if (sub.reward_sort) sub.reward_sort += first;
      if (sub.reward_press) sub.reward_press += first;
      if (sub.sorted_true) sub.sorted_true += first;
      info = &sub;
    }
  }
  StepLaunch l{state, actions, obs, reward, terminated, mask, info, replay, &h->step_variant, h->imported ? 0 : 1,
               h->policy_set ? h->policy_host : nullptr,
               (h->policy_set && h->policy_tc_ok && h->policy_tc_enabled) ? h->policy_tc_dev : nullptr, h->persist_per_sm};
  l.fused = fused;
  if (l.policy_tc && h->policy_tc_split && h->mode_scratch) l.mode_scratch = h->mode_scratch + first;   // this launch's slice of the handle's scratch
  if (fused) {
    cudaError_t e = launch_step(d, l, h->cfg.rng_mode, (cudaStream_t)stream);
    if (e == cudaErrorNotSupported)
      return fail(MSORT_E_UNSUPPORTED, "msort_rollout_step: the fused kernel exists for Env_3 in PHILOX mode with action masking and auto-reset on, "
                                       "no overflow check, a mask output and no per-step info arrays; use msort_policy_act + msort_step");
    MSORT_TRY_CUDA(e, "fused rollout kernel");
    h->launches += 1;
    return MSORT_OK;
  }
  MSORT_TRY_CUDA(launch_step(d, l, h->cfg.rng_mode, (cudaStream_t)stream), "step kernel");
  h->launches += h->step_variant == MSORT_STEP_HOT_TENSOR_SPLIT ? 2 : 1;
  return MSORT_OK;
}

extern "C" int msort_step(msort_t* h, void* state, const int64_t* actions, float* obs, float* reward, uint8_t* terminated,
                          uint8_t* mask, const msort_info_out_t* info, const msort_replay_t* replay, void* stream) {
  if (!h) return fail(MSORT_E_INVALID, "msort_step: NULL required argument");
  return step_impl(h, 0, h->dev.n, state, actions, obs, reward, terminated, mask, info, replay, stream);
}

extern "C" int msort_step_range(msort_t* h, int64_t first_env, int64_t num_envs, void* state, const int64_t* actions, float* obs,
                                float* reward, uint8_t* terminated, uint8_t* mask, const msort_info_out_t* info, void* stream) {
  if (!h) return fail(MSORT_E_INVALID, "msort_step_range: NULL handle");
  if (h->cfg.rng_mode != MSORT_RNG_PHILOX) return fail(MSORT_E_INVALID, "msort_step_range: PHILOX mode only");
  if (first_env < 0 || num_envs <= 0 || first_env + num_envs > h->dev.n || first_env % kTile != 0)
    return fail(MSORT_E_INVALID, "msort_step_range: range must lie inside the batch and start on a multiple of %d envs", kTile);
  return step_impl(h, first_env, num_envs, state, actions, obs, reward, terminated, mask, info, nullptr, stream);
}

// ---------------------------------------------------------------- fused rollout step (Env_3)
extern "C" int msort_rollout_pack(const float* params, uint32_t* packed, void* stream) {
  if (!params || !packed) return fail(MSORT_E_INVALID, "msort_rollout_pack: NULL argument");
  if (!aligned(packed, 16)) return fail(MSORT_E_INVALID, "msort_rollout_pack: packed must be 16-byte aligned");
  MSORT_TRY_CUDA(launch_pack_fused(params, packed, (cudaStream_t)stream), "pack kernel");
  return MSORT_OK;
}

extern "C" int msort_rollout_policy(msort_t* h, int64_t first_env, int64_t num_envs, const float* obs, const uint8_t* mask,
                                    const uint32_t* packed, uint64_t seed, uint32_t t, int deterministic, int64_t* actions, float* logp,
                                    float* value, void* stream) {
  if (!h || !obs || !mask || !packed || !actions || !logp || !value) return fail(MSORT_E_INVALID, "msort_rollout_policy: NULL argument");
  if (h->dev.kind != MSORT_ENV_MONO) return fail(MSORT_E_UNSUPPORTED, "msort_rollout_policy: Env_3_Monolith (29 observations, 22 actions) only");
  if (!aligned(obs, 4) || !aligned(packed, 16) || !aligned(actions, 8) || !aligned(logp, 4) || !aligned(value, 4))
    return fail(MSORT_E_INVALID, "msort_rollout_policy: misaligned buffer");
  if (first_env < 0 || num_envs <= 0 || first_env + num_envs > h->dev.n || first_env % kTile != 0)
    return fail(MSORT_E_INVALID, "msort_rollout_policy: range must lie inside the batch and start on a multiple of %d envs", kTile);
  DevConfig d = h->dev;
  d.n = num_envs; d.gid0 += first_env;
  MSORT_TRY_CUDA(launch_rollout_policy(d, obs + first_env * 29, mask + first_env * 22, packed, seed, t, h->draw_counter, deterministic,
                                       actions + first_env, logp + first_env, value + first_env, (cudaStream_t)stream),
                 "rollout policy kernel");
  h->launches += 1;
  return MSORT_OK;
}

extern "C" int msort_rollout_step(msort_t* h, void* state, const int64_t* actions, float* obs, float* reward, uint8_t* terminated,
                                  uint8_t* mask, const msort_info_out_t* info, const uint32_t* packed, uint64_t seed, uint32_t t,
                                  int deterministic, int64_t* next_actions, float* next_logp, float* next_value, void* stream) {
  if (!h || !packed || !next_actions || !next_logp || !next_value || !mask)
    return fail(MSORT_E_INVALID, "msort_rollout_step: NULL argument");
  if (!aligned(packed, 16) || !aligned(next_actions, 8) || !aligned(next_logp, 4) || !aligned(next_value, 4))
    return fail(MSORT_E_INVALID, "msort_rollout_step: misaligned buffer");
  if (h->dev.kind != MSORT_ENV_MONO || h->cfg.rng_mode != MSORT_RNG_PHILOX)
    return fail(MSORT_E_UNSUPPORTED, "msort_rollout_step: Env_3_Monolith in PHILOX mode only");
  const FusedLaunch f{packed, next_actions, next_logp, next_value, seed, t, h->draw_counter, deterministic};
  return step_impl(h, 0, h->dev.n, state, actions, obs, reward, terminated, mask, info, nullptr, stream, &f);
}

// ---------------------------------------------------------------- host-buffer step
// device scratch layout (every section 256-byte aligned): actions i64 | actions u8 | flag words
namespace {
struct HostScratch { size_t act64, act8, flags, total; };
HostScratch host_scratch_layout(const msort_t* h) {
  const size_t n = (size_t)h->dev.n_pad;
  auto up = [](size_t x) { return (x + 255) / 256 * 256; };
  HostScratch s;
  size_t o = 0;
  s.act64 = o; o += up(n * 8);
  s.act8 = o; o += up(n);
  s.flags = o; o += up(n * 2);
  s.total = o;
  return s;
}
}  // namespace

extern "C" size_t msort_host_scratch_bytes(const msort_t* h) { return h ? host_scratch_layout(h).total : 0; }

extern "C" int msort_step_host(msort_t* h, void* state, void* scratch, const msort_host_io_t* io, const msort_info_out_t* info,
                               void* stream) {
  if (!h || !state || !scratch || !io) return fail(MSORT_E_INVALID, "msort_step_host: NULL argument");
  if (io->struct_size != sizeof(msort_host_io_t)) return fail(MSORT_E_INVALID, "msort_step_host: bad io struct_size");
  if ((io->actions_u8 != nullptr) == (io->actions_i64 != nullptr))
    return fail(MSORT_E_INVALID, "msort_step_host: exactly one of actions_u8 / actions_i64 must be given");
  if (!io->obs || !io->reward || !io->flags) return fail(MSORT_E_INVALID, "msort_step_host: NULL output buffer");
  if (!io->dev_obs || !io->dev_reward || !io->dev_terminated || !io->dev_mask)
    return fail(MSORT_E_INVALID, "msort_step_host: NULL device output buffer");
  if (!aligned(scratch, 256)) return fail(MSORT_E_INVALID, "msort_step_host: scratch must be 256-byte aligned");
  if (h->cfg.rng_mode != MSORT_RNG_PHILOX) return fail(MSORT_E_INVALID, "msort_step_host: PHILOX mode only");
  cudaStream_t user = (cudaStream_t)stream;
  if (!h->host_ready) {
    MSORT_TRY_CUDA(cudaEventCreateWithFlags(&h->host_start, cudaEventDisableTiming), "cudaEventCreate");
    for (int k = 0; k < msort_handle::kHostStreams; ++k) {
      MSORT_TRY_CUDA(cudaStreamCreateWithFlags(&h->host_stream[k], cudaStreamNonBlocking), "cudaStreamCreate");
      MSORT_TRY_CUDA(cudaEventCreateWithFlags(&h->host_done[k], cudaEventDisableTiming), "cudaEventCreate");
    }
    h->host_ready = true;
  }
  const HostScratch L = host_scratch_layout(h);
  char* base = (char*)scratch;
  int64_t* d_act = (int64_t*)(base + L.act64);
  uint8_t* d_act8 = (uint8_t*)(base + L.act8);
  float* d_obs = io->dev_obs;
  float* d_rew = io->dev_reward;
  uint8_t* d_term = io->dev_terminated;
  uint8_t* d_mask = io->dev_mask;
  uint16_t* d_flags = (uint16_t*)(base + L.flags);
  const long long n = h->dev.n;
  const int D = msort_obs_dim(h), A = msort_num_actions(h);
  // default 2: measured at 1 048 576 Env_3 envs (profiles/e2e_chunks_r02.txt) 1 / 2 / 4 / 8 / 16 / 32 ranges cost 2.44 / 2.43 /
  // 2.44 / 2.47 / 2.60 / 2.70 ms per step — the D2H engine is the bottleneck (128 MB at 56 GB/s = 2.26 ms), so all a second
  // range buys is the first range's copy running beside the second range's kernel; more ranges only add copy launches
  int chunks = io->chunks ? (int)io->chunks : 2;
  long long per = (n + chunks - 1) / chunks;
  per = (per + kTile - 1) / kTile * kTile;                       // ranges start on whole tiles
  MSORT_TRY_CUDA(cudaEventRecord(h->host_start, user), "cudaEventRecord");
  int used = 0, c = 0;
  for (long long lo = 0; lo < n; lo += per, ++c) {
    const long long cnt = std::min(per, n - lo);
    const int k = c % msort_handle::kHostStreams;
    cudaStream_t s = h->host_stream[k];
    if (c < msort_handle::kHostStreams) { MSORT_TRY_CUDA(cudaStreamWaitEvent(s, h->host_start, 0), "cudaStreamWaitEvent"); used = c + 1; }
    if (io->actions_u8) {
      MSORT_TRY_CUDA(cudaMemcpyAsync(d_act8 + lo, io->actions_u8 + lo, (size_t)cnt, cudaMemcpyHostToDevice, s), "H2D actions");
      MSORT_TRY_CUDA(launch_widen_actions(d_act8 + lo, d_act + lo, cnt, s), "widen kernel");
      h->launches += 1;
    } else {
      MSORT_TRY_CUDA(cudaMemcpyAsync(d_act + lo, io->actions_i64 + lo, (size_t)cnt * 8, cudaMemcpyHostToDevice, s), "H2D actions");
    }
    int rc = step_impl(h, lo, cnt, state, d_act, d_obs, d_rew, d_term, d_mask, info, nullptr, s);
    if (rc != MSORT_OK) return rc;
    MSORT_TRY_CUDA(launch_pack_flags(h->dev.kind, d_mask + lo * A, d_term + lo, d_flags + lo, cnt, s), "pack kernel");
    h->launches += 1;
    MSORT_TRY_CUDA(cudaMemcpyAsync(io->obs + lo * D, d_obs + lo * D, (size_t)cnt * D * 4, cudaMemcpyDeviceToHost, s), "D2H obs");
    MSORT_TRY_CUDA(cudaMemcpyAsync(io->reward + lo, d_rew + lo, (size_t)cnt * 4, cudaMemcpyDeviceToHost, s), "D2H reward");
    MSORT_TRY_CUDA(cudaMemcpyAsync(io->flags + lo, d_flags + lo, (size_t)cnt * 2, cudaMemcpyDeviceToHost, s), "D2H flags");
  }
  for (int k = 0; k < used; ++k) {
    MSORT_TRY_CUDA(cudaEventRecord(h->host_done[k], h->host_stream[k]), "cudaEventRecord");
    MSORT_TRY_CUDA(cudaStreamWaitEvent(user, h->host_done[k], 0), "cudaStreamWaitEvent");
  }
  return MSORT_OK;
}

extern "C" int msort_set_policy(msort_t* h, const float* weights, int weights_on_device, void* stream) {
  if (!h || !weights) return fail(MSORT_E_INVALID, "msort_set_policy: NULL argument");
  // The step kernel receives the weights as a kernel parameter, so the library keeps a host copy.  A
  // device-resident source is copied back once, here (the one place besides msort_sync_check that waits).
  float sb3[MSORT_POLICY_WEIGHTS];   // SB3 order, as include/msort.h documents
  if (weights_on_device) {
    MSORT_TRY_CUDA(cudaMemcpyAsync(sb3, weights, sizeof(sb3), cudaMemcpyDeviceToHost, (cudaStream_t)stream),
                   "cudaMemcpyAsync(policy)");
    MSORT_TRY_CUDA(cudaStreamSynchronize((cudaStream_t)stream), "cudaStreamSynchronize(policy)");
  } else {
    memcpy(sb3, weights, sizeof(sb3));
  }
  pack_policy_pairs(sb3, h->policy_host);
  // tensor-core form (Env_2's persistent HOT kernel): fp16-split weight tiles in a small device buffer the handle owns
  h->policy_tc_ok = pack_policy_tc(sb3, h->policy_tc_host);
  if (h->policy_tc_ok) {
    if (!h->policy_tc_dev) MSORT_TRY_CUDA(cudaMalloc(&h->policy_tc_dev, sizeof(h->policy_tc_host)), "cudaMalloc(tensor-core policy)");
    if (!h->mode_scratch) MSORT_TRY_CUDA(cudaMalloc(&h->mode_scratch, (size_t)h->dev.n_pad), "cudaMalloc(sort-mode scratch)");
    MSORT_TRY_CUDA(cudaMemcpyAsync(h->policy_tc_dev, h->policy_tc_host, sizeof(h->policy_tc_host), cudaMemcpyHostToDevice,
                                   (cudaStream_t)stream), "cudaMemcpyAsync(tensor-core policy)");
  }
  h->policy_set = true;
  return MSORT_OK;
}

extern "C" int msort_debug_policy_logits(msort_t* h, const float* sort_obs, int64_t count, float* logits, void* stream) {
  if (!h || !sort_obs || !logits || count < 0) return fail(MSORT_E_INVALID, "msort_debug_policy_logits: bad argument");
  if (!h->policy_set || !h->policy_tc_ok || !h->policy_tc_dev)
    return fail(MSORT_E_UNSUPPORTED, "msort_debug_policy_logits: no tensor-core policy (msort_set_policy not called, or the weights do not fit the fp16 split)");
  MSORT_TRY_CUDA(launch_tc_logits(sort_obs, h->policy_tc_dev, count, logits, h->sm_count, (cudaStream_t)stream), "tc_logits kernel");
  h->launches += 1;
  return MSORT_OK;
}

static int observe_impl(msort_t* h, const void* state, float* obs, uint8_t* mask, int after_shift, void* stream, const char* who) {
  if (!h || !state) return fail(MSORT_E_INVALID, "%s: NULL handle/state", who);
  if (!aligned(state, 16) || (obs && !aligned(obs, 16)) || (mask && !aligned(mask, 16)))
    return fail(MSORT_E_INVALID, "%s: state / obs / mask must be 16-byte aligned", who);
  MSORT_TRY_CUDA(launch_observe(h->dev, state, obs, mask, after_shift, (cudaStream_t)stream), "observe kernel");
  h->launches += 1;
  return MSORT_OK;
}

extern "C" int msort_observe(msort_t* h, const void* state, float* obs, uint8_t* mask, void* stream) {
  return observe_impl(h, state, obs, mask, 0, stream, "msort_observe");
}

extern "C" int msort_observe_after_shift(msort_t* h, const void* state, float* obs, uint8_t* mask, void* stream) {
  return observe_impl(h, state, obs, mask, 1, stream, "msort_observe_after_shift");
}

extern "C" int msort_sample_actions(msort_t* h, const uint8_t* mask, int64_t* actions, uint64_t seed, uint32_t t,
                                    void* stream) {
  if (!h || !mask || !actions) return fail(MSORT_E_INVALID, "msort_sample_actions: NULL argument");
  if (!aligned(mask, 4) || !aligned(actions, 8)) return fail(MSORT_E_INVALID, "msort_sample_actions: misaligned buffer");
  MSORT_TRY_CUDA(launch_sample(h->dev, mask, actions, seed, t, (cudaStream_t)stream), "sample kernel");
  h->launches += 1;
  return MSORT_OK;
}

extern "C" int msort_generate_streams(msort_t* h, uint32_t episode, uint32_t first_step, uint32_t num_steps, uint32_t* input_counts,
                                      double* noise_u, uint32_t* draw_words, uint8_t* first_pattern, void* stream) {
  if (!h) return fail(MSORT_E_INVALID, "msort_generate_streams: NULL handle");
  if ((input_counts && !aligned(input_counts, 4)) || (noise_u && !aligned(noise_u, 16)) || (draw_words && !aligned(draw_words, 16)))
    return fail(MSORT_E_INVALID, "msort_generate_streams: misaligned buffer");
  MSORT_TRY_CUDA(launch_generate_streams(h->dev, episode, first_step, num_steps, input_counts, noise_u, draw_words, first_pattern,
                                         (cudaStream_t)stream), "generate_streams kernel");
  h->launches += 1;
  return MSORT_OK;
}

extern "C" int msort_rule_based_actions(msort_t* h, const void* state, int after_shift, int64_t* actions, void* stream) {
  if (!h || !state || !actions) return fail(MSORT_E_INVALID, "msort_rule_based_actions: NULL argument");
  if (!aligned(state, 16) || !aligned(actions, 8)) return fail(MSORT_E_INVALID, "msort_rule_based_actions: misaligned buffer");
  MSORT_TRY_CUDA(launch_rule_actions(h->dev, state, after_shift, actions, (cudaStream_t)stream), "rule-based action kernel");
  h->launches += 1;
  return MSORT_OK;
}

static int policy_act_impl(msort_t* h, long long first, long long count, const float* obs, const uint8_t* mask,
                           const float* packed_weights, uint64_t seed, uint32_t t, int deterministic, int64_t* actions,
                           float* logp, float* value, void* stream) {
  if (!h || !obs || !mask || !packed_weights || !actions || !logp || !value)
    return fail(MSORT_E_INVALID, "msort_policy_act: NULL argument");
  if (!aligned(obs, 4) || !aligned(packed_weights, 16) || !aligned(actions, 8) || !aligned(logp, 4) || !aligned(value, 4))
    return fail(MSORT_E_INVALID, "msort_policy_act: misaligned buffer");
  const int D = msort_obs_dim(h), A = msort_num_actions(h);
  DevConfig d = h->dev;
  d.n = count; d.gid0 += first;                     // the draw is keyed by the global env id: ranges compose to the whole
  MSORT_TRY_CUDA(launch_policy_act(d, obs + first * D, mask + first * A, packed_weights, D, A, seed, t, h->draw_counter, deterministic,
                                   actions + first, logp + first, value + first, h->sm_count, (cudaStream_t)stream),
                 "policy_act kernel");
  h->launches += 1;
  return MSORT_OK;
}

extern "C" int msort_policy_act(msort_t* h, const float* obs, const uint8_t* mask, const float* packed_weights,
                                uint64_t seed, uint32_t t, int deterministic, int64_t* actions, float* logp,
                                float* value, void* stream) {
  if (!h) return fail(MSORT_E_INVALID, "msort_policy_act: NULL argument");
  return policy_act_impl(h, 0, h->dev.n, obs, mask, packed_weights, seed, t, deterministic, actions, logp, value, stream);
}

extern "C" int msort_policy_act_range(msort_t* h, int64_t first_env, int64_t num_envs, const float* obs, const uint8_t* mask,
                                      const float* packed_weights, uint64_t seed, uint32_t t, int deterministic,
                                      int64_t* actions, float* logp, float* value, void* stream) {
  if (!h) return fail(MSORT_E_INVALID, "msort_policy_act_range: NULL handle");
  if (first_env < 0 || num_envs <= 0 || first_env + num_envs > h->dev.n || first_env % kTile != 0)
    return fail(MSORT_E_INVALID, "msort_policy_act_range: range must lie inside the batch and start on a multiple of %d envs", kTile);
  return policy_act_impl(h, first_env, num_envs, obs, mask, packed_weights, seed, t, deterministic, actions, logp, value, stream);
}

extern "C" int msort_policy_eval(msort_t* h, int obs_dim, int num_actions, int64_t num_rows, const float* obs, int64_t obs_row_stride,
                                 const uint8_t* mask, int64_t mask_row_stride, const float* packed_weights, uint64_t seed, uint32_t t,
                                 int deterministic, int64_t* actions, float* logp, float* value, void* stream) {
  if (!h || !obs || !packed_weights || !actions || !logp || !value) return fail(MSORT_E_INVALID, "msort_policy_eval: NULL argument");
  if (!((obs_dim == 29 && num_actions == 22) || (obs_dim == 16 && num_actions == 11) || (obs_dim == 13 && num_actions == 2)))
    return fail(MSORT_E_UNSUPPORTED, "msort_policy_eval: (obs_dim, num_actions) must be (13,2), (16,11) or (29,22)");
  if (num_rows < 0 || obs_row_stride < obs_dim || (mask && mask_row_stride < num_actions))
    return fail(MSORT_E_INVALID, "msort_policy_eval: bad row count / stride");
  if (!aligned(obs, 4) || !aligned(packed_weights, 16) || !aligned(actions, 8) || !aligned(logp, 4) || !aligned(value, 4))
    return fail(MSORT_E_INVALID, "msort_policy_eval: misaligned buffer");
  MSORT_TRY_CUDA(launch_policy_eval(h->dev, obs_dim, num_actions, num_rows, obs, obs_row_stride, mask, mask_row_stride, packed_weights,
                                    seed, t, deterministic, actions, logp, value, h->sm_count, (cudaStream_t)stream),
                 "policy_eval kernel");
  h->launches += 1;
  return MSORT_OK;
}

extern "C" int msort_export_state(msort_t* h, const void* state, msort_env_state_t* out, void* stream) {
  if (!h || !state || !out) return fail(MSORT_E_INVALID, "msort_export_state: NULL argument");
  if (!aligned(state, 16) || !aligned(out, 8)) return fail(MSORT_E_INVALID, "msort_export_state: misaligned buffer");
  MSORT_TRY_CUDA(launch_export(h->dev, state, out, (cudaStream_t)stream), "export kernel");
  h->launches += 1;
  return MSORT_OK;
}

extern "C" int msort_gather_state(msort_t* h, const void* state, const int64_t* env_ids, int64_t count,
                                  msort_env_state_t* out, void* stream) {
  if (!h || !state || (count > 0 && (!env_ids || !out))) return fail(MSORT_E_INVALID, "msort_gather_state: NULL argument");
  if (count < 0) return fail(MSORT_E_INVALID, "msort_gather_state: negative count");
  if (!aligned(state, 16) || !aligned(out, 8) || !aligned(env_ids, 8)) return fail(MSORT_E_INVALID, "msort_gather_state: misaligned buffer");
  MSORT_TRY_CUDA(launch_gather(h->dev, state, env_ids, count, out, (cudaStream_t)stream), "gather kernel");
  h->launches += 1;
  return MSORT_OK;
}

extern "C" int msort_import_state(msort_t* h, void* state, const msort_env_state_t* in, void* stream) {
  if (!h || !state || !in) return fail(MSORT_E_INVALID, "msort_import_state: NULL argument");
  if (!aligned(state, 16) || !aligned(in, 8)) return fail(MSORT_E_INVALID, "msort_import_state: misaligned buffer");
  MSORT_TRY_CUDA(launch_import(h->dev, state, in, (cudaStream_t)stream), "import kernel");
  h->imported = true;
  h->launches += 1;
  return MSORT_OK;
}

extern "C" int msort_reduce_stats(msort_t* h, const void* state, double* out16, void* stream) {
  if (!h || !state || !out16) return fail(MSORT_E_INVALID, "msort_reduce_stats: NULL argument");
  if (!aligned(state, 16) || !aligned(out16, 8)) return fail(MSORT_E_INVALID, "msort_reduce_stats: misaligned buffer");
  MSORT_TRY_CUDA(launch_stats(h->dev, state, out16, h->sm_count, (cudaStream_t)stream), "stats kernel");
  h->launches += 1;
  return MSORT_OK;
}

extern "C" int msort_sync_check(msort_t* h, void* stream) {
  if (!h) return fail(MSORT_E_INVALID, "msort_sync_check: NULL handle");
  MSORT_TRY_CUDA(cudaStreamSynchronize((cudaStream_t)stream), "cudaStreamSynchronize");
  MSORT_TRY_CUDA(cudaGetLastError(), "sticky error");
  return MSORT_OK;
}

// ---------------------------------------------------------------- MaskablePPO update kernels (msort_ppo.cu)
static int ppo_check_batch(const msort_ppo_batch_t* b, bool update, const char* who) {
  if (!b) return fail(MSORT_E_INVALID, "%s: NULL batch", who);
  if (b->struct_size != sizeof(msort_ppo_batch_t)) return fail(MSORT_E_INVALID, "%s: bad batch struct_size", who);
  if (ppo_param_count(b->obs_dim, b->num_actions) <= 0 ||
      !((b->obs_dim == 29 && b->num_actions == 22) || (b->obs_dim == 16 && b->num_actions == 11) || (b->obs_dim == 13 && b->num_actions == 2)))
    return fail(MSORT_E_UNSUPPORTED, "%s: (obs_dim, num_actions) must be (13,2), (16,11) or (29,22)", who);
  if (b->num_rows <= 0 || !b->obs || !b->mask || !b->actions) return fail(MSORT_E_INVALID, "%s: NULL / empty batch buffers", who);
  if (update && (!b->old_logp || !b->adv || !b->ret)) return fail(MSORT_E_INVALID, "%s: old_logp / adv / ret are required", who);
  return MSORT_OK;
}

extern "C" int msort_ppo_param_count(int obs_dim, int num_actions) { return ppo_param_count(obs_dim, num_actions); }
extern "C" int msort_ppo_scratch_floats(int obs_dim, int num_actions) { return ppo_scratch_floats(obs_dim, num_actions); }

extern "C" int msort_ppo_forward(const msort_ppo_batch_t* batch, const float* params, float* logp_out, float* value_out, void* stream) {
  int rc = ppo_check_batch(batch, false, "msort_ppo_forward");
  if (rc != MSORT_OK) return rc;
  if (!params) return fail(MSORT_E_INVALID, "msort_ppo_forward: NULL params");
  MSORT_TRY_CUDA(ppo_forward(*batch, params, logp_out, value_out, (cudaStream_t)stream), "ppo forward kernel");
  return MSORT_OK;
}

extern "C" int msort_ppo_gae(int32_t T, int64_t n, const float* rew, const float* val, const uint8_t* done, const float* last_val,
                             float gamma, float gae_lambda, float* adv, float* ret, void* stream) {
  if (T <= 0 || n <= 0 || !rew || !val || !done || !last_val || !adv || !ret) return fail(MSORT_E_INVALID, "msort_ppo_gae: bad argument");
  MSORT_TRY_CUDA(ppo_gae(T, n, rew, val, done, last_val, gamma, gae_lambda, adv, ret, (cudaStream_t)stream), "gae kernel");
  return MSORT_OK;
}

extern "C" int msort_ppo_gradient(const msort_ppo_batch_t* batch, const msort_ppo_hparams_t* hp, const float* params, float* grads,
                                  const int64_t* idx, int64_t first, int64_t count, float* scratch, float* stats, void* stream) {
  int rc = ppo_check_batch(batch, true, "msort_ppo_gradient");
  if (rc != MSORT_OK) return rc;
  if (!hp || hp->struct_size != sizeof(msort_ppo_hparams_t)) return fail(MSORT_E_INVALID, "msort_ppo_gradient: bad hparams");
  if (!params || !grads || !scratch || count <= 0 || first < 0 || first + count > batch->num_rows)
    return fail(MSORT_E_INVALID, "msort_ppo_gradient: bad argument");
  if (!aligned(scratch, 16)) return fail(MSORT_E_INVALID, "msort_ppo_gradient: scratch must be 16-byte aligned");
  MSORT_TRY_CUDA(ppo_gradient(*batch, *hp, params, grads, idx, first, count, scratch, stats, (cudaStream_t)stream), "ppo gradient kernel");
  return MSORT_OK;
}

extern "C" int msort_ppo_update(const msort_ppo_batch_t* batch, const msort_ppo_hparams_t* hp, float* params, float* grads, float* adam_m,
                                float* adam_v, int32_t* step, const int64_t* perms, int32_t n_epochs, int64_t batch_size, float* scratch,
                                float* stats, void* stream) {
  int rc = ppo_check_batch(batch, true, "msort_ppo_update");
  if (rc != MSORT_OK) return rc;
  if (!hp || hp->struct_size != sizeof(msort_ppo_hparams_t)) return fail(MSORT_E_INVALID, "msort_ppo_update: bad hparams");
  if (!params || !grads || !adam_m || !adam_v || !step || !perms || !scratch || n_epochs <= 0 || batch_size <= 0)
    return fail(MSORT_E_INVALID, "msort_ppo_update: bad argument");
  if (!aligned(scratch, 16)) return fail(MSORT_E_INVALID, "msort_ppo_update: scratch must be 16-byte aligned");
  const int P = ppo_param_count(batch->obs_dim, batch->num_actions);
  const long long N = batch->num_rows;
  for (int e = 0; e < n_epochs; ++e)
    for (long long s = 0; s < N; s += batch_size) {
      const long long cnt = std::min<long long>(batch_size, N - s);
      MSORT_TRY_CUDA(ppo_gradient(*batch, *hp, params, grads, perms + (long long)e * N, s, cnt, scratch, stats, (cudaStream_t)stream),
                     "ppo gradient kernel");
      MSORT_TRY_CUDA(ppo_adam(params, grads, adam_m, adam_v, step, P, *hp, (cudaStream_t)stream), "adam kernel");
    }
  return MSORT_OK;
}
