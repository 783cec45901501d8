// libmbe.so -- C ABI (include/mbe.h) over the sm_100a kernels in mbe_step.cuh.
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>
#include <vector>

#include "../../include/mbe.h"
#include "mbe_step.cuh"
#include "mbe_step_spec.cuh"
#include "mbe_step_big.cuh"
#include "mbe_step_upt.cuh"
#include "mbe_step_tpe.cuh"
#include "mbe_host_wire.cuh"

#include <memory>
#include <sched.h>

namespace {

thread_local std::string g_err;

int fail(const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  g_err = buf;
  return 1;
}

#define MBE_CUDA(expr)                                                                     \
  do {                                                                                     \
    cudaError_t e_ = (expr);                                                               \
    if (e_ != cudaSuccess) return fail("%s failed: %s", #expr, cudaGetErrorString(e_));    \
  } while (0)

// Makes the handle's device current for the duration of a call and restores the caller's device
// afterwards (a process may drive several devices / handles from one thread).
struct DeviceGuard {
  int prev = -1;
  bool switched = false;
  cudaError_t err = cudaSuccess;
  explicit DeviceGuard(int device) {
    err = cudaGetDevice(&prev);
    if (err == cudaSuccess && prev != device) {
      err = cudaSetDevice(device);
      switched = err == cudaSuccess;
    }
  }
  ~DeviceGuard() {
    if (switched) cudaSetDevice(prev);
  }
};

#define MBE_ON_DEVICE(dev)                                                                          \
  DeviceGuard guard_(dev);                                                                          \
  if (guard_.err != cudaSuccess) return fail("cudaSetDevice(%d): %s", dev, cudaGetErrorString(guard_.err))

}  // namespace

struct mbe_env {
  mbe_config cfg;
  mbe_buffers bufs;
  bool bound = false;
  mbe::StepArgs args;
  std::vector<void*> luts;  // device tables owned by the handle
  uint8_t* d_bs_class = nullptr;
  uint8_t* d_ue_class = nullptr;
  size_t smem = 0;
  int grid = 0;
  int64_t launches = 0;
  // specialised fused kernel for this shape (nullptr: generic kernel only)
  void (*spec)(mbe::StepArgs) = nullptr;
  size_t spec_smem = 0;
  void (*pipe)(mbe::StepArgs) = nullptr;
  size_t pipe_smem = 0;
  int pipe_grid = 0;
  bool big = false;  // block-per-env kernel (wide shapes, ProportionalFair)
  // several-UEs-per-thread kernel (GYM scenario shapes, shared layout)
  void (*upt)(mbe::StepArgs) = nullptr;
  size_t upt_smem = 0;
  int upt_epb = 0;
  int upt_pf = 0, spec_pf = 0;  // L2 prefetch distance (CTAs) of the two kernel families, 0 = off
  bool pf_bound_ok = false;     // the prefetched streams are 16-byte aligned
  // thread-per-env FORK kernel (per-env layouts, E % 32 == 0, 16-byte aligned buffers)
  void (*tpe)(mbe::StepArgs) = nullptr;
  size_t tpe_smem = 0;
  bool tpe_bound_ok = false;
  void (*big_fn)(mbe::StepArgs) = nullptr;       // block-per-env kernel instance (UEs per thread by U)
  size_t big_smem = 0;
  bool tpe_shared = false;                       // the fused episode of a shared layout (no per-env BS table)
  void (*tpe_rollout)(mbe::StepArgs) = nullptr;  // the same kernel looping over the steps of an episode
  size_t tpe_rollout_smem = 0;
  // mbe_step_host pipeline: second stream + fork/join events (created on first use)
  cudaStream_t host_stream = nullptr;
  cudaEvent_t host_fork = nullptr, host_join = nullptr;
  // compact observation wire format of mbe_step_host (mbe_host_wire.cuh; created on first use)
  mbe::WireShape wire_shape = {};
  unsigned char* wire_dev = nullptr;     // [E * bytes_per_env]
  unsigned char* wire_pinned = nullptr;  // the same, pinned host staging
  std::vector<cudaEvent_t> wire_events;  // one per env window
  std::unique_ptr<mbe::ExpandCrew> wire_pool;
};

namespace {

// Programmatic dependent launch (the next step's CTAs become resident while this grid drains and
// prefetch their state slices into L2 before the dependency wait).  Measured on B200
// (profiles/README.md): 3-7% (graph replay) to 10% (plain launches) faster for the medium GYM
// kernels, 14% for the small shape, neutral for the large one, 2% slower for the FORK kernels (no
// prefetch there) -> default on for the GYM step kernels only; MBE_PDL=0 turns it off everywhere,
// MBE_PDL=1 on everywhere.
int pdl_mode() {
  static const int mode = []() {
    const char* v = std::getenv("MBE_PDL");
    return !v ? -1 : (v[0] == '1' ? 1 : 0);
  }();
  return mode;
}
bool pdl_enabled() { return pdl_mode() == 1; }
bool pdl_enabled_gym() { return pdl_mode() != 0; }

struct SpecEntry {
  int mode, handler, U, B, per_env;
  void (*fn)(mbe::StepArgs);
  size_t smem;
  void (*pipe_fn)(mbe::StepArgs);  // persistent, TMA-pipelined variant (E % EPB == 0)
  size_t pipe_smem;
  void (*mc_fn)(mbe::StepArgs);    // several BS classes in a shared layout (nullptr for per-env layouts)
};

#define MBE_SPEC(MODE, HANDLER, U, B, PE)                                             \
  SpecEntry {                                                                         \
    MODE, HANDLER, U, B, PE, mbe::step_spec_kernel<MODE, HANDLER, U, B, (PE != 0)>,   \
        mbe::spec_smem_bytes<HANDLER, U, B, (PE != 0)>(MODE == 1),                    \
        mbe::step_pipe_kernel<MODE, HANDLER, U, B, (PE != 0)>,                        \
        mbe::pipe_smem_bytes<HANDLER, U, B, (PE != 0)>(MODE == 1),                    \
        mbe::spec_mc_kernel<MODE, HANDLER, U, B, (PE != 0)>()                         \
  }

// shapes with a compile-time specialisation: the scenario sizes of BASELINE.json (small 3x5,
// medium 4x15, large 13x30) and the fork's MComCustom (7 UEs, <= 10 random BSs per env)
struct UptEntry {
  int handler, U, B, K;
  void (*fn)(mbe::StepArgs);
  size_t smem;
};
#define MBE_UPT(HANDLER, U, B, K) \
  UptEntry { HANDLER, U, B, K, mbe::step_upt_kernel<HANDLER, U, B, K>, mbe::upt_smem_bytes<HANDLER, U, B, K>() }
#ifndef MBE_UPT_K_MEDIUM
#define MBE_UPT_K_MEDIUM 5
#endif
// (small 5x3 with K=1 and large 30x13 with K=10 were measured slower than one thread per UE)
const UptEntry kUpts[] = {MBE_UPT(0, 15, 4, MBE_UPT_K_MEDIUM), MBE_UPT(1, 15, 4, MBE_UPT_K_MEDIUM)};

const SpecEntry kSpecs[] = {
    MBE_SPEC(1, 0, 5, 3, 0),  MBE_SPEC(1, 1, 5, 3, 0),  MBE_SPEC(1, 0, 15, 4, 0),  MBE_SPEC(1, 1, 15, 4, 0),
    MBE_SPEC(1, 0, 30, 13, 0), MBE_SPEC(1, 1, 30, 13, 0), MBE_SPEC(0, 0, 7, 10, 1), MBE_SPEC(0, 0, 5, 3, 0),
    MBE_SPEC(0, 0, 15, 4, 0), MBE_SPEC(0, 0, 30, 13, 0), MBE_SPEC(1, 0, 7, 10, 1),  MBE_SPEC(1, 1, 7, 10, 1),
};

}  // namespace

extern "C" {

int mbe_abi_version(void) { return MBE_ABI_VERSION; }

const char* mbe_build_info(void) {
#define MBE_STR2(x) #x
#define MBE_STR(x) MBE_STR2(x)
  return "libmbe abi " MBE_STR(MBE_ABI_VERSION) " | sm_100a | built " __DATE__ " " __TIME__;
}

int mbe_struct_size(int which) {
  switch (which) {
    case 0: return (int)sizeof(mbe_config);
    case 1: return (int)sizeof(mbe_buffers);
    case 2: return (int)sizeof(mbe_link_class);
    case 3: return (int)sizeof(mbe_ue_class);
    case 4: return (int)sizeof(mbe_rollout_out);
    default: return -1;
  }
}

const char* mbe_last_error(void) { return g_err.c_str(); }

int mbe_create(const mbe_config* cfg, mbe_env** out) {
  if (!cfg || !out) return fail("mbe_create: null argument");
  *out = nullptr;
  if (cfg->abi_version != MBE_ABI_VERSION)
    return fail("mbe_create: abi_version %d, library is %d", cfg->abi_version, MBE_ABI_VERSION);
  if (cfg->num_envs <= 0) return fail("mbe_create: num_envs must be > 0");
  if (cfg->num_ues <= 0 || cfg->num_ues > mbe::kBigMaxI * mbe::kBigThreads)
    return fail("mbe_create: num_ues=%d out of range (1..%d)", cfg->num_ues, mbe::kBigMaxI * mbe::kBigThreads);
  if (cfg->num_bs <= 0 || cfg->num_bs > mbe::kBigMaxB)
    return fail("mbe_create: num_bs=%d out of range (1..%d)", cfg->num_bs, mbe::kBigMaxB);
  if (cfg->mode != MBE_MODE_FORK && cfg->mode != MBE_MODE_GYM) return fail("mbe_create: bad mode %d", cfg->mode);
  if (cfg->handler != MBE_HANDLER_CENTRAL && cfg->handler != MBE_HANDLER_MA)
    return fail("mbe_create: bad handler %d", cfg->handler);
  if (cfg->scheduler != MBE_SCHED_RESOURCE_FAIR && cfg->scheduler != MBE_SCHED_PROPORTIONAL_FAIR &&
      cfg->scheduler != MBE_SCHED_RATE_FAIR)
    return fail("mbe_create: scheduler %d not available", cfg->scheduler);
  const int nuc = cfg->num_ue_classes > 1 ? cfg->num_ue_classes : 1;
  if (nuc > MBE_MAX_UE_CLASSES) return fail("mbe_create: num_ue_classes=%d out of range", cfg->num_ue_classes);
  if (nuc > 1 && !cfg->ue_class) return fail("mbe_create: num_ue_classes > 1 needs ue_class");
  if (cfg->num_classes < 1 || cfg->num_classes * nuc > MBE_MAX_CLASSES)
    return fail("mbe_create: %d BS classes x %d UE classes exceed %d link classes", cfg->num_classes, nuc,
                MBE_MAX_CLASSES);
  const int nlink = cfg->num_classes * nuc;
  if (!(cfg->width > 0 && cfg->width <= 32767 && cfg->height > 0 && cfg->height <= 32767))
    return fail("mbe_create: map %gx%g does not fit int16 coordinates", cfg->width, cfg->height);
  if (cfg->ep_time <= 0) return fail("mbe_create: ep_time must be > 0");
  if (cfg->bs_random_max > 0 &&
      (cfg->bs_layout != MBE_BS_PER_ENV || cfg->bs_random_min < 1 || cfg->bs_random_max > cfg->num_bs ||
       cfg->bs_random_min > cfg->bs_random_max))
    return fail("mbe_create: random BS layouts need bs_layout=PER_ENV and 1 <= min <= max <= num_bs");
  if (!(cfg->util_upper > cfg->util_lower) || !(cfg->util_w3 > 0) || cfg->util_w3 == 1.0)
    return fail("mbe_create: bad utility parameters");
  for (int c = 0; c < nlink; ++c) {
    if (cfg->classes[c].d2max >= 0 && !cfg->classes[c].rate_lut)
      return fail("mbe_create: link class %d has no rate_lut", c);
    if (cfg->classes[c].log2snr_lut && cfg->classes[c].log2snr_len <= 0)
      return fail("mbe_create: link class %d: log2snr_lut needs log2snr_len > 0", c);
  }
  if (nuc > 1)
    for (int u = 0; u < cfg->num_ues; ++u)
      if (cfg->ue_class[u] >= nuc) return fail("mbe_create: ue_class[%d]=%d >= num_ue_classes", u, cfg->ue_class[u]);

  if (cfg->env_offset < 0 || cfg->env_offset + (int64_t)cfg->num_envs > (int64_t)0xffffffffll)
    return fail("mbe_create: env_offset + num_envs must stay below 2^32 (global env ids are 32-bit Philox counter words)");
  MBE_ON_DEVICE(cfg->device);
  mbe_env* env = new (std::nothrow) mbe_env();
  if (!env) return fail("mbe_create: out of host memory");
  env->cfg = *cfg;
  std::memset(&env->bufs, 0, sizeof env->bufs);
  mbe::StepArgs& a = env->args;
  std::memset(&a, 0, sizeof a);
  a.E = cfg->num_envs;
  a.U = cfg->num_ues;
  a.B = cfg->num_bs;
  const bool gym = cfg->mode == MBE_MODE_GYM, ma = cfg->handler == MBE_HANDLER_MA;
  env->big = a.U > 32 || a.B > 32 || cfg->scheduler != MBE_SCHED_RESOURCE_FAIR;
  a.scheduler = cfg->scheduler;
  a.F = gym ? (ma ? 4 * a.B + 1 : 2 * a.B + 1) : 0;
  a.epw = env->big ? 1 : 32 / a.U;
  a.epb = a.epw * mbe::kWarpsPerBlock;
  a.env_offset = (unsigned)cfg->env_offset;
  a.traj_gid_mask = (cfg->flags & MBE_FLAG_SHARED_TRAJECTORY) ? 0u : 0xffffffffu;
  a.ep_time = cfg->ep_time;
  a.autoreset = cfg->autoreset;
  a.reset_rng_episode = cfg->reset_rng_episode;
  a.bs_per_env = cfg->bs_layout == MBE_BS_PER_ENV;
  a.bs_rand_min = cfg->bs_random_min;
  a.bs_rand_max = cfg->bs_random_max;
  a.seed_lo = (unsigned)(cfg->seed & 0xffffffffu);
  a.seed_hi = (unsigned)(cfg->seed >> 32);
  a.width = cfg->width;
  a.height = cfg->height;
  a.wh_int = (cfg->width == std::floor(cfg->width) && cfg->height == std::floor(cfg->height)) ? 1 : 0;
  a.wh_int_w = (int)cfg->width;
  a.wh_int_h = (int)cfg->height;
  for (int c = 0; c < nuc; ++c) {
    const double vel = nuc > 1 ? cfg->ue_classes[c].velocity : cfg->velocity;
    mbe::MoveDev& m = a.mv[c];
    m.velocity = vel;
    m.velocity_f = (float)vel;
    m.tie_eps = (float)(vel * 1e-6 + 1e-6);
    m.move_d2max = nuc > 1 ? cfg->ue_classes[c].move_d2max : cfg->move_d2max;
    // is an axis-aligned step exactly +-velocity in the reference's FP64 chain (movement.py:58-59)?
    volatile double v = vel;
    int ok = 1;
    const int dmax = (int)std::max(cfg->width, cfg->height) + 1;
    for (int d = 1; d <= dmax && ok; ++d) {
      volatile double prod = v * (double)d;
      volatile double q = prod / (double)d;
      if (q != v) ok = 0;
    }
    m.axis_exact = ok;
  }
  a.util_c = (float)(cfg->util_w1 * std::log(2.0) / std::log(cfg->util_w3));
  a.util_w2 = (float)cfg->util_w2;
  a.util_lo = (float)cfg->util_lower;
  a.util_hi = (float)cfg->util_upper;
  a.util_scale = (float)(2.0 / (cfg->util_upper - cfg->util_lower));
  a.inv_U = 1.0f / (float)a.U;
  a.n_classes = cfg->num_classes;
  a.n_ue_classes = nuc;
  bool any_ltab = false;
  for (int c = 0; c < nlink; ++c) {
    const mbe_link_class& h = cfg->classes[c];
    mbe::ClassDev& d = a.cls[c];
    d.l0_hi = (float)h.l0;
    d.l0_lo = (float)(h.l0 - (double)d.l0_hi);
    d.k_hi = (float)h.k;
    d.k_lo = (float)(h.k - (double)d.k_hi);
    d.l_zero = (float)h.l_zero;
    d.d2max = h.d2max;
    d.stride = h.d2max + 1;
    d.lutn = nullptr;
    d.lut0 = nullptr;
    d.ltab = nullptr;
    d.ltab_len = 0;
    if (h.log2snr_lut) {  // log2 snr per d2 (loss not affine in log-distance)
      float* pt = nullptr;
      size_t bytes = (size_t)h.log2snr_len * sizeof(float);
      cudaError_t e0 = cudaMalloc(&pt, bytes);
      if (e0 == cudaSuccess) e0 = cudaMemcpy(pt, h.log2snr_lut, bytes, cudaMemcpyHostToDevice);
      if (e0 != cudaSuccess) {
        mbe_destroy(env);
        return fail("mbe_create: log2snr_lut upload failed: %s", cudaGetErrorString(e0));
      }
      env->luts.push_back(pt);
      d.ltab = pt;
      d.ltab_len = h.log2snr_len;
      any_ltab = true;
    }
    if (h.d2max >= 0) {  // the raw table, for the block-per-env kernel
      double* p0 = nullptr;
      size_t bytes0 = (size_t)d.stride * sizeof(double);
      cudaError_t e0 = cudaMalloc(&p0, bytes0);
      if (e0 == cudaSuccess) e0 = cudaMemcpy(p0, h.rate_lut, bytes0, cudaMemcpyHostToDevice);
      if (e0 != cudaSuccess) {
        mbe_destroy(env);
        return fail("mbe_create: rate_lut upload failed: %s", cudaGetErrorString(e0));
      }
      env->luts.push_back(p0);
      d.lut0 = p0;
    }
    if (h.d2max >= 0 && !env->big) {
      // device table: [0] = 0.0, then rows n = 1..U of round(rate_lut[d2] / n, 2) in FP64
      // (schedules.py:20-22, base.py:435); kernels index it as lutn[n*stride + d2]
      const int rows = a.U;
      std::vector<double> tab((size_t)rows * d.stride + 1);
      tab[0] = 0.0;
      for (int n = 1; n <= rows; ++n)
        for (int i = 0; i < d.stride; ++i) {
          volatile double share = h.rate_lut[i] / (double)n;
          volatile double scaled = share * 100.0;
          tab[1 + (size_t)(n - 1) * d.stride + i] = std::nearbyint(scaled) / 100.0;
        }
      double* p = nullptr;
      size_t bytes = tab.size() * sizeof(double);
      cudaError_t e = cudaMalloc(&p, bytes);
      if (e == cudaSuccess) e = cudaMemcpy(p, tab.data(), bytes, cudaMemcpyHostToDevice);
      if (e != cudaSuccess) {
        mbe_destroy(env);
        return fail("mbe_create: rate table upload failed: %s", cudaGetErrorString(e));
      }
      env->luts.push_back(p);
      d.lutn = p + 1 - d.stride;  // biased: row n starts at lutn + n*stride, lutn[stride-1] == 0.0
    }
  }
  if (cfg->bs_class) {
    for (int b = 0; b < cfg->num_bs; ++b)
      if (cfg->bs_class[b] >= cfg->num_classes) {
        mbe_destroy(env);
        return fail("mbe_create: bs_class[%d]=%d >= num_classes", b, cfg->bs_class[b]);
      }
    cudaError_t e = cudaMalloc(&env->d_bs_class, cfg->num_bs);
    if (e == cudaSuccess) e = cudaMemcpy(env->d_bs_class, cfg->bs_class, cfg->num_bs, cudaMemcpyHostToDevice);
    if (e != cudaSuccess) {
      mbe_destroy(env);
      return fail("mbe_create: bs_class upload failed: %s", cudaGetErrorString(e));
    }
    a.bs_class = env->d_bs_class;
  }
  if (nuc > 1) {
    cudaError_t e = cudaMalloc(&env->d_ue_class, cfg->num_ues);
    if (e == cudaSuccess) e = cudaMemcpy(env->d_ue_class, cfg->ue_class, cfg->num_ues, cudaMemcpyHostToDevice);
    if (e != cudaSuccess) {
      mbe_destroy(env);
      return fail("mbe_create: ue_class upload failed: %s", cudaGetErrorString(e));
    }
    a.ue_class = env->d_ue_class;
  }
  for (int b = 0; b < mbe::kMaxSlots; ++b) {
    const int c = (cfg->bs_class && b < cfg->num_bs) ? cfg->bs_class[b] : 0;
    const mbe::ClassDev& d = a.cls[c * nuc];  // (the slot table serves the one-UE-class kernels)
    mbe::SlotDev& sl = a.slot[b];
    sl.x = sl.y = 0;  // coordinates of a shared layout arrive with mbe_bind (host copy of bs_xy)
    sl.d2max = d.d2max;
    sl.stride = d.stride;
    sl.k = d.k_hi;
    sl.l0 = d.l0_hi;
    sl.xf = sl.yf = 0.0f;
    sl.lutn = d.lutn;
  }
  // the specialised kernels assume one UE class, log2 snr affine in log-distance and squared
  // distances exact in FP32; the warp-segment ones take per-BS classes (folded per slot), the
  // UEs-per-thread and thread-per-env ones a single class
  const bool spec_ok = nuc == 1 && !any_ltab &&
                       std::floor(cfg->width) * std::floor(cfg->width) + std::floor(cfg->height) * std::floor(cfg->height) <
                           16777216.0;
  if (spec_ok && !(cfg->flags & MBE_FLAG_GENERIC_KERNEL)) {
    for (const SpecEntry& sp : kSpecs)
      if (sp.mode == cfg->mode && (sp.handler == cfg->handler || !gym) && sp.U == a.U && sp.B == a.B &&
          sp.per_env == a.bs_per_env) {
        if (cfg->num_classes > 1 && !sp.mc_fn) break;  // per-env layouts have one class: generic kernel
        env->spec = cfg->num_classes > 1 ? sp.mc_fn : sp.fn;
        env->spec_smem = sp.smem;
        if (a.E % a.epb == 0 && cfg->num_classes == 1) {
          env->pipe = sp.pipe_fn;
          env->pipe_smem = sp.pipe_smem;
        }
        break;
      }
  }
  {
    const char* v = std::getenv("MBE_UPT");
    const bool on = !(v && v[0] == '0');
    if (on && spec_ok && cfg->num_classes == 1 && gym && !a.bs_per_env && !env->big &&
        !(cfg->flags & MBE_FLAG_GENERIC_KERNEL))
      for (const UptEntry& t : kUpts)
        if (t.handler == cfg->handler && t.U == a.U && t.B == a.B) {
          env->upt = t.fn;
          env->upt_smem = t.smem;
          env->upt_epb = (32 / t.K) * MBE_UPT_WARPS;
        }
  }
  {
    const char* v = std::getenv("MBE_TPE");
    const bool on = !(v && v[0] == '0');
    if (on && !gym && a.bs_per_env && spec_ok && cfg->num_classes == 1 && !env->big && a.E % 32 == 0 && a.U == 7 && a.B == 10 &&
        cfg->width <= 2048 && cfg->height <= 2048 &&
        !(cfg->flags & MBE_FLAG_GENERIC_KERNEL)) {
      env->tpe = mbe::step_tpe_fork_kernel<7, 10>;
      env->tpe_smem = sizeof(mbe::TpeForkSmem<7, 10>);
      env->tpe_rollout = mbe::step_tpe_fork_kernel<7, 10, true>;
      env->tpe_rollout_smem = sizeof(mbe::TpeRolloutSmem<7, 10>);
    }
    // the scenario shapes in FORK mode (one layout shared by all envs): the fused episode only (5 x 3 and
    // 15 x 4); their single step stays on the warp-segment kernel
    if (on && !gym && !a.bs_per_env && spec_ok && cfg->num_classes == 1 && !env->big && a.E % 32 == 0 &&
        cfg->width <= 2048 && cfg->height <= 2048 && !(cfg->flags & MBE_FLAG_GENERIC_KERNEL)) {
      env->tpe_shared = true;
      if (a.U == 5 && a.B == 3) {
        env->tpe_rollout = mbe::step_tpe_fork_kernel<5, 3, true, true>;
        env->tpe_rollout_smem = sizeof(mbe::TpeRolloutSmem<5, 3>);
      } else if (a.U == 15 && a.B == 4) {
        env->tpe_rollout = mbe::step_tpe_fork_kernel<15, 4, true, true>;
        env->tpe_rollout_smem = sizeof(mbe::TpeRolloutSmem<15, 4>);
      } else if (a.U == 30 && a.B == 13 && std::getenv("MBE_TPE_LARGE") && std::getenv("MBE_TPE_LARGE")[0] == '1') {
        // opt-in: one env per lane walks 30 x 13 pairs in scalar code at 208 registers -- 1.43 ms per
        // 20-step episode of 65,536 envs against 0.83 ms for 20 launches of the warp-segment kernel
        // (profiles/r02_j_fork_rollout.txt); the small and medium shapes win 7.1x and 1.8x
        env->tpe_rollout = mbe::step_tpe_fork_kernel<30, 13, true, true>;
        env->tpe_rollout_smem = sizeof(mbe::TpeRolloutSmem<30, 13>);
      } else {
        env->tpe_shared = false;
      }
    }
  }
  if (env->big && std::floor(cfg->width) * std::floor(cfg->width) + std::floor(cfg->height) * std::floor(cfg->height) >=
                      16777216.0) {
    mbe_destroy(env);
    return fail("mbe_create: the block-per-env kernel (wide shapes, ProportionalFair / RateFair) needs a map whose "
                "squared diagonal is below 2^24 (distances are exact FP32 integers)");
  }
  env->smem = env->big ? 0 : mbe::smem_bytes(gym, ma, a.epb, a.U, a.B, a.F, a.bs_per_env);
  env->grid = env->big ? a.E : (a.E + a.epb - 1) / a.epb;
  if (env->big) env->spec = nullptr;
  if (env->smem > 200 * 1024) {
    mbe_destroy(env);
    return fail("mbe_create: needs %zu bytes of shared memory per block", env->smem);
  }
  const void* fn = gym ? (ma ? (const void*)mbe::step_kernel<1, 1> : (const void*)mbe::step_kernel<1, 0>)
                       : (const void*)mbe::step_kernel<0, 0>;
  cudaError_t e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)env->smem);
  if (e == cudaSuccess && env->big) {
    // UEs per thread held in registers: 2 (64 registers, 4 CTAs per SM) up to 512 UEs, else 4
// 1 = always the 4-UEs-per-thread instance (80 registers, 3 CTAs per SM).  With the observation rows
// leaving as aligned bulk tiles the 2-UEs-per-thread instance for U <= 512 (64 registers, 4 CTAs per SM,
// smaller shared-memory arrays) is the faster one: 802 vs 920 us per 16,384-env synthetic step
// (profiles/r02_zb_variants.txt); while the rows were direct 4-byte stores it was the slower one
// (1105 vs 1046 us, r02_q): more resident CTAs only meant more contention on misaligned sectors
#ifndef MBE_BIG_FORCE_MAXI4
#define MBE_BIG_FORCE_MAXI4 0
#endif
    const bool two = a.U <= 2 * mbe::kBigThreads && !MBE_BIG_FORCE_MAXI4;
    env->big_fn = gym ? (ma ? (two ? mbe::step_big_kernel<1, 1, 2> : mbe::step_big_kernel<1, 1, 4>)
                            : (two ? mbe::step_big_kernel<1, 0, 2> : mbe::step_big_kernel<1, 0, 4>))
                      : (two ? mbe::step_big_kernel<0, 0, 2> : mbe::step_big_kernel<0, 0, 4>);
    env->big_smem = two ? mbe::big_smem_bytes<2>() : mbe::big_smem_bytes<4>();
    e = cudaFuncSetAttribute((const void*)env->big_fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)env->big_smem);
  }
  if (e == cudaSuccess && env->spec)
    e = cudaFuncSetAttribute((const void*)env->spec, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)env->spec_smem);
  if (e == cudaSuccess && env->upt)
    e = cudaFuncSetAttribute((const void*)env->upt, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)env->upt_smem);
  if (e == cudaSuccess && env->tpe)
    e = cudaFuncSetAttribute((const void*)env->tpe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)env->tpe_smem);
  if (e == cudaSuccess && env->tpe_rollout)
    e = cudaFuncSetAttribute((const void*)env->tpe_rollout, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)env->tpe_rollout_smem);
  if (e == cudaSuccess && env->pipe) {
    // measured slower than the one-chunk-per-CTA kernel on B200 (profiles/README.md): opt-in only
    const char* v = std::getenv("MBE_PIPE");
    if (!(v && v[0] == '1')) env->pipe = nullptr;
  }
  if (e == cudaSuccess && env->pipe) {
    e = cudaFuncSetAttribute((const void*)env->pipe, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)env->pipe_smem);
    int per_sm = 0, sms = 0;
    if (e == cudaSuccess)
      e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, (const void*)env->pipe, mbe::kThreads,
                                                        env->pipe_smem);
    if (e == cudaSuccess) e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, cfg->device);
    // persistent grid: one full wave of resident CTAs, never more CTAs than chunks
    env->pipe_grid = std::max(1, std::min(a.E / a.epb, per_sm * sms));
  }
  if (e == cudaSuccess) {
    // L2 prefetch distance: 0.75 waves of resident CTAs (measured best of 0.25 .. 2 waves on B200,
    // profiles/README.md); MBE_PREFETCH=0 turns it off, MBE_PREFETCH=n sets the distance in CTAs
    const char* v = std::getenv("MBE_PREFETCH");
    const int forced = v ? std::atoi(v) : -1;
    int sms = 0;
    e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, cfg->device);
    auto dist = [&](const void* fn, int threads, size_t smem) {
      int per_sm = 0;
      if (e == cudaSuccess) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fn, threads, smem);
      return forced >= 0 ? forced : per_sm * sms * 3 / 4;
    };
    if (env->upt) env->upt_pf = dist((const void*)env->upt, 32 * MBE_UPT_WARPS, env->upt_smem);
    if (env->spec) env->spec_pf = dist((const void*)env->spec, mbe::kThreads, env->spec_smem);
  }
  if (e != cudaSuccess) {
    mbe_destroy(env);
    return fail("mbe_create: cudaFuncSetAttribute(%zu B smem): %s", env->smem, cudaGetErrorString(e));
  }
  *out = env;
  return 0;
}

void mbe_destroy(mbe_env* env) {
  if (!env) return;
  for (void* p : env->luts) cudaFree(p);
  if (env->d_bs_class) cudaFree(env->d_bs_class);
  if (env->d_ue_class) cudaFree(env->d_ue_class);
  if (env->host_stream) cudaStreamDestroy(env->host_stream);
  if (env->host_fork) cudaEventDestroy(env->host_fork);
  if (env->host_join) cudaEventDestroy(env->host_join);
  env->wire_pool.reset();
  for (cudaEvent_t e : env->wire_events) cudaEventDestroy(e);
  if (env->wire_dev) cudaFree(env->wire_dev);
  if (env->wire_pinned) cudaFreeHost(env->wire_pinned);
  delete env;
}

int mbe_bind(mbe_env* env, const mbe_buffers* b) {
  if (!env || !b) return fail("mbe_bind: null argument");
  MBE_ON_DEVICE(env->cfg.device);
  const bool gym = env->cfg.mode == MBE_MODE_GYM;
  if (!b->pos || !b->wp || !b->t || !b->episode || !b->bs_xy || !b->utility || !b->done)
    return fail("mbe_bind: pos, wp, t, episode, bs_xy, utility and done are required");
  if (gym && (!b->conn || !b->actions || !b->obs || !b->reward))
    return fail("mbe_bind: GYM mode needs conn, actions, obs and reward");
  if (!gym && !b->assoc) return fail("mbe_bind: FORK mode needs assoc");
  if (env->cfg.bs_random_max > 0 && !b->nbs) return fail("mbe_bind: random BS layouts need nbs");
  if (b->inj_wp && (!b->wp_cnt || b->inj_k <= 0)) return fail("mbe_bind: inj_wp needs wp_cnt and inj_k > 0");
  if (((uintptr_t)b->pos | (uintptr_t)b->wp | (uintptr_t)b->bs_xy | (uintptr_t)b->inj_wp) & 3)
    return fail("mbe_bind: int16 pair buffers must be 4-byte aligned");
  if (b->metrics && ((uintptr_t)b->metrics & 15)) return fail("mbe_bind: metrics must be 16-byte aligned");
  if (b->rate && ((uintptr_t)b->rate & 7)) return fail("mbe_bind: rate must be 8-byte aligned");
  env->bufs = *b;
  mbe::StepArgs& a = env->args;
  a.pos = reinterpret_cast<uint32_t*>(b->pos);
  a.wp = reinterpret_cast<uint32_t*>(b->wp);
  a.t = b->t;
  a.episode = b->episode;
  a.bs_xy = reinterpret_cast<uint32_t*>(b->bs_xy);
  a.nbs = b->nbs;
  a.conn = b->conn;
  a.assoc = b->assoc;
  a.actions = b->actions;
  a.rate = b->rate;
  a.utility = b->utility;
  a.obs = b->obs;
  a.reward = b->reward;
  a.done = b->done;
  a.metrics = b->metrics;
  a.dbg_snr = b->dbg_snr;
  a.inj_wp = reinterpret_cast<const uint32_t*>(b->inj_wp);
  a.wp_cnt = b->wp_cnt;
  a.inj_k = b->inj_k;
  a.obs_bulk_ok = b->obs && (((uintptr_t)b->obs & 15) == 0);
  {
    uintptr_t all = (uintptr_t)b->pos | (uintptr_t)b->wp | (uintptr_t)b->t | (uintptr_t)b->episode |
                    (uintptr_t)b->bs_xy | (uintptr_t)b->nbs | (uintptr_t)b->assoc | (uintptr_t)b->rate |
                    (uintptr_t)b->utility | (uintptr_t)b->done | (uintptr_t)b->metrics;
    // bulk async copies need 16-byte aligned global addresses; the FORK thread-per-env kernel also needs nbs
    env->tpe_bound_ok = (all & 15) == 0 && b->nbs != nullptr;
    if (env->tpe_shared) {
      all = (uintptr_t)b->pos | (uintptr_t)b->wp | (uintptr_t)b->t | (uintptr_t)b->episode | (uintptr_t)b->assoc |
            (uintptr_t)b->rate | (uintptr_t)b->utility | (uintptr_t)b->done | (uintptr_t)b->metrics;
      env->tpe_bound_ok = (all & 15) == 0;
    }
    uintptr_t pf = (uintptr_t)b->pos | (uintptr_t)b->wp | (uintptr_t)b->conn | (uintptr_t)b->actions;
    env->pf_bound_ok = (pf & 15) == 0;
  }
  if (!a.bs_per_env) {
    // a shared layout is constant for the life of the binding: fold the coordinates into the
    // kernel parameters (constant bank) for the specialised kernels
    int16_t xy[2 * mbe::kBigMaxB];
    MBE_CUDA(cudaMemcpy(xy, b->bs_xy, sizeof(int16_t) * 2 * a.B, cudaMemcpyDeviceToHost));
    for (int i = 0; i < a.B && i < mbe::kMaxSlots; ++i) {
      a.slot[i].x = xy[2 * i];
      a.slot[i].y = xy[2 * i + 1];
      a.slot[i].xf = (float)xy[2 * i];
      a.slot[i].yf = (float)xy[2 * i + 1];
    }
  }
  env->bound = true;
  return 0;
}

// Restricts the launch arguments to the env window [first, first + count): every stream is
// env-major, so a window is a pointer shift; Philox counters keep using the global env id.
static void shift_window(mbe::StepArgs& a, bool ma, size_t first, int count) {
  const size_t U = a.U, MW = (a.B + 31) >> 5;
  auto adv = [&](auto*& p, size_t n) {
    if (p) p += n;
  };
  adv(a.pos, first * U);
  adv(a.wp, first * U);
  adv(a.t, first);
  adv(a.episode, first);
  if (a.bs_per_env) adv(a.bs_xy, first * a.B);
  adv(a.nbs, first);
  adv(a.conn, first * U * MW);
  adv(a.assoc, first * U);
  adv(a.actions, first * U);
  adv(a.rate, first * U);
  adv(a.utility, first * U);
  adv(a.obs, first * U * a.F);
  adv(a.reward, ma ? first * U : first);
  adv(a.done, first);
  adv(a.metrics, first * 4);
  adv(a.dbg_snr, first * U * a.B);
  adv(a.inj_wp, first * U * (size_t)a.inj_k);
  adv(a.wp_cnt, first * U);
  a.env_offset += (unsigned)first;
  a.E = count;
}

// window_count == 0: all envs; otherwise the full step of envs [window_first, +window_count)
// (window_first a multiple of 32 keeps every bulk-copy address 16-byte aligned)
static int launch(mbe_env* env, int op, int phases, const uint8_t* mask, void* stream, size_t window_first = 0,
                  int window_count = 0) {
  if (!env) return fail("null handle");
  if (!env->bound) return fail("mbe_bind has not been called");
  MBE_ON_DEVICE(env->cfg.device);
  mbe::StepArgs a = env->args;
  a.op = op;
  a.phases = phases;
  a.reset_mask = mask;
  const bool window = window_count > 0;
  if (window) shift_window(a, env->cfg.handler == MBE_HANDLER_MA && env->cfg.mode == MBE_MODE_GYM, window_first, window_count);
  const int grid = env->big ? a.E : (a.E + a.epb - 1) / a.epb;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const bool gym = env->cfg.mode == MBE_MODE_GYM, ma = env->cfg.handler == MBE_HANDLER_MA;
  if (env->big) {
    env->big_fn<<<grid, mbe::kBigThreads, env->big_smem, st>>>(a);
    MBE_CUDA(cudaGetLastError());
    env->launches += 1;
    return 0;
  }
  // (a window that is not a whole number of 32-env warps falls through to the warp-segment kernels,
  // which handle ragged tails)
  if (env->tpe && env->tpe_bound_ok && a.E % 32 == 0 && op == mbe::OP_STEP && phases == MBE_PHASE_ALL && !a.dbg_snr) {
    cudaLaunchConfig_t lc = {};
    lc.gridDim = dim3(a.E / 32);
    lc.blockDim = dim3(32);
    lc.dynamicSmemBytes = env->tpe_smem;
    lc.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    lc.attrs = attr;
    lc.numAttrs = pdl_enabled() ? 1 : 0;
    MBE_CUDA(cudaLaunchKernelEx(&lc, env->tpe, a));
    env->launches += 1;
    return 0;
  }
  if (env->upt && op == mbe::OP_STEP && phases == MBE_PHASE_ALL && !a.dbg_snr) {
    a.epb = env->upt_epb;  // envs per CTA of this mapping (bounds and size of the obs bulk store)
    a.pf_dist = env->pf_bound_ok ? env->upt_pf : 0;
    cudaLaunchConfig_t lc = {};
    lc.gridDim = dim3((a.E + a.epb - 1) / a.epb);
    lc.blockDim = dim3(32 * MBE_UPT_WARPS);
    lc.dynamicSmemBytes = env->upt_smem;
    lc.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    lc.attrs = attr;
    lc.numAttrs = pdl_enabled_gym() ? 1 : 0;
    MBE_CUDA(cudaLaunchKernelEx(&lc, env->upt, a));
    env->launches += 1;
    return 0;
  }
  // the debug SNR output only exists in the generic kernel
  if (env->spec && op == mbe::OP_STEP && phases == MBE_PHASE_ALL && !a.dbg_snr)
  {
    const bool pdl = gym ? pdl_enabled_gym() : pdl_enabled();
    cudaLaunchConfig_t lc = {};
    const bool use_pipe = env->pipe != nullptr && !window;
    a.pf_dist = env->pf_bound_ok ? env->spec_pf : 0;
    lc.gridDim = dim3(use_pipe ? env->pipe_grid : grid);
    lc.blockDim = dim3(mbe::kThreads);
    lc.dynamicSmemBytes = use_pipe ? env->pipe_smem : env->spec_smem;
    lc.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    lc.attrs = attr;
    lc.numAttrs = pdl ? 1 : 0;
    MBE_CUDA(cudaLaunchKernelEx(&lc, use_pipe ? env->pipe : env->spec, a));
  }
  else if (!gym)
    mbe::step_kernel<0, 0><<<grid, mbe::kThreads, env->smem, st>>>(a);
  else if (!ma)
    mbe::step_kernel<1, 0><<<grid, mbe::kThreads, env->smem, st>>>(a);
  else
    mbe::step_kernel<1, 1><<<grid, mbe::kThreads, env->smem, st>>>(a);
  MBE_CUDA(cudaGetLastError());
  env->launches += 1;
  return 0;
}

int mbe_reset(mbe_env* env, const uint8_t* env_mask, void* stream) {
  return launch(env, mbe::OP_RESET, 0, env_mask, stream);
}

int mbe_step(mbe_env* env, void* stream) { return launch(env, mbe::OP_STEP, MBE_PHASE_ALL, nullptr, stream); }

int mbe_step_window(mbe_env* env, int first_env, int num_envs, void* stream) {
  if (!env) return fail("null handle");
  if (first_env < 0 || num_envs <= 0 || (long long)first_env + num_envs > env->cfg.num_envs)
    return fail("mbe_step_window: window [%d, +%d) outside 0..%d", first_env, num_envs, env->cfg.num_envs);
  if (first_env % 32) return fail("mbe_step_window: first_env must be a multiple of 32 (16-byte aligned slices)");
  return launch(env, mbe::OP_STEP, MBE_PHASE_ALL, nullptr, stream, (size_t)first_env, num_envs);
}

int mbe_stage(mbe_env* env, int phase_mask, void* stream) {
  if (phase_mask <= 0 || phase_mask > MBE_PHASE_ALL) return fail("mbe_stage: bad phase mask %d", phase_mask);
  return launch(env, mbe::OP_STEP, phase_mask, nullptr, stream);
}

int mbe_observe(mbe_env* env, void* stream) {
  if (env && env->cfg.mode != MBE_MODE_GYM) return fail("mbe_observe: GYM mode only");
  return launch(env, mbe::OP_OBSERVE, 0, nullptr, stream);
}

int mbe_channel(mbe_env* env, float* out_snr, uint32_t* out_elig, void* stream) {
  if (!env) return fail("null handle");
  if (!env->bound) return fail("mbe_bind has not been called");
  MBE_ON_DEVICE(env->cfg.device);
  const mbe::StepArgs& a = env->args;
  if (out_elig && a.B > 32) return fail("mbe_channel: the connectable bitmask output needs num_bs <= 32");
  const size_t total = (size_t)a.E * a.U;
  const int grid = (int)((total + mbe::kThreads - 1) / mbe::kThreads);
  const size_t smem = 4 * (size_t)(a.bs_per_env ? (mbe::kThreads / a.U + 2) * a.B : a.B) + a.B + 16;
  mbe::channel_kernel<<<grid, mbe::kThreads, smem, static_cast<cudaStream_t>(stream)>>>(a, out_snr, out_elig);
  MBE_CUDA(cudaGetLastError());
  env->launches += 1;
  return 0;
}

int mbe_accumulate_qoe(mbe_env* env, float* acc, float threshold, void* stream) {
  if (!env || !acc) return fail("mbe_accumulate_qoe: null argument");
  if (!env->bound) return fail("mbe_bind has not been called");
  if ((uintptr_t)acc & 15) return fail("mbe_accumulate_qoe: acc must be 16-byte aligned");
  MBE_ON_DEVICE(env->cfg.device);
  const mbe::StepArgs& a = env->args;
  const int per_blk = mbe::kThreads / 32;
  mbe::qoe_accumulate_kernel<<<(a.E + per_blk - 1) / per_blk, mbe::kThreads, 0, static_cast<cudaStream_t>(stream)>>>(
      a.utility, reinterpret_cast<float4*>(acc), a.E, a.U, threshold);
  MBE_CUDA(cudaGetLastError());
  env->launches += 1;
  return 0;
}

int mbe_rollout(mbe_env* env, int steps, float* qoe_acc, float threshold, const mbe_rollout_out* out, void* stream) {
  if (!env) return fail("null handle");
  if (!env->bound) return fail("mbe_bind has not been called");
  if (env->cfg.mode != MBE_MODE_FORK) return fail("mbe_rollout: FORK mode only (a GYM step needs fresh actions)");
  if (steps <= 0) return fail("mbe_rollout: steps must be positive");
  if ((uintptr_t)qoe_acc & 15) return fail("mbe_rollout: qoe_acc must be 16-byte aligned");
  MBE_ON_DEVICE(env->cfg.device);
  const mbe_rollout_out none = {};
  const mbe_rollout_out& o = out ? *out : none;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const mbe::StepArgs& a0 = env->args;
  const size_t EU = (size_t)a0.E * a0.U;
  const uintptr_t series = (uintptr_t)o.pos | (uintptr_t)o.wp | (uintptr_t)o.assoc | (uintptr_t)o.rate | (uintptr_t)o.utility;
  if (env->tpe_rollout && env->tpe_bound_ok && !a0.dbg_snr && (series & 15) == 0 && (o.rate == nullptr || a0.rate)) {
    // fused episode: one launch, state read and written once
    mbe::StepArgs a = a0;
    a.op = mbe::OP_STEP;
    a.phases = MBE_PHASE_ALL;
    a.reset_mask = nullptr;
    a.ro_steps = steps;
    a.qoe_thr = threshold;
    a.qoe_acc = reinterpret_cast<float4*>(qoe_acc);
    a.ro_pos = reinterpret_cast<uint32_t*>(o.pos);
    a.ro_wp = reinterpret_cast<uint32_t*>(o.wp);
    a.ro_assoc = o.assoc;
    a.ro_rate = o.rate;
    a.ro_util = o.utility;
    env->tpe_rollout<<<a.E / 32, 32, env->tpe_rollout_smem, st>>>(a);
    MBE_CUDA(cudaGetLastError());
    env->launches += 1;
    return 0;
  }
  // other shapes: the same episode as a sequence of step launches
  if ((o.pos || o.wp) && env->cfg.autoreset)
    return fail("mbe_rollout: per-step positions with autoreset need the fused kernel (MComCustom's 7 UEs x 10 BS slots "
                "or a scenario shape, num_envs a multiple of 32)");
  if (o.rate && !a0.rate) return fail("mbe_rollout: a rate series needs the rate buffer bound");
  for (int s = 0; s < steps; ++s) {
    if (int rc = launch(env, mbe::OP_STEP, MBE_PHASE_ALL, nullptr, stream)) return rc;
    if (qoe_acc)
      if (int rc = mbe_accumulate_qoe(env, qoe_acc, threshold, stream)) return rc;
    if (o.pos) MBE_CUDA(cudaMemcpyAsync(o.pos + 2 * EU * s, a0.pos, EU * 4, cudaMemcpyDeviceToDevice, st));
    if (o.wp) MBE_CUDA(cudaMemcpyAsync(o.wp + 2 * EU * s, a0.wp, EU * 4, cudaMemcpyDeviceToDevice, st));
    if (o.assoc) MBE_CUDA(cudaMemcpyAsync(o.assoc + EU * s, a0.assoc, EU * 4, cudaMemcpyDeviceToDevice, st));
    if (o.rate) MBE_CUDA(cudaMemcpyAsync(o.rate + EU * s, a0.rate, EU * 8, cudaMemcpyDeviceToDevice, st));
    if (o.utility) MBE_CUDA(cudaMemcpyAsync(o.utility + EU * s, a0.utility, EU * 4, cudaMemcpyDeviceToDevice, st));
  }
  return 0;
}

int64_t mbe_launch_count(const mbe_env* env) { return env ? env->launches : 0; }

const char* mbe_step_kernel_name(const mbe_env* env) {  // mirrors the dispatch order of launch()
  if (!env) return "";
  if (env->big) return "step_big_kernel";
  if (env->tpe && (!env->bound || env->tpe_bound_ok) && !env->args.dbg_snr) return "step_tpe_fork_kernel";
  if (env->upt && !env->args.dbg_snr) return "step_upt_kernel";
  if (env->spec && !env->args.dbg_snr) return env->pipe ? "step_pipe_kernel" : "step_spec_kernel";
  return "step_kernel";
}

int mbe_step_host(mbe_env* env, const int32_t* actions_host, float* obs_host, float* reward_host,
                  uint8_t* done_host, void* stream) {
  if (!env) return fail("null handle");
  if (!env->bound) return fail("mbe_bind has not been called");
  MBE_ON_DEVICE(env->cfg.device);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const mbe::StepArgs& a = env->args;
  const bool gym = env->cfg.mode == MBE_MODE_GYM, ma = env->cfg.handler == MBE_HANDLER_MA;
  if (!gym && (actions_host || obs_host || reward_host))
    return fail("mbe_step_host: actions, obs and reward only exist in GYM mode");
  // Env windows on two streams: the action upload and the step of window c+1 overlap the result
  // download of window c (PCIe is full duplex).  With observations the call is bound by their
  // download: plain FP32 rows by DMA straight into obs_host (48 GB/s on the B200 box, no host thread
  // touches them).  MBE_HOST_WIRE=compact ships them in the compact wire format instead
  // (mbe_host_wire.cuh, a third fewer bytes) and host threads expand window c into obs_host while
  // window c+1 is still arriving: that pays only on a host whose cores can write the FP32 rows faster
  // than PCIe delivers them -- on the 16-vCPU B200 box it measured 0.52-0.70x the DMA path
  // (profiles/README.md), hence opt-in.  MBE_HOST_WINDOWS / MBE_HOST_THREADS override the defaults.
  const char* wire_v = std::getenv("MBE_HOST_WIRE");
  const bool compact = obs_host != nullptr && wire_v && std::strcmp(wire_v, "compact") == 0;
  const char* wv = std::getenv("MBE_HOST_WINDOWS");
  const int want = wv ? std::max(1, std::atoi(wv)) : (obs_host ? (compact ? 8 : 4) : 1);
  constexpr int kAlign = 384;  // multiple of every kernel's envs-per-CTA and of 32 (16-byte aligned slices)
  int windows = std::min(want, std::max(1, a.E / (env->big ? kAlign : 8 * kAlign)));
  const int per = ((a.E + windows - 1) / windows + kAlign - 1) / kAlign * kAlign;
  windows = (a.E + per - 1) / per;
  if (compact && !env->wire_dev) {
    mbe::WireShape& w = env->wire_shape;
    w.U = a.U, w.B = a.B, w.F = a.F, w.MW = (a.B + 31) >> 5, w.ma = ma ? 1 : 0;
    w.W = w.MW * (ma ? 2 : 1);
    const size_t bytes = (size_t)a.E * w.bytes_per_env();
    MBE_CUDA(cudaMalloc(&env->wire_dev, bytes));
    MBE_CUDA(cudaMemset(env->wire_dev, 0, bytes));
    MBE_CUDA(cudaHostAlloc(&env->wire_pinned, bytes, cudaHostAllocDefault));
    const char* tv = std::getenv("MBE_HOST_THREADS");
    cpu_set_t set;
    int cpus = (sched_getaffinity(0, sizeof set, &set) == 0) ? CPU_COUNT(&set) : (int)std::thread::hardware_concurrency();
    // spinning workers: keep clear of oversubscription (driver threads, other ranks); 8 threads expand a
    // 65,536-env medium batch in 0.3 ms on the B200 box's host (profiles/host_expand_bench.cu)
    const int threads = tv ? std::max(1, std::atoi(tv)) : std::max(1, std::min(cpus / 2, 8));
    env->wire_pool.reset(new mbe::ExpandCrew(threads - 1));  // the calling thread is the last worker
  }
  if (compact)
    while ((int)env->wire_events.size() < windows) {
      cudaEvent_t e;
      MBE_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
      env->wire_events.push_back(e);
    }
  if (windows > 1 && !env->host_stream) {
    MBE_CUDA(cudaStreamCreateWithFlags(&env->host_stream, cudaStreamNonBlocking));
    MBE_CUDA(cudaEventCreateWithFlags(&env->host_fork, cudaEventDisableTiming));
    MBE_CUDA(cudaEventCreateWithFlags(&env->host_join, cudaEventDisableTiming));
  }
  // the expansion crew is woken first: its wake-up overlaps the launches below.  Any early return
  // (a failing CUDA call) releases it through the guard.
  struct CrewGuard {
    mbe::ExpandCrew* crew;
    ~CrewGuard() {
      if (crew) crew->abort();
    }
  } crew_guard{nullptr};
  if (compact) {
    mbe::ExpandCrew* crew = env->wire_pool.get();
    if (windows > mbe::ExpandCrew::kMaxWindows) return fail("mbe_step_host: at most %d env windows", mbe::ExpandCrew::kMaxWindows);
    const mbe::WireShape ws = env->wire_shape;
    const int chunks = std::max(1, 4 * (crew->size() + 1));
    const unsigned char* pinned = env->wire_pinned;
    const int E = a.E, UF = a.U * a.F;
    crew->begin(windows, chunks, [=](int w, int c) {
      const int first = w * per, count = std::min(per, E - first);
      const int lo = (int)((long long)count * c / chunks), hi = (int)((long long)count * (c + 1) / chunks);
      if (hi > lo)
        mbe::wire_expand_any(pinned + (size_t)first * ws.bytes_per_env(), obs_host + (size_t)first * UF, count, lo, hi, ws);
    });
    crew_guard.crew = crew;
  }
  if (windows > 1) {  // the side stream starts after everything already queued on the caller's stream
    MBE_CUDA(cudaEventRecord(env->host_fork, st));
    MBE_CUDA(cudaStreamWaitEvent(env->host_stream, env->host_fork, 0));
  }
  const size_t U = a.U;
  for (int w = 0, first = 0; first < a.E; ++w, first += per) {
    const int count = std::min(per, a.E - first);
    cudaStream_t s = (w & 1) ? env->host_stream : st;
    const size_t fu = (size_t)first * U, cu = (size_t)count * U;
    if (actions_host)
      MBE_CUDA(cudaMemcpyAsync(const_cast<int32_t*>(env->bufs.actions) + fu, actions_host + fu, cu * 4,
                               cudaMemcpyHostToDevice, s));
    if (int rc = (windows > 1 ? launch(env, mbe::OP_STEP, MBE_PHASE_ALL, nullptr, s, (size_t)first, count)
                              : launch(env, mbe::OP_STEP, MBE_PHASE_ALL, nullptr, s)))
      return rc;
    if (compact) {
      const mbe::WireShape& ws = env->wire_shape;
      const size_t off = (size_t)first * ws.bytes_per_env(), len = (size_t)count * ws.bytes_per_env();
      const int rows = count * a.U;
      mbe::wire_pack_kernel<<<(rows + 255) / 256, 256, 0, s>>>(env->bufs.obs, env->wire_dev + off, first, count, ws);
      MBE_CUDA(cudaGetLastError());
      MBE_CUDA(cudaMemcpyAsync(env->wire_pinned + off, env->wire_dev + off, len, cudaMemcpyDeviceToHost, s));
      MBE_CUDA(cudaEventRecord(env->wire_events[w], s));  // the window's observation bytes have landed
    } else if (obs_host) {
      MBE_CUDA(cudaMemcpyAsync(obs_host + fu * a.F, env->bufs.obs + fu * a.F, cu * a.F * 4, cudaMemcpyDeviceToHost, s));
    }
    if (reward_host) {
      const size_t f = ma ? fu : (size_t)first, c = ma ? cu : (size_t)count;
      MBE_CUDA(cudaMemcpyAsync(reward_host + f, env->bufs.reward + f, c * 4, cudaMemcpyDeviceToHost, s));
    }
    if (done_host)
      MBE_CUDA(cudaMemcpyAsync(done_host + first, env->bufs.done + first, (size_t)count, cudaMemcpyDeviceToHost, s));
  }
  if (windows > 1) {
    MBE_CUDA(cudaEventRecord(env->host_join, env->host_stream));
    MBE_CUDA(cudaStreamWaitEvent(st, env->host_join, 0));
  }
  if (compact) {
    // publish window after window as it lands; the crew expands window c while window c+1 is in flight
    mbe::ExpandCrew* crew = env->wire_pool.get();
    for (int w = 0; w < windows; ++w) {
      MBE_CUDA(cudaEventSynchronize(env->wire_events[w]));
      crew->publish(w + 1);
    }
    crew->finish();
    crew_guard.crew = nullptr;
  }
  MBE_CUDA(cudaStreamSynchronize(st));
  return 0;
}

}  // extern "C"
