/*
 * mbe.h -- C ABI of libmbe.so: batched, GPU-resident replacement for the per-step hot path
 * of mobile-env (reference fork yang-peilin/mobile-env-gan), i.e. MComCore.step over N
 * independent environments (reference mobile_env/core/base.py:230-296).
 *
 * The reference is pure Python and has no FFI; the entry points below are what a binding
 * for this path would need (INTEGRATION.md shows the ctypes stub a maintainer would add to
 * mobile_env/core/base.py).  Each entry cites the reference code it replaces.
 *
 * Conventions
 *   - plain C types only; every pointer inside mbe_buffers is a DEVICE pointer owned by the
 *     caller (e.g. torch tensors) and must stay alive while bound;
 *   - all calls are stream-ordered on the given cudaStream_t (passed as void*), enqueue
 *     only, never synchronise (except the *_host convenience entry, which says so);
 *   - return 0 on success, non-zero on error; mbe_last_error() gives a thread-local text;
 *   - a handle belongs to the device named in mbe_config.device; every call makes that device current
 *     for its duration and restores the caller's current device before returning;
 *   - a handle is not thread-safe, different handles are.
 *
 * Shapes: U <= 32 and B <= 32 run on the warp-segment kernels (an env is a slice of a warp; the
 * scenario shapes have fused compile-time variants: several UEs per thread for 15 x 4, a thread per
 * env for the fork's 7 UEs x 10 per-env BS slots -- mbe_step_kernel_name() tells which);
 * 32 < U <= 1024 or 32 < B <= 64, and the ProportionalFair / RateFair schedulers, run on the
 * block-per-env kernel (every entry point; needs width^2 + height^2 < 2^24).
 *
 * Data layout (structure of arrays, env-major; E = envs on this rank, U = UEs, B = BS slots,
 * MW = ceil(B/32), F = 2B+1 (central) or 4B+1 (multi-agent)):
 *   pos      int16 [E,U,2]   UE position (integers after the first move, movement.py:60)
 *   wp       int16 [E,U,2]   current waypoint; x < 0 = none (movement.py:44-47,54-56)
 *   t        int32 [E]       step clock of the episode (base.py:280, "time")
 *   episode  int32 [E]       episode counter (-1 before the first reset)
 *   bs_xy    int16 [B,2] or [E,B,2]  int-truncated BS coordinates (entities.py:24-26); a shared
 *                            [B,2] layout is read once by mbe_bind (re-bind after editing it)
 *   nbs      int32 [E]       live BS slots per env (random layouts, custom.py:68-77) or NULL
 *   conn     uint32 [E,U,MW] GYM: connection bitmask (bs2ue_connections, base.py:76)
 *   assoc    int32 [E,U]     FORK: BS index the UE is attached to, -1 = none (base.py:236-241)
 *   actions  int32 [E,U]     GYM: 0 = NOOP (base.py:29), a>0 toggles BS a-1
 *   rate     f64  [E,U]      per-UE total of the 2-decimal-rounded pair rates (base.py:413-435)
 *   utility  f32  [E,U]      scaled utility in [-1,1] (utilities.py:44-55; base.py:253-258)
 *   obs      f32  [E,U,F]    GYM observation written straight into the policy input
 *   reward   f32  [E] (central) or [E,U] (multi-agent)
 *   done     uint8 [E]       time_is_up (base.py:407-409)
 *   metrics  f32  [E,4]      number connections, number connected, mean utility, mean
 *                            datarate (metrics.py:5-28)
 */
#ifndef MBE_H_
#define MBE_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MBE_ABI_VERSION 2

enum { MBE_MODE_FORK = 0, MBE_MODE_GYM = 1 };
enum { MBE_HANDLER_CENTRAL = 0, MBE_HANDLER_MA = 1 };
enum { MBE_SCHED_RESOURCE_FAIR = 0, MBE_SCHED_PROPORTIONAL_FAIR = 1, MBE_SCHED_RATE_FAIR = 2 };
enum { MBE_BS_SHARED = 0, MBE_BS_PER_ENV = 1 };
enum { MBE_MAX_CLASSES = 16 };   /* link classes = BS classes x UE classes */
enum { MBE_MAX_UE_CLASSES = 8 };
/* mbe_config.flags */
enum {
  MBE_FLAG_GENERIC_KERNEL = 1,    /* never use the shape-specialised fused kernels */
  MBE_FLAG_SHARED_TRAJECTORY = 2, /* all envs draw the SAME UE initial positions and waypoints (the fork: movement
                                     reset_rng_episode=True makes every epoch replay one UE trajectory, base.py:130-134;
                                     with env index = epoch number only the BS layouts differ) */
};

/* phases of one step (bit mask for mbe_stage); the fused step runs all of them in one launch.
 * FORK order: MOVE, PRE, CLOCK.  GYM order: PRE, MOVE, CLOCK, POST. */
enum {
  MBE_PHASE_MOVE = 1,  /* RandomWaypointMovement.move         movement.py:42-62            */
  MBE_PHASE_PRE = 2,   /* connectivity / association / action, scheduler split, utility,
                          reward, metrics                     base.py:212-258, 421-435     */
  MBE_PHASE_CLOCK = 4, /* time += 1, departures, done          base.py:280-291, 407-409     */
  MBE_PHASE_POST = 8,  /* observation writer (GYM)                                          */
  MBE_PHASE_ALL = 15
};

/* One LINK class: all (BS, UE) pairs whose BS shares bw/freq/tx/height (entities.py:6-29) and whose UE
 * shares snr_threshold/noise/height (entities.py:32-57).  Link class of the pair (b, u) =
 * bs_class[b] * num_ue_classes + ue_class[u].  The tables are folded on the host in FP64 from the
 * channel's power_loss (channels.py:18-21, 132-146) through Channel.calculateSNR (channels.py:24-27),
 * evaluated at the integer offsets that realise each squared distance d2 (entities.py:24-26,52-54):
 *    d2max        largest d2 with snr > snr_threshold (base.py:212-214), -1: never connectable;
 *    rate_lut     HOST pointer, d2max+1 doubles, copied by mbe_create: Channel.datarate =
 *                 bw*log2(1+snr) (channels.py:78-83);
 *    log2(snr(d2)) for the FP32 observation path: l0 - k*log2(d2) for d2 >= 1 and l_zero for
 *                 d2 == 0 (EPSILON, channels.py:8) when log2snr_lut is NULL, else the table
 *                 log2snr_lut[d2] (HOST pointer, log2snr_len floats covering every d2 of the map,
 *                 copied by mbe_create) for losses that are not affine in log-distance. */
typedef struct mbe_link_class {
  double l0;
  double k;
  double l_zero;
  int32_t d2max; /* -1: never connectable */
  int32_t log2snr_len;
  const double* rate_lut;
  const float* log2snr_lut;
} mbe_link_class;
typedef mbe_link_class mbe_bs_class; /* ABI 1 name */

/* One UE class: the movement parameters of the UEs that share a parameter set (entities.py:32-57) */
typedef struct mbe_ue_class {
  double velocity;    /* ue.velocity (base.py:119) */
  int32_t move_d2max; /* largest integer d2 with sqrt(d2) <= velocity (movement.py:54) */
  int32_t reserved;
} mbe_ue_class;

typedef struct mbe_config {
  int32_t abi_version; /* MBE_ABI_VERSION */
  int32_t device;      /* CUDA device ordinal */
  int32_t num_envs;    /* E on this rank */
  int32_t num_ues;     /* U */
  int32_t num_bs;      /* B (slots) */
  int32_t mode;        /* MBE_MODE_* */
  int32_t handler;     /* MBE_HANDLER_* (GYM) */
  int32_t scheduler;   /* MBE_SCHED_* */
  int32_t bs_layout;   /* MBE_BS_* */
  int32_t bs_random_min, bs_random_max; /* >0: reset draws nbs in [min,max] and their
                                           coordinates per env and episode (custom.py:68-77) */
  int32_t autoreset;   /* GYM/FORK: re-initialise an env in the step that ends its episode */
  int32_t reset_rng_episode; /* movement_params.reset_rng_episode (base.py:130-134) */
  int32_t ep_time;     /* min(EP_MAX_TIME, max departure) (base.py:407-409) */
  int32_t move_d2max;  /* largest integer d2 with sqrt(d2) <= velocity (movement.py:54) */
  int64_t env_offset;  /* global id of local env 0 (multi-GPU sharding; Philox counters use it). Global env ids are
                          one 32-bit counter word: 0 <= env_offset and env_offset + num_envs <= 2^32 - 1, else
                          mbe_create fails (4.29e9 envs across the job) */
  uint64_t seed;       /* movement seed = config seed + 4 (base.py:155-170) */
  double width, height;/* map (base.py:103) */
  double velocity;     /* ue.velocity (base.py:119) of UE class 0 */
  double util_lower, util_upper, util_w1, util_w2, util_w3; /* utilities.py:30-55 */
  int32_t num_classes; /* number of BS classes, >= 1; num_classes * max(1, num_ue_classes) <= MBE_MAX_CLASSES */
  int32_t flags;       /* MBE_FLAG_* */
  mbe_link_class classes[MBE_MAX_CLASSES]; /* [bs class][ue class], ue class fastest */
  const uint8_t* bs_class; /* HOST [B] class id per BS slot, or NULL = all class 0 */
  int32_t num_ue_classes;  /* 0 or 1: all UEs alike (velocity / move_d2max above); else 2..MBE_MAX_UE_CLASSES */
  int32_t reserved2;
  const uint8_t* ue_class; /* HOST [U] class id per UE (required when num_ue_classes > 1) */
  mbe_ue_class ue_classes[MBE_MAX_UE_CLASSES]; /* used when num_ue_classes > 1 */
} mbe_config;

typedef struct mbe_buffers {
  int16_t* pos;
  int16_t* wp;
  int32_t* t;
  int32_t* episode;
  int16_t* bs_xy;
  int32_t* nbs;          /* optional */
  uint32_t* conn;        /* GYM */
  int32_t* assoc;        /* FORK */
  const int32_t* actions;/* GYM */
  double* rate;          /* optional in GYM */
  float* utility;        /* required (carried between phases) */
  float* obs;            /* GYM */
  float* reward;         /* GYM */
  uint8_t* done;
  float* metrics;        /* optional */
  float* dbg_snr;        /* optional f32 [E,U,B]: SNR at the positions the PRE phase sees */
  /* replay of reference trajectories: waypoints injected instead of Philox draws */
  const int16_t* inj_wp; /* optional int16 [E,U,K,2] */
  int32_t* wp_cnt;       /* int32 [E,U] draws consumed so far (required with inj_wp) */
  int32_t inj_k;         /* K */
  int32_t reserved;
} mbe_buffers;

typedef struct mbe_env mbe_env;

/* library / build information */
int mbe_abi_version(void);
const char* mbe_build_info(void);
const char* mbe_last_error(void);
/* sizeof of the ABI structs as the library was compiled (0 mbe_config, 1 mbe_buffers, 2 mbe_link_class,
 * 3 mbe_ue_class, 4 mbe_rollout_out; -1 otherwise): lets a binding in another language check its layout */
int mbe_struct_size(int which);

/* replaces MComCore.__init__ (base.py:32-100): validates, copies the constant tables */
int mbe_create(const mbe_config* cfg, mbe_env** out);
void mbe_destroy(mbe_env* env);

/* binds caller-owned device buffers (the SoA state of entities.py:6-57 / base.py:69-79) */
int mbe_bind(mbe_env* env, const mbe_buffers* bufs);

/* replaces MComCore.reset + MComCustom.reset (base.py:172-209, custom.py:40-62) for the envs
 * whose mask byte is non-zero (DEVICE pointer, NULL = all): clock, initial positions
 * (movement.py:64-72), optional BS layout, cleared connections, reset observation. */
int mbe_reset(mbe_env* env, const uint8_t* env_mask, void* stream);

/* replaces MComCore.step (base.py:230-296): one fused launch over all bound envs */
int mbe_step(mbe_env* env, void* stream);

/* the same fused step for the envs [first_env, first_env + num_envs) only (first_env a multiple of
 * 32; envs are independent, base.py:69-79): lets one handle keep several env groups in flight on
 * different streams (step group A while a policy reads group B's observations) */
int mbe_step_window(mbe_env* env, int first_env, int num_envs, void* stream);

/* the same step split into phases (MBE_PHASE_* mask), for per-stage parity and profiling */
int mbe_stage(mbe_env* env, int phase_mask, void* stream);

/* Channel.calculateSNR for every UE x BS pair (channels.py:24-27): f32 [E,U,B] into out_snr,
 * connectable bitmask uint32 [E,U,MW] into out_elig (either may be NULL) */
int mbe_channel(mbe_env* env, float* out_snr, uint32_t* out_elig, void* stream);

/* recompute the observation of the current state (after the caller edited pos / conn) */
int mbe_observe(mbe_env* env, void* stream);

/* adds this step's two-decimal QoE values (base.py:269) to acc f32 [E,4] = (sum q, sum q^2,
 * #q < threshold, #values): the statistics of chooseBaseStation.ipynb cell 5 `qoeValue` */
int mbe_accumulate_qoe(mbe_env* env, float* acc, float threshold, void* stream);

/* per-step series of a fused episode: DEVICE pointers, any may be NULL; 16-byte aligned */
typedef struct mbe_rollout_out {
  int16_t* pos;   /* [T,E,U,2] positions after each step's move (what base.py:298-404 dumps per step) */
  int16_t* wp;    /* [T,E,U,2] waypoints after the move; (-1,-1) = arrived this step (movement.py:54-56) */
  int32_t* assoc; /* [T,E,U] serving BS or -1 (base.py:236-241) */
  double* rate;   /* [T,E,U] rounded data rates (base.py:435) */
  float* utility; /* [T,E,U] scaled utility = QoE (base.py:253-258) */
} mbe_rollout_out;

/* FORK mode: `steps` consecutive MComCore.step calls (base.py:230-296; the fork's collect loop
 * `for s in range(20): env.step(e, s)`) in ONE launch -- state stays on chip between the steps.
 * qoe_acc (f32 [E,4], may be NULL) accumulates the statistics of mbe_accumulate_qoe over all steps;
 * `out` (may be NULL) receives the per-step series.  After the call the bound buffers hold exactly
 * what `steps` calls of mbe_step would have left.  Shapes without the fused kernel run the same
 * episode as a sequence of step launches (same results). */
int mbe_rollout(mbe_env* env, int steps, float* qoe_acc, float threshold, const mbe_rollout_out* out, void* stream);

/* number of kernel launches this handle has enqueued so far */
int64_t mbe_launch_count(const mbe_env* env);

/* which kernel family mbe_step dispatches to for this handle's shape and bound buffers, e.g.
 * "step_upt_kernel" (several UEs per thread), "step_spec_kernel" (warp segment, compile-time shape),
 * "step_tpe_fork_kernel" (thread per env), "step_big_kernel" (block per env), "step_kernel" (generic) */
const char* mbe_step_kernel_name(const mbe_env* env);

/* Host-buffer convenience: H2D actions, step, D2H obs/reward/done, then synchronises the
 * stream.  Host pointers should be pinned.  NULL pointers are skipped.  Large batches with
 * observations are processed as env windows on two streams so uploads and the step overlap the
 * download (MBE_HOST_WINDOWS=n overrides the window count). */
int mbe_step_host(mbe_env* env, const int32_t* actions_host, float* obs_host,
                  float* reward_host, uint8_t* done_host, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MBE_H_ */
