// Device-side building blocks of the batched MComCore.step (sm_100a).
// Reference semantics are cited per function (paths relative to the reference repo).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace mbe {

constexpr int kWarpsPerBlock = 8;
constexpr int kThreads = kWarpsPerBlock * 32;
constexpr unsigned kFull = 0xffffffffu;

enum Op : int { OP_STEP = 0, OP_RESET = 1, OP_OBSERVE = 2 };
enum Purpose : unsigned { P_WAYPOINT = 0, P_INITPOS = 1, P_BSLAYOUT = 2 };

struct ClassDev {
  float l0_hi, l0_lo;  // log2 snr at d2 = 1, split hi+lo so the FP32 chain keeps ~2^-30 of it
  float k_hi, k_lo;    // slope per log2(d2)
  float l_zero;        // log2 snr at d2 == 0
  int d2max;           // connectable iff d2 <= d2max
  const double* lut;   // rate_lut[d2], d2 in [0, d2max]
};

struct StepArgs {
  // geometry of the launch
  int E, U, B, F;
  int epw;  // envs per warp
  int epb;  // envs per block
  int op, phases;
  unsigned env_offset;
  // scenario
  int ep_time, autoreset, reset_rng_episode, bs_per_env, bs_rand_min, bs_rand_max;
  unsigned seed_lo, seed_hi;
  double width, height, velocity;
  int move_d2max;
  // utility: u = clip(util_c * log2(w2 + r), lo, hi); scaled = (u - lo) * util_scale - 1
  float util_c, util_w2, util_lo, util_hi, util_scale;
  int n_classes;
  ClassDev cls[8];
  const uint8_t* bs_class;  // device [B] or nullptr
  // bound buffers (see include/mbe.h)
  uint32_t* pos;
  uint32_t* wp;
  int32_t* t;
  int32_t* episode;
  uint32_t* bs_xy;
  int32_t* nbs;
  uint32_t* conn;
  int32_t* assoc;
  const int32_t* actions;
  double* rate;
  float* utility;
  float* obs;
  float* reward;
  uint8_t* done;
  float* metrics;
  float* dbg_snr;
  const uint32_t* inj_wp;
  int32_t* wp_cnt;
  int inj_k;
  const uint8_t* reset_mask;
  int obs_bulk_ok;  // obs base is 16B aligned -> whole-block TMA bulk store allowed
};

// ---------------------------------------------------------------------------------------
// Philox4x32-10 (Salmon et al. 2011): counter-based replacement for the shared PCG64 stream
// behind RandomWaypointMovement (movement.py:16-18,44-47,64-72).
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k) {
  constexpr unsigned M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    unsigned hi0 = __umulhi(M0, c.x), lo0 = M0 * c.x;
    unsigned hi1 = __umulhi(M1, c.z), lo1 = M1 * c.z;
    c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
    k.x += W0;
    k.y += W1;
  }
  return c;
}

// x = int(u0 * W), y = int(u1 * H) with u = r * 2^-32 in FP64: the counter-based analogue of
// int(rng.uniform(0, W)) (movement.py:45-46, 69-70).  Counter = (env gid, ue, t, purpose+4*salt).
__device__ __forceinline__ void philox_point(const StepArgs& a, unsigned gid, unsigned ue, unsigned t,
                                             unsigned purpose, unsigned salt, int& x, int& y) {
  uint4 r = philox4x32_10(make_uint4(gid, ue, t, purpose + 4u * salt), make_uint2(a.seed_lo, a.seed_hi));
  x = (int)((double)r.x * 0x1p-32 * a.width);
  y = (int)((double)r.y * 0x1p-32 * a.height);
}

__device__ __forceinline__ int philox_bs_count(const StepArgs& a, unsigned gid, unsigned salt) {
  uint4 r = philox4x32_10(make_uint4(gid, 0xFFFFu, 0xFFFFu, P_BSLAYOUT + 4u * salt),
                          make_uint2(a.seed_lo, a.seed_hi));
  return a.bs_rand_min + (int)((double)r.x * 0x1p-32 * (double)(a.bs_rand_max - a.bs_rand_min + 1));
}

__device__ __forceinline__ uint32_t pack_xy(int x, int y) {
  return ((uint32_t)x & 0xffffu) | ((uint32_t)y << 16);
}
__device__ __forceinline__ void unpack_xy(uint32_t p, int& x, int& y) {
  x = (int)(int16_t)(p & 0xffffu);
  y = (int)(int16_t)(p >> 16);
}

// ---------------------------------------------------------------------------------------
// RandomWaypointMovement.move once the waypoint exists (movement.py:49-62).  FP64 with the
// reference's operation order: pos + (velocity * v) / norm(v), np.round (half-even), astype(int).
// "norm <= velocity" is the integer test d2 <= move_d2max (host-folded, exact).
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ bool move_exact(const StepArgs& a, int& x, int& y, int wx, int wy) {
  int dx = wx - x, dy = wy - y;
  int d2 = dx * dx + dy * dy;
  if (d2 <= a.move_d2max) {
    x = wx;
    y = wy;
    return true;  // arrived: snap and pop the waypoint (movement.py:54-56)
  }
  double norm = sqrt((double)d2);
  x = (int)rint((double)x + (a.velocity * (double)dx) / norm);
  y = (int)rint((double)y + (a.velocity * (double)dy) / norm);
  return false;
}

// log2 of the SNR (channels.py:24-27 + 132-146 folded): l0 - k*log2(d2); SFU lg2 on FP32.
__device__ __forceinline__ float log2_snr(const ClassDev& c, int d2) {
  if (d2 == 0) return c.l_zero;
  float lg = __log2f((float)d2);
  float l = fmaf(-c.k_hi, lg, c.l0_hi);
  return fmaf(-c.k_lo, lg, l) + c.l0_lo;
}

// BoundedLogUtility.calculateUtility + scaleUtility (utilities.py:44-55)
__device__ __forceinline__ float scaled_utility(const StepArgs& a, double rate) {
  if (rate <= 0.0) return -1.0f;  // lower -> scaled -1
  float u = a.util_c * log2f(a.util_w2 + (float)rate);
  u = fminf(fmaxf(u, a.util_lo), a.util_hi);
  return fmaf(u - a.util_lo, a.util_scale, -1.0f);
}

// round(rate, 2) of np.float64: rint(x*100)/100 (base.py:435)
__device__ __forceinline__ double round2(double r) { return rint(r * 100.0) / 100.0; }

// Sum over the lanes of one env (contiguous segment of U lanes inside the warp), result
// broadcast to every lane of the segment.  Fixed tree order => deterministic.
__device__ __forceinline__ float seg_sum(float v, int u, int U, int lane) {
  for (int off = 1; off < U; off <<= 1) {
    float o = __shfl_down_sync(kFull, v, off);
    if (u + off < U) v += o;
  }
  return __shfl_sync(kFull, v, lane - u);
}

}  // namespace mbe
