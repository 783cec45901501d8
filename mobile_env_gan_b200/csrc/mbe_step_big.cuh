// Block-per-env fused step kernel for the wide shapes (32 < U <= 1024 UEs or 32 < B <= 64 BS
// slots, e.g. the synthetic 64 x 512 scale-up): one CTA of 256 threads owns one env, every thread
// keeps up to 4 UEs in registers, the BS table and the per-BS accumulators live in shared memory.
// Per-BS counts / proportional-fair totals / multi-agent BS utilities are integer (fixed-point)
// shared-memory atomics, so every result is independent of the order threads arrive in.
// Same arithmetic helpers as the warp-segment kernels (mbe_device.cuh).
#pragma once
#include "mbe_device.cuh"

namespace mbe {

constexpr int kBigThreads = 256;
constexpr int kBigMaxI = 4;     // UEs per thread  => U <= 1024
constexpr int kBigMaxB = 64;    // BS slots        => two 32-bit mask words

struct BigSmem {
  uint32_t bs[kBigMaxB];
  int cnt[kBigMaxB];
  unsigned long long pf_tot[kBigMaxB];
  long long bsu_acc[kBigMaxB];
  float bsu[kBigMaxB];
  uint8_t cls[kBigMaxB];
  float red_f[3][kBigThreads / 32];
  int red_i[2][kBigThreads / 32];
  float usum, rsum;
  int csum, ncon;
  // per-warp rows of the observation writer: first the log2(snr) of every BS (one pass), then
  // reused as the transpose tile of the other feature segments (odd stride: conflict-free)
  float rows[kBigThreads / 32][32][kBigMaxB + 1];
};

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) v += __shfl_down_sync(kFull, v, off);
  return v;
}
__device__ __forceinline__ int warp_sum_i(int v) {
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) v += __shfl_down_sync(kFull, v, off);
  return v;
}

// one link's share of its BS (schedules.py:20-22 / proportional fair), rounded like base.py:435
__device__ __forceinline__ double link_share(const StepArgs& a, const ClassDev& c, int d2, int n,
                                             unsigned long long pf_tot) {
  const double raw = c.lut0[d2];  // Channel.datarate (channels.py:78-83), FP64 table
  double share;
  if (a.scheduler == 1) {
    const double tot = __ll2double_rn((long long)pf_tot) * (1.0 / 1048576.0);
    share = (raw * raw) / tot;
  } else if (a.scheduler == 2) {  // RateFair: the same 1 / sum(1/r) for every UE of the BS
    const double tot = __ll2double_rn((long long)pf_tot) * 0x1p-50;
    share = 1.0 / tot;
  } else {
    share = raw / (double)n;
  }
  return rint(share * 100.0) / 100.0;
}

template <int MODE, int HANDLER>
__global__ void __launch_bounds__(kBigThreads) step_big_kernel(const __grid_constant__ StepArgs a) {
  constexpr bool GYM = (MODE == 1);
  constexpr bool MA = (HANDLER == 1);
  extern __shared__ __align__(128) unsigned char smem_raw[];
  BigSmem& s = *reinterpret_cast<BigSmem*>(smem_raw);
  const int U = a.U, B = a.B, F = a.F, MW = (B + 31) >> 5;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int env = blockIdx.x;
  const unsigned gid = a.env_offset + (unsigned)env;
  const int op = a.op;

  // ---- P0: env scalars, BS table, accumulators ----
  int t_e = a.t[env], epi = a.episode[env];
  int nb = a.nbs ? a.nbs[env] : B;
  if (tid < B) {
    s.bs[tid] = a.bs_per_env ? a.bs_xy[(size_t)env * B + tid] : a.bs_xy[tid];
    s.cls[tid] = a.bs_class ? a.bs_class[tid] : (uint8_t)0;
    s.cnt[tid] = 0;
    s.pf_tot[tid] = 0ull;
    s.bsu_acc[tid] = 0ll;
    s.bsu[tid] = -1.0f;
  }
  bool touched = true;
  if (op == OP_RESET) touched = (a.reset_mask == nullptr) || (a.reset_mask[env] != 0);
  __syncthreads();

  // ---- per-UE registers ----
  int x[kBigMaxI], y[kBigMaxI], wx[kBigMaxI], wy[kBigMaxI];
  uint32_t c0[kBigMaxI], c1[kBigMaxI], e0[kBigMaxI], e1[kBigMaxI];
  float util[kBigMaxI];
  double rate[kBigMaxI];
  int best[kBigMaxI], bestd2[kBigMaxI];
#pragma unroll
  for (int i = 0; i < kBigMaxI; ++i) {
    const int u = tid + i * kBigThreads;
    x[i] = y[i] = 0;
    wx[i] = wy[i] = -1;
    c0[i] = c1[i] = e0[i] = e1[i] = 0;
    util[i] = -1.0f;
    rate[i] = 0.0;
    best[i] = -1;
    bestd2[i] = 0x7fffffff;
    if (u < U) {
      const size_t idx = (size_t)env * U + u;
      unpack_xy(a.pos[idx], x[i], y[i]);
      unpack_xy(a.wp[idx], wx[i], wy[i]);
      if (GYM) {
        c0[i] = a.conn[idx * MW];
        if (MW > 1) c1[i] = a.conn[idx * MW + 1];
      }
    }
  }
  bool done = false, fresh = false;
  const bool one_class = a.n_classes == 1 && a.n_ue_classes == 1;
  const int d2max0 = a.cls[0].d2max;
  // link class of (b, UE i of this thread) = bs_class[b] * n_ue_classes + ue_class[u] (entities.py:6-57)
  int ucls[kBigMaxI];
#pragma unroll
  for (int i = 0; i < kBigMaxI; ++i) {
    const int u = tid + i * kBigThreads;
    ucls[i] = (a.ue_class && u < U) ? (int)a.ue_class[u] : 0;
  }
  const int nuc = a.n_ue_classes;
  auto link = [&](int i, int b) -> const ClassDev& { return a.cls[(int)s.cls[b] * nuc + ucls[i]]; };

  auto d2_to = [&](int i, int b) {
    int bx, by;
    unpack_xy(s.bs[b], bx, by);
    int dx = x[i] - bx, dy = y[i] - by;
    return dx * dx + dy * dy;
  };
  auto has_bit = [&](uint32_t lo, uint32_t hi, int b) { return (((b < 32) ? (lo >> b) : (hi >> (b - 32))) & 1u) != 0; };

  auto phase_move = [&]() {
#pragma unroll
    for (int i = 0; i < kBigMaxI; ++i) {
      const int u = tid + i * kBigThreads;
      if (u >= U) continue;
      if (wx[i] < 0) next_waypoint(a, gid, (unsigned)u, (size_t)env * U + u, t_e, epi, true, wx[i], wy[i]);
      if (move_ue(a.mv[ucls[i]], x[i], y[i], wx[i], wy[i])) wx[i] = wy[i] = -1;
    }
  };

  // block-wide sums in a fixed order (warp tree, then warp partials in ascending order)
  auto block_sums = [&](float fu, float fr, int ic, int in) {
    fu = warp_sum(fu);
    fr = warp_sum(fr);
    ic = warp_sum_i(ic);
    in = warp_sum_i(in);
    if (lane == 0) {
      s.red_f[0][warp] = fu;
      s.red_f[1][warp] = fr;
      s.red_i[0][warp] = ic;
      s.red_i[1][warp] = in;
    }
    __syncthreads();
    if (tid == 0) {
      float su = 0.0f, sr = 0.0f;
      int sc = 0, sn = 0;
      for (int w = 0; w < kBigThreads / 32; ++w) {
        su += s.red_f[0][w];
        sr += s.red_f[1][w];
        sc += s.red_i[0][w];
        sn += s.red_i[1][w];
      }
      s.usum = su;
      s.rsum = sr;
      s.csum = sc;
      s.ncon = sn;
    }
    __syncthreads();
  };

  auto reinit_all = [&]() {
    // MComCore.reset + MComCustom.reset for this env (base.py:172-209, custom.py:40-62)
    epi += 1;
    t_e = 0;
    fresh = true;
    if (a.bs_rand_max > 0 && a.bs_per_env) {
      nb = philox_bs_count(a, gid, (unsigned)epi);
      __syncthreads();
      if (tid < B) {
        int bx = 0, by = 0;
        if (tid < nb) philox_point(a, gid, (unsigned)tid, 0u, P_BSLAYOUT, (unsigned)epi, bx, by);
        uint32_t p = pack_xy(bx, by);
        s.bs[tid] = p;
        a.bs_xy[(size_t)env * B + tid] = p;
      }
      if (tid == 0 && a.nbs) a.nbs[env] = nb;
    }
    __syncthreads();
    if (tid < B) {
      s.cnt[tid] = 0;
      s.bsu[tid] = -1.0f;
    }
#pragma unroll
    for (int i = 0; i < kBigMaxI; ++i) {
      const int u = tid + i * kBigThreads;
      if (u >= U) continue;
      c0[i] = c1[i] = 0;
      wx[i] = wy[i] = -1;
      philox_point(a, gid, (unsigned)u, 0u, P_INITPOS, a.reset_rng_episode ? 0u : (unsigned)epi, x[i], y[i]);
      if (a.inj_wp) a.wp_cnt[(size_t)env * U + u] = 0;
    }
    __syncthreads();
  };

  if (op == OP_RESET) {
    if (touched) {
      reinit_all();
#pragma unroll
      for (int i = 0; i < kBigMaxI; ++i) {
        const int u = tid + i * kBigThreads;
        if (u >= U) continue;
        const size_t idx = (size_t)env * U + u;
        a.utility[idx] = -1.0f;
        if (a.rate) a.rate[idx] = 0.0;
        if (!GYM) a.assoc[idx] = -1;
        if (GYM && MA) a.reward[idx] = 0.0f;
      }
      if (tid == 0) {
        a.done[env] = 0;
        if (GYM && !MA) a.reward[env] = 0.0f;
      }
    }
  } else {
    // ================= one step =================
    if (!GYM) phase_move();  // FORK moves first (base.py:232-233)

    // ---- P1: connectivity, association / action, per-BS accumulators ----
#pragma unroll
    for (int i = 0; i < kBigMaxI; ++i) {
      const int u = tid + i * kBigThreads;
      if (u >= U) continue;
      const size_t idx = (size_t)env * U + u;
      for (int w = 0; w < 2; ++w) {  // one 32-bit mask word at a time (no per-BS word select)
        uint32_t ew = 0;
        const int b_lo = 32 * w, b_hi = min(nb, 32 * w + 32);
        for (int b = b_lo; b < b_hi; ++b) {
          const int d2 = d2_to(i, b);
          if (d2 <= (one_class ? d2max0 : link(i, b).d2max)) {  // check_connectivity (base.py:212-214)
            ew |= 1u << (b - b_lo);
            if (!GYM && d2 < bestd2[i]) {  // nearest connectable BS, first minimum (base.py:240)
              best[i] = b;
              bestd2[i] = d2;
            }
          }
        }
        if (w) e1[i] = ew; else e0[i] = ew;
      }
      if (GYM) {
        c0[i] &= e0[i];  // update_connections (base.py:221-227)
        c1[i] &= e1[i];
        const int act = a.actions[idx];
        if (act > 0 && act <= nb) {  // NOOP_ACTION = 0 (base.py:29)
          const int b = act - 1;
          uint32_t& cw = (b < 32) ? c0[i] : c1[i];
          const uint32_t ew = (b < 32) ? e0[i] : e1[i];
          const uint32_t bit = 1u << (b & 31);
          cw = (cw & bit) ? (cw & ~bit) : (cw | (ew & bit));
        }
      } else if (best[i] >= 0) {
        if (best[i] < 32) c0[i] = 1u << best[i]; else c1[i] = 1u << (best[i] - 32);
      }
      for (int w = 0; w < 2; ++w) {
        uint32_t m = w ? c1[i] : c0[i];
        while (m) {
          const int b = (__ffs(m) - 1) + 32 * w;
          m &= m - 1;
          atomicAdd(&s.cnt[b], 1);
          if (a.scheduler == 1) {
            const double raw = link(i, b).lut0[d2_to(i, b)];
            atomicAdd(&s.pf_tot[b], (unsigned long long)__double2ll_rn(raw * 1048576.0));
          } else if (a.scheduler == 2) {
            const double raw = link(i, b).lut0[d2_to(i, b)];
            atomicAdd(&s.pf_tot[b], (unsigned long long)__double2ll_rn(0x1p50 / raw));
          }
        }
      }
    }
    __syncthreads();

    // ---- P2: split, rounding, utility (base.py:421-435, 413-418, 253-258) ----
    float fu = 0.0f, fr = 0.0f;
    int ic = 0, in = 0;
#pragma unroll
    for (int i = 0; i < kBigMaxI; ++i) {
      const int u = tid + i * kBigThreads;
      if (u >= U) continue;
      const size_t idx = (size_t)env * U + u;
      double r = 0.0;
      for (int w = 0; w < 2; ++w) {
        uint32_t m = w ? c1[i] : c0[i];
        while (m) {  // ascending bs order = the reference's bs-major accumulation
          const int b = (__ffs(m) - 1) + 32 * w;
          m &= m - 1;
          r += link_share(a, link(i, b), d2_to(i, b), s.cnt[b], s.pf_tot[b]);
        }
      }
      rate[i] = r;
      util[i] = scaled_utility(a, r);
      if (a.rate) a.rate[idx] = r;
      a.utility[idx] = util[i];
      if (!GYM) a.assoc[idx] = best[i];
      fu += util[i];
      fr += (float)r;
      ic += __popc(c0[i]) + __popc(c1[i]);
      in += (c0[i] | c1[i]) ? 1 : 0;
      if (GYM && MA) {
        const long long q = __double2ll_rn((double)util[i] * 4294967296.0);
        for (int w = 0; w < 2; ++w) {
          uint32_t m = w ? c1[i] : c0[i];
          while (m) {
            const int b = (__ffs(m) - 1) + 32 * w;
            m &= m - 1;
            atomicAdd((unsigned long long*)&s.bsu_acc[b], (unsigned long long)q);
          }
        }
      }
    }
    block_sums(fu, fr, ic, in);
    if (GYM && MA) {
      if (tid < B)  // allStationUtilities (base.py:438-447)
        s.bsu[tid] = s.cnt[tid] ? (float)(__ll2double_rn(s.bsu_acc[tid]) * (1.0 / 4294967296.0) / (double)s.cnt[tid])
                                : -1.0f;
      __syncthreads();
#pragma unroll
      for (int i = 0; i < kBigMaxI; ++i) {
        const int u = tid + i * kBigThreads;
        if (u >= U) continue;
        float nu = 0.0f;
        int ncnt = 0;
        for (int b = 0; b < nb; ++b)
          if (has_bit(e0[i], e1[i], b)) {  // available_connections (base.py:216-218)
            nu += s.bsu[b];
            ncnt += s.cnt[b];
          }
        a.reward[(size_t)env * U + u] = (nu + util[i]) / (float)(ncnt + 1);
      }
    }
    if (tid == 0) {
      if (GYM && !MA) a.reward[env] = s.usum * a.inv_U;  // mean utility (metrics.py:25-28)
      if (a.metrics) {
        const float nc = (float)s.ncon;
        reinterpret_cast<float4*>(a.metrics)[env] =
            make_float4((float)s.csum, nc, s.usum * a.inv_U, mean_or_zero(s.rsum, nc));
      }
    }

    if (GYM) phase_move();

    // ---- CLOCK (base.py:280-291, 407-409) ----
    t_e += 1;
    done = t_e >= a.ep_time;
    if (tid == 0) a.done[env] = done ? 1 : 0;
    if (done) {
      __syncthreads();
#pragma unroll
      for (int i = 0; i < kBigMaxI; ++i) c0[i] = c1[i] = 0;  // everyone leaves at ep_time
      if (tid < B) {
        s.cnt[tid] = 0;
        s.bsu[tid] = -1.0f;
      }
      if (a.autoreset) reinit_all();
      else __syncthreads();
    }
  }

  // ---- POST: observation rows (GYM) ----
  // A lane owns one UE row of F floats; the rows of the warp's 32 UEs are contiguous in HBM.  Each
  // feature segment (connections | snr ratios | utility | broadcast utilities | broadcast counts)
  // is produced 32 columns at a time into a padded shared tile and written back transposed, so a
  // store instruction covers 128 contiguous bytes of one row.  log2(snr) is evaluated once per
  // pair and parked in shared memory until the row maximum is known.
  if (GYM && touched) {
    float (*tile)[kBigMaxB + 1] = s.rows[warp];
#pragma unroll
    for (int i = 0; i < kBigMaxI; ++i) {
      const int u0 = warp * 32 + i * kBigThreads;  // first UE of this warp's row group
      if (u0 >= U) continue;                        // warp-uniform
      const bool active = (u0 + lane < U) && !(done && !fresh);
      float* gbase = a.obs + ((size_t)env * U + u0) * F;
      const int nrows = min(32, U - u0);
      // transposed write-back of tile columns [c0, c0 + n) to feature columns [f0, f0 + n)
      // (full 32-row tiles of the BASELINE synthetic shape, F = 129 / 257, unroll completely with the
      // row stride as an immediate: one shared load + one store per row, no address arithmetic)
      auto flush = [&](int f0, int c0_, int n) {
        __syncwarp();
        if (lane < n) {
          float* g = gbase + f0 + lane;
          const float* t = &tile[0][c0_ + lane];
          if (nrows == 32 && F == 129) {
#pragma unroll
            for (int r = 0; r < 32; ++r) g[r * 129] = t[r * (kBigMaxB + 1)];
          } else if (nrows == 32 && F == 257) {
#pragma unroll
            for (int r = 0; r < 32; ++r) g[r * 257] = t[r * (kBigMaxB + 1)];
          } else {
#pragma unroll 8
            for (int r = 0; r < nrows; ++r) g[(size_t)r * F] = t[r * (kBigMaxB + 1)];
          }
        }
        __syncwarp();
      };
      // pass over the BSs: log2 snr, its maximum, connectable mask, MA count total
      float lmax = -INFINITY, tot = 0.0f;
      uint32_t k0 = 0, k1 = 0;
      if (active)
        for (int b = 0; b < nb; ++b) {
          const ClassDev& c = link(i, b);
          const int d2 = d2_to(i, b);
          const float l = log2_snr_obs(c, d2);
          tile[lane][b] = l;
          lmax = fmaxf(lmax, l);
          if (d2 <= c.d2max) {
            if (b < 32) k0 |= 1u << b; else k1 |= 1u << (b - 32);
            tot += (float)s.cnt[b];
          }
        }
      const float inv_tot = 1.0f / fmaxf(1.0f, tot);
      // (2) snr / max snr, in place over the parked log2 values
      for (int b = 0; b < B; ++b) tile[lane][b] = (active && b < nb) ? ex2_sfu(tile[lane][b] - lmax) : 0.0f;
      for (int b0 = 0; b0 < B; b0 += 32) flush(B + b0, b0, min(32, B - b0));
      // (1) connection one-hot
      {
        const uint32_t m0 = active ? c0[i] : 0u, m1 = active ? c1[i] : 0u;
        for (int b = 0; b < min(B, 32); ++b) tile[lane][b] = ((m0 >> b) & 1u) ? 1.0f : 0.0f;
        for (int b = 32; b < B; ++b) tile[lane][b] = ((m1 >> (b - 32)) & 1u) ? 1.0f : 0.0f;
      }
      for (int b0 = 0; b0 < B; b0 += 32) flush(b0, b0, min(32, B - b0));
      tile[lane][0] = active ? ((fresh || t_e == 0) ? -1.0f : util[i]) : 0.0f;  // (3) own utility
      flush(2 * B, 0, 1);
      if (MA) {
        // (4) broadcast BS utilities
        for (int b = 0; b < B; ++b) tile[lane][b] = active ? (has_bit(k0, k1, b) ? s.bsu[b] : -1.0f) : 0.0f;
        for (int b0 = 0; b0 < B; b0 += 32) flush(2 * B + 1 + b0, b0, min(32, B - b0));
        // (5) broadcast connection counts, normalised
        for (int b = 0; b < B; ++b)
          tile[lane][b] = (active && has_bit(k0, k1, b)) ? (float)s.cnt[b] * inv_tot : 0.0f;
        for (int b0 = 0; b0 < B; b0 += 32) flush(3 * B + 1 + b0, b0, min(32, B - b0));
      }
    }
  }

  // ---- store state ----
  if (touched) {
#pragma unroll
    for (int i = 0; i < kBigMaxI; ++i) {
      const int u = tid + i * kBigThreads;
      if (u >= U) continue;
      const size_t idx = (size_t)env * U + u;
      a.pos[idx] = pack_xy(x[i], y[i]);
      a.wp[idx] = pack_xy(wx[i], wy[i]);
      if (GYM) {
        a.conn[idx * MW] = c0[i];
        if (MW > 1) a.conn[idx * MW + 1] = c1[i];
      }
    }
    if (tid == 0) {
      a.t[env] = t_e;
      if (fresh) a.episode[env] = epi;
    }
  }
}

}  // namespace mbe
