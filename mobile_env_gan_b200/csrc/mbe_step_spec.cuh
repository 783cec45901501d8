// Specialised fused step kernels: U (UEs) and B (BS slots) are template constants, so every
// loop over b unrolls, BS coordinates of a shared layout are constant-bank operands (no
// shared-memory table, no unpacking) and the per-pair values stay in registers.
// Same warp-segment mapping and the same arithmetic helpers as the generic kernel
// (mbe_step.cuh), which remains the path for reset / observe / split phases and for shapes
// without an instantiation.  Only the whole fused step (OP_STEP, all phases) runs here.
// Preconditions checked by the dispatcher (mbe.cu): one BS class, width^2+height^2 < 2^24
// (squared distances are exact in FP32), no debug SNR output bound.
#pragma once
#include "mbe_device.cuh"

// Resident CTAs (128 threads) per SM the register allocation is held to, tuned per kernel on B200
// (profiles/README.md): the small central shapes run at full occupancy with 32 registers (10..16 CTAs
// measure the same within noise), the multi-agent and wide shapes need more registers for their
// per-BS arrays (large central 9 CTAs = 56 registers, large multi-agent 8 = 64).
#ifndef MBE_SPEC_MIN_BLOCKS
#define MBE_SPEC_MIN_BLOCKS(MODE, HANDLER, B) \
  ((B) <= 4 ? ((HANDLER) == 1 ? 12 : MBE_SMALL_C_BLOCKS) : (B) <= 10 ? ((MODE) == 0 ? 11 : 10) : ((HANDLER) == 1 ? MBE_LARGE_MA_BLOCKS : MBE_LARGE_C_BLOCKS))
#endif
#ifndef MBE_SMALL_C_BLOCKS
#define MBE_SMALL_C_BLOCKS 16
#endif
#ifndef MBE_LARGE_C_BLOCKS
#define MBE_LARGE_C_BLOCKS 9
#endif
#ifndef MBE_LARGE_MA_BLOCKS
#define MBE_LARGE_MA_BLOCKS 8
#endif

namespace mbe {

template <int HANDLER, int U, int B, bool PER_ENV>
__host__ __device__ constexpr size_t spec_smem_bytes(bool gym) {
  constexpr int EPB = (32 / U) * kWarpsPerBlock;
  constexpr int F = (HANDLER == 1 ? 4 : 2) * B + 1;
  size_t n = gym ? (size_t)EPB * U * F * 4 : 0;
  n = (n + 15) & ~(size_t)15;
  if (PER_ENV) n += (size_t)EPB * B * 4;
  if (gym && HANDLER == 1) n += (size_t)EPB * B * 8;
  return (n + 15) & ~(size_t)15;
}

// Shared / staged memory one chunk of envs (the envs of one CTA iteration) works on.
struct ChunkMem {
  float* obs;      // [EPB*U*F] observation staging block
  uint32_t* bs;    // [EPB*B] per-env BS table (PER_ENV)
  float* bsu;      // [EPB*B] multi-agent BS utilities
  int* bsn;        // [EPB*B] multi-agent BS connection counts
  // state staged in shared memory by bulk async copies (PIPE), else unused
  const uint32_t* st_pos;
  const uint32_t* st_wp;
  const uint32_t* st_conn;
  const int32_t* st_act;
  const int32_t* st_t;
  const int32_t* st_epi;
  const int32_t* st_nbs;
};

// One fused step of the EPB envs starting at env_base.  PIPE: the state comes from shared memory
// (ChunkMem::st_*), otherwise straight from global memory.  The observation block is left in
// m.obs; the caller stores it.
template <int MODE, int HANDLER, int U, int B, bool PER_ENV, bool PIPE, bool MC = false>
__device__ __forceinline__ void step_chunk(const StepArgs& a, const int env_base, const ChunkMem& m) {
  constexpr bool GYM = (MODE == 1);
  constexpr bool MA = (HANDLER == 1);
  constexpr int EPW = 32 / U;
  constexpr int EPB = EPW * kWarpsPerBlock;
  constexpr int F = GYM ? ((MA ? 4 : 2) * B + 1) : 0;
  constexpr unsigned SEG = (U == 32) ? kFull : ((1u << U) - 1u);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  float* s_obs = m.obs;
  uint32_t* s_bs = m.bs;
  float* s_bsu = m.bsu;
  int* s_bsn = m.bsn;

  int seg = lane / U;
  int u = lane - seg * U;
  if (seg >= EPW) {
    seg = EPW;
    u = lane - EPW * U;
  }
  const int env_in_blk = warp * EPW + seg;
  const int env = env_base + env_in_blk;
  const bool valid = (seg < EPW) && (env < a.E);
  const unsigned segmask = valid ? (SEG << (seg * U)) : 0u;
  // idle lanes (32 - EPW*U per warp, and the tail of the last block) run the same code on a
  // clamped index and only their stores are predicated off
  const int env_ld = valid ? env : 0;
  const unsigned idx = (unsigned)env_ld * U + (valid ? u : 0);
  const unsigned gid = a.env_offset + (unsigned)env;
  uint32_t* bs_env = PER_ENV ? s_bs + (size_t)min(env_in_blk, EPB - 1) * B : nullptr;
  // the slot's class constants.  MC: the BSs of a shared layout differ in bw/freq/tx/height
  // (entities.py:6-29) -- every slot has its own folded class, constant-bank operands once the loops
  // over b unroll; otherwise all slots share slot 0's (one register set)
  auto SL = [&](int b) -> const SlotDev& { return a.slot[MC ? b : 0]; };

  // ---- load state (all loads issued before any use) ----
  const unsigned lidx = valid ? (unsigned)(env_in_blk * U + u) : 0u;  // index inside the chunk
  const int lenv = valid ? env_in_blk : 0;
  const uint32_t pos_in = PIPE ? m.st_pos[lidx] : a.pos[idx];
  const uint32_t wp_in = PIPE ? m.st_wp[lidx] : a.wp[idx];
  uint32_t conn = 0;
  int act = 0;
  if (GYM) {
    conn = PIPE ? m.st_conn[lidx] : a.conn[idx];
    act = PIPE ? m.st_act[lidx] : a.actions[idx];
  }
  int t_e = PIPE ? m.st_t[lenv] : a.t[env_ld];
  int epi = PIPE ? m.st_epi[lenv] : a.episode[env_ld];
  int nb = B;
  if (PER_ENV && a.nbs) nb = PIPE ? m.st_nbs[lenv] : a.nbs[env_ld];
  int x, y, wx, wy;
  unpack_xy(pos_in, x, y);
  unpack_xy(wp_in, wx, wy);
  if (!valid) {
    conn = 0;
    act = 0;
  }
  const uint32_t conn_in = conn;
  bool done = false, fresh = false;
  float util = -1.0f;

  auto bs_xy_of = [&](int b, int& bx, int& by) {
    if (PER_ENV) {
      unpack_xy(bs_env[b], bx, by);
    } else {
      bx = a.slot[b].x;
      by = a.slot[b].y;
    }
  };
  auto d2_to = [&](int b) {
    int bx, by;
    bs_xy_of(b, bx, by);
    int dx = x - bx, dy = y - by;
    return dx * dx + dy * dy;
  };

  auto phase_move = [&]() {
    if (wx < 0) next_waypoint(a, gid, (unsigned)u, idx, t_e, epi, valid, wx, wy);  // movement.py:44-47
    if (move_ue(a.mv[0], x, y, wx, wy)) wx = wy = -1;
  };

  auto phase_clock = [&]() {
    t_e += 1;
    done = valid && (t_e >= a.ep_time);
    if (done) conn = 0;  // everyone leaves at ep_time (arrival.py:32-36, base.py:283-285)
    if (valid && u == 0) a.done[env] = done ? 1 : 0;
    if (__any_sync(kFull, done && a.autoreset))
      reinit_env(a, done && a.autoreset, gid, u, idx, env, bs_env, epi, t_e, conn, x, y, wx, wy, nb, fresh);
  };

  if (!GYM) {
    // ================= FORK: move -> associate -> split -> utility (base.py:230-296) =================
    phase_move();
    int best = -1, bestd2 = 0x7fffffff;
#pragma unroll
    for (int b = 0; b < B; ++b) {
      int d2 = d2_to(b);
      bool ok = (d2 <= SL(b).d2max) && (d2 < bestd2);  // strict <: first minimum wins (base.py:240)
      if (PER_ENV) ok = ok && (b < nb);
      if (ok) {
        best = b;
        bestd2 = d2;
      }
    }
    const bool has = valid && best >= 0;
    unsigned peers = __match_any_sync(kFull, has ? (seg * 64 + best) : (0x10000 + lane));
    double rate = 0.0;
    if (has) {
      const SlotDev& sb = SL(best);
      rate = sb.lutn[(unsigned)__popc(peers) * (unsigned)sb.stride + (unsigned)bestd2];
    }
    util = scaled_utility(a, rate);
    if (valid) {
      a.assoc[idx] = best;
      if (a.rate) a.rate[idx] = rate;
      a.utility[idx] = util;
    }
    if (a.metrics) {
      unsigned cm = __ballot_sync(kFull, has) & segmask;
      float usum = seg_sum_head<U>(valid ? util : 0.0f, u);
      float rsum = seg_sum_head<U>((float)rate, u);
      if (valid && u == 0) {
        float nc = (float)__popc(cm);
        reinterpret_cast<float4*>(a.metrics)[env] = make_float4(nc, nc, usum * a.inv_U, mean_or_zero(rsum, nc));
      }
    }
    phase_clock();
  } else {
    // ================= GYM: actions -> split -> utility -> reward -> move -> obs =================
    uint32_t elig = 0;
    int d2pre[B];
#pragma unroll
    for (int b = 0; b < B; ++b) {
      d2pre[b] = d2_to(b);
      bool ok = d2pre[b] <= SL(b).d2max;  // check_connectivity (base.py:212-214)
      if (PER_ENV) ok = ok && (b < nb);
      elig |= ok ? (1u << b) : 0u;
    }
    conn &= elig;  // update_connections (base.py:221-227)
    if (act > 0 && act <= nb) {  // NOOP_ACTION = 0 (base.py:29)
      uint32_t bit = 1u << (act - 1);
      conn = (conn & bit) ? (conn & ~bit) : (conn | (elig & bit));
    }
    // |connections(b)| for every BS of the env, then allocateDataRate2User (base.py:421-435):
    // the link's share comes from the pre-rounded table, bs-major accumulation (413-418); an
    // unconnected slot reads the table's 0.0 entry, so the loads need no predication
    int cnt[B];
    int csum_i = 0;
#pragma unroll
    for (int b = 0; b < B; ++b) {
      if constexpr (EPW == 1)  // the env is the whole warp (idle lanes hold conn = 0): one warp-wide integer add
        cnt[b] = (int)__reduce_add_sync(kFull, (conn >> b) & 1u);
      else
        cnt[b] = __popc(__ballot_sync(kFull, (conn >> b) & 1u) & segmask);
      csum_i += cnt[b];
    }
    double rate = 0.0;
#pragma unroll
    for (int b = 0; b < B; ++b) {
      const unsigned stride = (unsigned)SL(b).stride;
      unsigned off = (unsigned)cnt[b] * stride + (unsigned)d2pre[b];
      rate += SL(b).lutn[((conn >> b) & 1u) ? off : stride - 1u];
    }
    util = scaled_utility(a, rate);
    float usum = seg_sum_head<U>(valid ? util : 0.0f, u);
    if (valid) {
      if (a.rate) a.rate[idx] = rate;
      a.utility[idx] = util;
    }
    if (MA) {
#pragma unroll
      for (int b = 0; b < B; ++b) {
        bool bit = (conn >> b) & 1u;
        float sum = seg_sum_head<U>(bit ? util : 0.0f, u);
        if (valid && u == 0) {
          int n = cnt[b];
          s_bsn[env_in_blk * B + b] = n;
          s_bsu[env_in_blk * B + b] = n ? sum / (float)n : -1.0f;
        }
      }
      __syncwarp();
      if (valid) {
        float nu = 0.0f;
        int ncnt = 0;
#pragma unroll
        for (int b = 0; b < B; ++b)
          if ((elig >> b) & 1u) {  // available_connections (base.py:216-218)
            nu += s_bsu[env_in_blk * B + b];
            ncnt += s_bsn[env_in_blk * B + b];
          }
        a.reward[idx] = (nu + util) / (float)(ncnt + 1);
      }
    } else if (valid && u == 0) {
      a.reward[env] = usum * a.inv_U;  // mean utility (metrics.py:25-28)
    }
    if (a.metrics) {
      unsigned cm = __ballot_sync(kFull, conn != 0) & segmask;
      float csum = (float)csum_i;
      float rsum = seg_sum_head<U>((float)rate, u);
      if (valid && u == 0) {
        float nc = (float)__popc(cm);
        reinterpret_cast<float4*>(a.metrics)[env] = make_float4(csum, nc, usum * a.inv_U, mean_or_zero(rsum, nc));
      }
    }

    phase_move();
    phase_clock();

    // ---- observation of the new state ----
    if (MA) {
      // statistics change only for envs that ended (connections dropped / fresh episode)
      if (__any_sync(kFull, done)) {
        if (done && u == 0) {
#pragma unroll
          for (int b = 0; b < B; ++b) {
            s_bsn[env_in_blk * B + b] = 0;
            s_bsu[env_in_blk * B + b] = -1.0f;
          }
        }
        __syncwarp();
      }
    }
    if (valid) {
      float* row = s_obs + ((unsigned)env_in_blk * U + u) * F;
      if (done && !fresh) {  // inactive UEs observe zeros
#pragma unroll
        for (int f = 0; f < F; ++f) row[f] = 0.0f;
      } else {
        float l[B];
        float lmax = -INFINITY;
        uint32_t elig2 = 0;
        const float xf = (float)x, yf = (float)y;
#pragma unroll
        for (int b = 0; b < B; ++b) {
          float bxf, byf;
          if (PER_ENV) {
            int bx, by;
            unpack_xy(bs_env[b], bx, by);
            bxf = (float)bx;
            byf = (float)by;
          } else {
            bxf = a.slot[b].xf;
            byf = a.slot[b].yf;
          }
          float dx = xf - bxf, dy = yf - byf;
          float d2f = fmaf(dx, dx, dy * dy);  // exact: integers below 2^24
          l[b] = log2_snr_obs_f(SL(b).k, SL(b).l0, d2f);
          bool live = !PER_ENV || (b < nb);
          if (!live) l[b] = -INFINITY;
          lmax = fmaxf(lmax, l[b]);
          if (MA && live && d2f <= (float)SL(b).d2max) elig2 |= 1u << b;
        }
#pragma unroll
        for (int b = 0; b < B; ++b) {
          row[b] = ((conn >> b) & 1u) ? 1.0f : 0.0f;
          row[B + b] = ex2_sfu(l[b] - lmax);  // snr / max snr
        }
        row[2 * B] = fresh ? -1.0f : util;
        if (MA) {
          float cnt[B];
          float tot = 0.0f;
#pragma unroll
          for (int b = 0; b < B; ++b) {
            bool ok = (elig2 >> b) & 1u;
            cnt[b] = ok ? (float)s_bsn[env_in_blk * B + b] : 0.0f;
            row[2 * B + 1 + b] = ok ? s_bsu[env_in_blk * B + b] : -1.0f;
            tot += cnt[b];
          }
          float inv = 1.0f / fmaxf(1.0f, tot);
#pragma unroll
          for (int b = 0; b < B; ++b) row[3 * B + 1 + b] = cnt[b] * inv;
        }
      }
    }
  }

  // ---- store state (only what changed) ----
  if (valid) {
    uint32_t p = pack_xy(x, y), w = pack_xy(wx, wy);
    if (p != pos_in) a.pos[idx] = p;
    if (w != wp_in) a.wp[idx] = w;
    if (GYM && conn != conn_in) a.conn[idx] = conn;
    if (u == 0) {
      a.t[env] = t_e;
    }
  }
}

// ---- one chunk per CTA (works for any E; also the fallback of the pipelined kernel) ----
template <int MODE, int HANDLER, int U, int B, bool PER_ENV, bool MC = false>
__global__ void __launch_bounds__(kThreads, MBE_SPEC_MIN_BLOCKS(MODE, HANDLER, B)) step_spec_kernel(const __grid_constant__ StepArgs a) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  constexpr bool GYM = (MODE == 1);
  constexpr bool MA = (HANDLER == 1);
  constexpr int EPB = (32 / U) * kWarpsPerBlock;
  constexpr int F = GYM ? ((MA ? 4 : 2) * B + 1) : 0;
  // Programmatic dependent launch: let the next launch in the stream become resident while this
  // grid drains, and (as the dependent) wait for the previous grid's memory before touching state.
  // Both are no-ops when the launch does not carry the attribute.
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  prefetch_ahead<EPB, U, GYM>(a);
  asm volatile("griddepcontrol.wait;" ::: "memory");
  constexpr size_t OBS_BYTES = (((size_t)(GYM ? EPB * U * F * 4 : 0)) + 15) & ~(size_t)15;
  constexpr size_t BS_BYTES = PER_ENV ? (size_t)EPB * B * 4 : 0;
  ChunkMem m = {};
  m.obs = reinterpret_cast<float*>(smem_raw);
  m.bs = reinterpret_cast<uint32_t*>(smem_raw + OBS_BYTES);
  m.bsu = reinterpret_cast<float*>(smem_raw + OBS_BYTES + BS_BYTES);
  m.bsn = reinterpret_cast<int*>(smem_raw + OBS_BYTES + BS_BYTES + (size_t)EPB * B * 4);
  const int env_base = blockIdx.x * EPB;
  if (PER_ENV) {
    const int n = min(EPB, a.E - env_base) * B;
    const uint32_t* g = a.bs_xy + (size_t)env_base * B;
    for (int i = threadIdx.x; i < n; i += kThreads) m.bs[i] = g[i];
    __syncthreads();
  }
  step_chunk<MODE, HANDLER, U, B, PER_ENV, false, MC>(a, env_base, m);
  if (GYM) store_obs_block(a, m.obs, env_base, threadIdx.x, true);
}

// the multi-class instantiation of a shape (shared layouts only)
template <int MODE, int HANDLER, int U, int B, bool PER_ENV>
constexpr auto spec_mc_kernel() -> void (*)(StepArgs) {
  if constexpr (PER_ENV) return nullptr;
  else return step_spec_kernel<MODE, HANDLER, U, B, false, true>;
}

// ---- persistent, software-pipelined variant (E must be a multiple of EPB) ----
// Every CTA walks the env chunks with a grid stride.  While chunk k is computed, the state slices
// of chunk k+1 ([EPB,U] words of pos / wp / conn / actions, [EPB] clocks, per-env BS tables) are
// already in flight into the other shared-memory stage as bulk async copies (TMA) that signal an
// mbarrier, and the observation block of chunk k-1 is still draining from its own buffer as a bulk
// async store.  Global-memory latency is thus off the critical path of every chunk but the first.
template <int HANDLER, int U, int B, bool PER_ENV>
__host__ __device__ constexpr size_t pipe_stage_bytes(bool gym) {
  constexpr int EPB = (32 / U) * kWarpsPerBlock;
  size_t n = (size_t)EPB * U * 4 * (gym ? 4 : 2);  // pos, wp (, conn, actions)
  n += (size_t)EPB * 4 * 3;                        // t, episode, nbs
  if (PER_ENV) n += (size_t)EPB * B * 4;
  return (n + 15) & ~(size_t)15;
}
template <int HANDLER, int U, int B, bool PER_ENV>
__host__ __device__ constexpr size_t pipe_smem_bytes(bool gym) {
  constexpr int EPB = (32 / U) * kWarpsPerBlock;
  constexpr int F = (HANDLER == 1 ? 4 : 2) * B + 1;
  size_t obs = gym ? (((size_t)EPB * U * F * 4 + 15) & ~(size_t)15) : 0;
  size_t ma = (gym && HANDLER == 1) ? (size_t)EPB * B * 8 : 0;
  return 2 * obs + ((ma + 15) & ~(size_t)15) + 2 * pipe_stage_bytes<HANDLER, U, B, PER_ENV>(gym) + 16;
}

__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(bar)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = (uint32_t)__cvta_generic_to_shared(bar);
  uint32_t ok = 0;
  while (!ok) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(addr), "r"(parity)
        : "memory");
  }
}
__device__ __forceinline__ void bulk_load(void* sdst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   (uint32_t)__cvta_generic_to_shared(sdst)),
               "l"(gsrc), "r"(bytes), "r"((uint32_t)__cvta_generic_to_shared(bar))
               : "memory");
}

template <int MODE, int HANDLER, int U, int B, bool PER_ENV>
__global__ void __launch_bounds__(kThreads, MBE_SPEC_MIN_BLOCKS(MODE, HANDLER, B)) step_pipe_kernel(const __grid_constant__ StepArgs a) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  constexpr bool GYM = (MODE == 1);
  constexpr bool MA = (HANDLER == 1);
  constexpr int EPB = (32 / U) * kWarpsPerBlock;
  constexpr int F = GYM ? ((MA ? 4 : 2) * B + 1) : 0;
  constexpr size_t OBS_BYTES = GYM ? (((size_t)EPB * U * F * 4 + 15) & ~(size_t)15) : 0;
  constexpr size_t MA_BYTES = (((GYM && MA) ? (size_t)EPB * B * 8 : 0) + 15) & ~(size_t)15;
  constexpr size_t STAGE_BYTES = pipe_stage_bytes<HANDLER, U, B, PER_ENV>(GYM);
  constexpr uint32_t EU = EPB * U * 4;  // bytes of one [EPB,U] word slice
  constexpr uint32_t EE = EPB * 4;      // bytes of one [EPB] word slice
  const int tid = threadIdx.x;

  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  asm volatile("griddepcontrol.wait;" ::: "memory");

  unsigned char* obs_buf[2] = {smem_raw, smem_raw + OBS_BYTES};
  unsigned char* ma_buf = smem_raw + 2 * OBS_BYTES;
  unsigned char* stage[2] = {ma_buf + MA_BYTES, ma_buf + MA_BYTES + STAGE_BYTES};
  uint64_t* bar = reinterpret_cast<uint64_t*>(ma_buf + MA_BYTES + 2 * STAGE_BYTES);

  auto issue_loads = [&](int chunk, int sidx) {  // thread 0 only
    unsigned char* st = stage[sidx];
    const size_t e0 = (size_t)chunk * EPB;
    const uint32_t total = EU * (GYM ? 4 : 2) + EE * 2 + (PER_ENV ? EE + EPB * B * 4 : 0);
    mbar_expect_tx(&bar[sidx], total);
    bulk_load(st, a.pos + e0 * U, EU, &bar[sidx]);
    bulk_load(st + EU, a.wp + e0 * U, EU, &bar[sidx]);
    size_t off = 2 * (size_t)EU;
    if (GYM) {
      bulk_load(st + off, a.conn + e0 * U, EU, &bar[sidx]);
      bulk_load(st + off + EU, a.actions + e0 * U, EU, &bar[sidx]);
      off += 2 * (size_t)EU;
    }
    bulk_load(st + off, a.t + e0, EE, &bar[sidx]);
    bulk_load(st + off + EE, a.episode + e0, EE, &bar[sidx]);
    if (PER_ENV) {
      bulk_load(st + off + 2 * EE, a.nbs + e0, EE, &bar[sidx]);
      bulk_load(st + off + 3 * EE, a.bs_xy + e0 * B, EPB * B * 4, &bar[sidx]);
    }
  };

  const int nchunks = a.E / EPB;
  if (tid == 0) {
    mbar_init(&bar[0], 1);
    mbar_init(&bar[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    if ((int)blockIdx.x < nchunks) issue_loads(blockIdx.x, 0);
  }
  __syncthreads();

  int it = 0;
  for (int chunk = blockIdx.x; chunk < nchunks; chunk += gridDim.x, ++it) {
    const int sidx = it & 1;
    // the other stage was consumed in the previous iteration (its readers passed that iteration's
    // barrier), so it can be refilled now for the next chunk
    if (tid == 0 && chunk + (int)gridDim.x < nchunks) issue_loads(chunk + gridDim.x, sidx ^ 1);
    mbar_wait(&bar[sidx], (uint32_t)((it >> 1) & 1));

    unsigned char* st = stage[sidx];
    ChunkMem m = {};
    m.obs = reinterpret_cast<float*>(obs_buf[sidx]);
    m.bsu = reinterpret_cast<float*>(ma_buf);
    m.bsn = reinterpret_cast<int*>(ma_buf + (size_t)EPB * B * 4);
    m.st_pos = reinterpret_cast<const uint32_t*>(st);
    m.st_wp = reinterpret_cast<const uint32_t*>(st + EU);
    size_t off = 2 * (size_t)EU;
    if (GYM) {
      m.st_conn = reinterpret_cast<const uint32_t*>(st + off);
      m.st_act = reinterpret_cast<const int32_t*>(st + off + EU);
      off += 2 * (size_t)EU;
    }
    m.st_t = reinterpret_cast<const int32_t*>(st + off);
    m.st_epi = reinterpret_cast<const int32_t*>(st + off + EE);
    if (PER_ENV) {
      m.st_nbs = reinterpret_cast<const int32_t*>(st + off + 2 * EE);
      m.bs = reinterpret_cast<uint32_t*>(st + off + 3 * EE);
    }
    step_chunk<MODE, HANDLER, U, B, PER_ENV, true>(a, chunk * EPB, m);

    // the bulk store that last read this observation buffer was issued two iterations ago and
    // waited for one iteration ago (below), so the rows written above raced with nothing
    if (GYM) {
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      if (tid == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");  // previous chunk's store
    }
    __syncthreads();  // rows complete; stage sidx fully consumed
    if (GYM && tid == 0) {
      float* gdst = a.obs + (size_t)chunk * EPB * U * F;
      asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst),
                   "r"((uint32_t)__cvta_generic_to_shared(m.obs)), "r"((uint32_t)(EPB * U * F * 4))
                   : "memory");
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    }
  }
  if (GYM && tid == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}

}  // namespace mbe
