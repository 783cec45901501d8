/* TEST INFRASTRUCTURE ONLY -- compiled CPU restatement (plain C, FP64) of the reference's step.
 *
 * Only tests/ and the cpu_baseline leg of bench.py may load this library (through oracle/c_oracle.py);
 * the product (mobile_env_gan_b200/, libmbe.so) never does.
 *
 * Why it exists: the reference is scalar Python, so the faithful CPU baseline (oracle/mbe_oracle.py
 * ScalarEnv) mostly measures the interpreter.  This file restates the same arithmetic in C, one env
 * per OpenMP iteration, so that bench.py can also report what the host cores do with a COMPILED
 * implementation of the path -- and it is a third, independent restatement pinned to the same
 * golden vectors (tests/test_c_oracle.py: fork_*.json, gymref_*.json from the unmodified reference).
 *
 * Reference lines restated (mobile_env/core/...):
 *   movement.py:42-62      RandomWaypointMovement.move (norm <= velocity snap, np.round half-even)
 *   entities.py:24-26,52-54  int-truncated points; distance = Euclid on ints
 *   channels.py:132-146    OkumuraHata.power_loss;  channels.py:24-27 calculateSNR;  78-83 datarate
 *   base.py:212-214        check_connectivity (snr > threshold, strict)
 *   base.py:236-241        FORK association: nearest connectable BS, first minimum wins
 *   base.py:221-227        update_connections (GYM)
 *   base.py:421-435        allocateDataRate2User: ResourceFair share (schedules.py:20-22), round(.,2);
 *                          ProportionalFair / repaired RateFair as specified in oracle/mbe_oracle.py
 *   base.py:413-418        user_total_datarates (bs-major sum)
 *   utilities.py:44-55     BoundedLogUtility calculate / scale
 *   base.py:438-447        allStationUtilities;  metrics.py:5-28 monitor scalars
 *   base.py:280-291,407-409  clock, leaving UEs, time_is_up
 * GYM stage order / action semantics / observation layout: this build's specification, the same as
 * oracle/mbe_oracle.py ScalarEnv.step_gym (see its header for the parity status).
 *
 * Loop-invariant logarithms are hoisted per BS with the reference's own expressions (same values).
 * Build: gcc -O2 -ffp-contract=off -fopenmp -shared -fPIC (no -ffast-math, no -march: FP64 results
 * must not depend on contraction). */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define MBO_MAX_B 64
#define MBO_MAX_U 1024

typedef struct {
  double width, height, velocity, snr_tr, noise, ue_height;
  double util_lower, util_upper, w1, w2, w3;
  int32_t ep_time;
  int32_t handler;   /* 0 central, 1 multi-agent (GYM observations / reward) */
  int32_t scheduler; /* 0 ResourceFair (schedules.py:20-22); 1 ProportionalFair, 2 repaired RateFair: this
                        build's own specifications (oracle/mbe_oracle.py pf_total / rate_fair_share) */
  int32_t pad_;
} mbo_params;

typedef struct {
  double bw, tx, tmp1, tmp2;
} bs_fold;

/* OkumuraHata constants of one BS, channels.py:137-144 (operation order of the Python source) */
static void fold_bs(const mbo_params* p, const double* par /* bw, freq, tx, height */, bs_fold* f) {
  const double freq = par[1], hb = par[3];
  const double ch = 0.8 + (1.1 * log10(freq) - 0.7) * p->ue_height - 1.56 * log10(freq);
  f->tmp1 = 69.55 - ch + 26.16 * log10(freq) - 13.82 * log10(hb);
  f->tmp2 = 44.9 - 6.55 * log10(hb);
  f->bw = par[0];
  f->tx = par[2];
}

/* Optional PINNED tables (mbo_set_tables): snr and Shannon rate per BS slot (or one shared table) and
 * integer squared distance, computed by the caller with numpy in the reference's own operation order
 * (oracle/mbe_oracle.py snr_of / datarate_of, which the golden vectors pin bit for bit).  glibc's
 * log10 / pow / log2 may differ from numpy's in the last ulp; with the tables every rate this file
 * produces is the numpy value exactly, so full-size comparisons can demand bit equality.  Without
 * them the chain below runs on glibc (the independent restatement tests/test_c_oracle.py pins). */
static const double* g_snr_tab = NULL;
static const double* g_rate_tab = NULL;
static int g_ntabs = 0, g_tablen = 0;
void mbo_set_tables(const double* snr_tab, const double* rate_tab, int n_tabs, int tab_len) {
  g_snr_tab = snr_tab;
  g_rate_tab = rate_tab;
  g_ntabs = n_tabs;
  g_tablen = tab_len;
}
static inline size_t tab_index(int b, int d2) { return (size_t)(g_ntabs == 1 ? 0 : b) * (size_t)g_tablen + (size_t)d2; }

static inline double snr_at(const mbo_params* p, const bs_fold* f, int b, int dx, int dy) {
  const int d2 = dx * dx + dy * dy;
  if (g_snr_tab && d2 < g_tablen) return g_snr_tab[tab_index(b, d2)];
  const double dist = sqrt((double)d2);                         /* bs.point.distance(ue.point) */
  const double loss = f->tmp1 + f->tmp2 * log10(dist + 1e-16);  /* channels.py:146 */
  return pow(10.0, (f->tx - loss) / 10.0) / p->noise;           /* channels.py:26-27 */
}

static inline double datarate(const mbo_params* p, const bs_fold* f, int b, int d2, double snr) {
  if (g_rate_tab && d2 < g_tablen) return g_rate_tab[tab_index(b, d2)];
  return snr > p->snr_tr ? f->bw * log2(1.0 + snr) : 0.0; /* channels.py:80-83 */
}

static inline double round2(double v) { return rint(v * 100.0) / 100.0; } /* np.float64.__round__(2) */

/* Order-independent per-BS totals of the two extra schedulers (fixed point, exact integer sums):
 * ProportionalFair  total = sum(rint(r * 2^20)) * 2^-20,   share_u = r_u * r_u / total
 * RateFair          total = sum(rint(2^50 / r)) * 2^-50,   share   = 1 / total */
typedef unsigned __int128 fix_t;
static inline fix_t sched_term(int scheduler, double r) {
  return scheduler == 1 ? (fix_t)rint(r * 1048576.0) : scheduler == 2 ? (fix_t)rint(0x1p50 / r) : (fix_t)0;
}
static inline double link_share(int scheduler, double r, int n, fix_t tot) {
  if (scheduler == 1) return (r * r) / ((double)tot * (1.0 / 1048576.0));
  if (scheduler == 2) return 1.0 / ((double)tot * 0x1p-50);
  return r / (double)n;
}

static inline double scaled_utility(const mbo_params* p, double rate) {
  double u = p->util_lower;
  if (rate > 0.0) {
    u = p->w1 * log(p->w2 + rate) / log(p->w3);
    if (u < p->util_lower) u = p->util_lower;
    if (u > p->util_upper) u = p->util_upper;
  }
  return 2.0 * (u - p->util_lower) / (p->util_upper - p->util_lower) - 1.0;
}

/* movement.py:42-62 for one UE whose waypoint exists; returns 1 when it arrived (waypoint popped) */
static inline int move_ue(const mbo_params* p, int32_t* x, int32_t* y, int wx, int wy) {
  const int dx = wx - *x, dy = wy - *y;
  const double norm = sqrt((double)(dx * dx + dy * dy));
  if (norm <= p->velocity) {
    *x = wx;
    *y = wy;
    return 1;
  }
  *x = (int32_t)rint((double)*x + (p->velocity * (double)dx) / norm);
  *y = (int32_t)rint((double)*y + (p->velocity * (double)dy) / norm);
  return 0;
}

static void move_all(const mbo_params* p, int U, int32_t* pos, int32_t* wp, const int32_t* new_wp, int32_t* drew) {
  for (int u = 0; u < U; ++u) {
    int d = 0;
    if (wp[2 * u] < 0) { /* movement.py:44-47: draw (here: take the supplied draw) */
      wp[2 * u] = new_wp[2 * u];
      wp[2 * u + 1] = new_wp[2 * u + 1];
      d = 1;
    }
    if (move_ue(p, &pos[2 * u], &pos[2 * u + 1], wp[2 * u], wp[2 * u + 1])) wp[2 * u] = wp[2 * u + 1] = -1;
    if (drew) drew[u] = d;
  }
}

/* One FORK step (base.py:230-296) of E independent envs.
 * bs_par [B,4] = bw, freq, tx, height per BS slot; bs_xy [B,2] shared or [E,B,2] (+ nbs [E]) per env;
 * pos / wp [E,U,2] (wp x < 0: none) updated in place; new_wp [E,U,2]: the draw a UE without a waypoint
 * takes; drew [E,U] (may be NULL) tells which draws were consumed; t [E] clock, updated.
 * Outputs: assoc [E,U] (-1 none), rate [E,U], util [E,U], done [E], metrics [E,4] =
 * (#connections, #connected, mean utility, mean datarate). */
void mbo_fork_step(const mbo_params* p, int E, int U, int B, const double* bs_par, const int32_t* bs_xy,
                   int bs_per_env, const int32_t* nbs, int32_t* pos, int32_t* wp, const int32_t* new_wp,
                   int32_t* drew, int32_t* t, int32_t* assoc, double* rate, double* util, uint8_t* done,
                   double* metrics) {
  bs_fold fold[MBO_MAX_B];
  for (int b = 0; b < B; ++b) fold_bs(p, bs_par + 4 * b, &fold[b]);
#pragma omp parallel for schedule(static)
  for (int e = 0; e < E; ++e) {
    const int32_t* bs = bs_per_env ? bs_xy + (size_t)e * B * 2 : bs_xy;
    const int nb = nbs ? nbs[e] : B;
    int32_t* ps = pos + (size_t)e * U * 2;
    int32_t* as = assoc + (size_t)e * U;
    double* rt = rate + (size_t)e * U;
    double* ut = util + (size_t)e * U;
    double best_snr[MBO_MAX_U];
    int best_d2[MBO_MAX_U];
    int cnt[MBO_MAX_B];
    memset(cnt, 0, sizeof(int) * (size_t)B);
    move_all(p, U, ps, wp + (size_t)e * U * 2, new_wp + (size_t)e * U * 2, drew ? drew + (size_t)e * U : NULL);
    for (int u = 0; u < U; ++u) { /* base.py:236-241 */
      int best = -1, bestd2 = 0;
      for (int b = 0; b < nb; ++b) {
        const int dx = ps[2 * u] - bs[2 * b], dy = ps[2 * u + 1] - bs[2 * b + 1];
        const int d2 = dx * dx + dy * dy;
        if (best >= 0 && d2 >= bestd2) continue; /* not nearer than the current choice: first minimum wins */
        const double snr = snr_at(p, &fold[b], b, dx, dy);
        if (snr > p->snr_tr) { /* check_connectivity, base.py:212-214 */
          best = b;
          bestd2 = d2;
          best_snr[u] = snr;
          best_d2[u] = d2;
        }
      }
      as[u] = best;
      if (best >= 0) cnt[best] += 1;
    }
    int nconn = 0;
    double usum = 0.0, rsum = 0.0;
    fix_t tot[MBO_MAX_B];
    if (p->scheduler) {
      memset(tot, 0, sizeof(fix_t) * (size_t)B);
      for (int u = 0; u < U; ++u)
        if (as[u] >= 0) tot[as[u]] += sched_term(p->scheduler, datarate(p, &fold[as[u]], as[u], best_d2[u], best_snr[u]));
    }
    for (int u = 0; u < U; ++u) { /* base.py:421-435, 413-418, 253-258 */
      double r = 0.0;
      if (as[u] >= 0) {
        r = 0.0 + round2(link_share(p->scheduler, datarate(p, &fold[as[u]], as[u], best_d2[u], best_snr[u]), cnt[as[u]],
                                    p->scheduler ? tot[as[u]] : (fix_t)0));
        nconn += 1;
        rsum += r;
      }
      rt[u] = r;
      ut[u] = scaled_utility(p, r);
      usum += ut[u];
    }
    if (metrics) {
      double* m = metrics + (size_t)e * 4;
      m[0] = nconn;
      m[1] = nconn;
      m[2] = usum / U;
      m[3] = nconn ? rsum / nconn : 0.0;
    }
    t[e] += 1;
    done[e] = t[e] >= p->ep_time; /* base.py:407-409 */
  }
}

/* One GYM step (order of oracle/mbe_oracle.py ScalarEnv.step_gym) of E independent envs, shared BS
 * layout.  conn [E,U,B] bytes (0/1) in place; actions [E,U] (0 = NOOP, a > 0 toggles BS a-1).
 * Outputs: rate / util [E,U], reward [E] (central) or [E,U] (multi-agent), done [E], obs f32 [E,U,F]
 * with F = 2B+1 (central) / 4B+1 (multi-agent), bs_util [E,B] (may be NULL), metrics [E,4]. */
void mbo_gym_step(const mbo_params* p, int E, int U, int B, const double* bs_par, const int32_t* bs_xy,
                  int32_t* pos, int32_t* wp, const int32_t* new_wp, int32_t* drew, int32_t* t, uint8_t* conn,
                  const int32_t* actions, double* rate, double* util, double* reward, uint8_t* done, float* obs,
                  double* bs_util_out, double* metrics) {
  bs_fold fold[MBO_MAX_B];
  for (int b = 0; b < B; ++b) fold_bs(p, bs_par + 4 * b, &fold[b]);
  const int ma = p->handler == 1;
  const int F = (ma ? 4 : 2) * B + 1;
  const double idle = 2.0 * (p->util_lower - p->util_lower) / (p->util_upper - p->util_lower) - 1.0;
#pragma omp parallel for schedule(static)
  for (int e = 0; e < E; ++e) {
    int32_t* ps = pos + (size_t)e * U * 2;
    uint8_t* cn = conn + (size_t)e * U * B;
    const int32_t* ac = actions + (size_t)e * U;
    double* rt = rate + (size_t)e * U;
    double* ut = util + (size_t)e * U;
    double* snr = (double*)malloc(sizeof(double) * (size_t)U * B);
    int* d2s = (int*)malloc(sizeof(int) * (size_t)U * B);
    uint8_t* ok = (uint8_t*)malloc((size_t)U * B);
    int cnt[MBO_MAX_B];
    double bsu[MBO_MAX_B];
    for (int u = 0; u < U; ++u)
      for (int b = 0; b < B; ++b) {
        const int dx = ps[2 * u] - bs_xy[2 * b], dy = ps[2 * u + 1] - bs_xy[2 * b + 1];
        const double s = snr_at(p, &fold[b], b, dx, dy);
        snr[u * B + b] = s;
        d2s[u * B + b] = dx * dx + dy * dy;
        ok[u * B + b] = s > p->snr_tr;
      }
    /* (1) update_connections (base.py:221-227), (2) actions */
    for (int u = 0; u < U; ++u) {
      for (int b = 0; b < B; ++b) cn[u * B + b] = cn[u * B + b] && ok[u * B + b];
      const int a = ac[u];
      if (a > 0 && a <= B) {
        const int b = a - 1;
        if (cn[u * B + b]) cn[u * B + b] = 0;
        else if (ok[u * B + b]) cn[u * B + b] = 1;
      }
    }
    /* (3) allocation: ResourceFair + rounding per link, bs-major sum per UE; (4) utility */
    int nlinks = 0, nconn = 0;
    for (int b = 0; b < B; ++b) {
      cnt[b] = 0;
      for (int u = 0; u < U; ++u) cnt[b] += cn[u * B + b];
      nlinks += cnt[b];
    }
    double usum = 0.0, rsum = 0.0;
    fix_t tot[MBO_MAX_B];
    for (int b = 0; b < B; ++b) {
      tot[b] = 0;
      if (p->scheduler)
        for (int u = 0; u < U; ++u)
          if (cn[u * B + b]) tot[b] += sched_term(p->scheduler, datarate(p, &fold[b], b, d2s[u * B + b], snr[u * B + b]));
    }
    for (int u = 0; u < U; ++u) {
      double r = 0.0;
      int any = 0;
      for (int b = 0; b < B; ++b)
        if (cn[u * B + b]) {
          r += round2(link_share(p->scheduler, datarate(p, &fold[b], b, d2s[u * B + b], snr[u * B + b]), cnt[b], tot[b]));
          any = 1;
        }
      rt[u] = r;
      ut[u] = scaled_utility(p, r);
      usum += ut[u];
      if (any) {
        nconn += 1;
        rsum += r;
      }
    }
    for (int b = 0; b < B; ++b) { /* allStationUtilities, base.py:438-447 */
      double s = 0.0;
      for (int u = 0; u < U; ++u)
        if (cn[u * B + b]) s += ut[u];
      bsu[b] = cnt[b] ? s / cnt[b] : idle;
      if (bs_util_out) bs_util_out[(size_t)e * B + b] = bsu[b];
    }
    /* (5) reward */
    if (!ma) {
      reward[e] = usum / U; /* metrics.py:25-28 */
    } else {
      for (int u = 0; u < U; ++u) {
        double nu = 0.0;
        int nc = 0;
        for (int b = 0; b < B; ++b)
          if (ok[u * B + b]) {
            nu += bsu[b];
            nc += cnt[b];
          }
        reward[(size_t)e * U + u] = (nu + ut[u]) / (double)(nc + 1);
      }
    }
    if (metrics) {
      double* m = metrics + (size_t)e * 4;
      m[0] = nlinks;
      m[1] = nconn;
      m[2] = usum / U;
      m[3] = nconn ? rsum / nconn : 0.0;
    }
    /* (6) move, clock, departures (base.py:232-233, 280-291) */
    move_all(p, U, ps, wp + (size_t)e * U * 2, new_wp + (size_t)e * U * 2, drew ? drew + (size_t)e * U : NULL);
    t[e] += 1;
    done[e] = t[e] >= p->ep_time;
    float* ob = obs + (size_t)e * U * F;
    if (done[e]) {
      memset(cn, 0, (size_t)U * B);
      memset(ob, 0, sizeof(float) * (size_t)U * F);
    } else {
      for (int u = 0; u < U; ++u) {
        float* row = ob + (size_t)u * F;
        double s[MBO_MAX_B], mx = 0.0, tot = 0.0;
        uint8_t k[MBO_MAX_B];
        for (int b = 0; b < B; ++b) {
          s[b] = snr_at(p, &fold[b], b, ps[2 * u] - bs_xy[2 * b], ps[2 * u + 1] - bs_xy[2 * b + 1]);
          k[b] = s[b] > p->snr_tr;
          if (b == 0 || s[b] > mx) mx = s[b];
          if (k[b]) tot += (double)cnt[b];
        }
        if (tot < 1.0) tot = 1.0;
        for (int b = 0; b < B; ++b) {
          row[b] = cn[u * B + b] ? 1.0f : 0.0f;
          row[B + b] = (float)(s[b] / mx);
          if (ma) {
            row[2 * B + 1 + b] = (float)(k[b] ? bsu[b] : idle);
            row[3 * B + 1 + b] = (float)(k[b] ? (double)cnt[b] / tot : 0.0);
          }
        }
        row[2 * B] = (float)ut[u];
      }
    }
    free(snr);
    free(d2s);
    free(ok);
  }
}

int mbo_max_b(void) { return MBO_MAX_B; }
int mbo_max_u(void) { return MBO_MAX_U; }
