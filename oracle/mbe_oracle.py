"""TEST INFRASTRUCTURE ONLY -- CPU restatement (numpy FP64) of the reference's step.

Only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` may import this module.  The product path
(``mobile_env_gan_b200``) never does: it fails loudly when the CUDA library is absent.

PARITY STATUS
  * FORK mode (the reference's real ``MComCore.step``): **pinned**.  ``ScalarEnv`` is
    checked step-by-step against the unmodified reference run in the build container
    (``oracle/ref_harness.py``) and against the notebook known-answer vectors KAT-1/2/3
    (reference ``mobile_env/GNN/GNN.ipynb`` cell 3 / cell 17 outputs), see
    ``tests/test_oracle_golden.py`` and ``tests/golden/``.
  * GYM mode (actions, central / multi-agent observations and rewards): **parity
    unpinned** -- the reference fork has no handler package, no action argument and no
    observation (``core/base.py:230,296``).  The step order and feature definitions
    below are this build's specification of record, assembled
    from the pieces that do survive in the fork: ``update_connections``
    (base.py:221-227), ``check_connectivity`` (212-214), ``available_connections``
    (216-218), ``allStationUtilities`` (438-447), ``NOOP_ACTION`` (29),
    ``metrics.mean_utility`` (metrics.py:25-28).  The arithmetic of those stages IS pinned:
    ``ref_harness.record_gym_pieces_episode`` runs GYM-order episodes with the reference's own
    primitives (incl. UEs on several BSs through ``allocateDataRate2User`` /
    ``user_total_datarates``) and ``step_gym`` must replay ``tests/golden/gymref_*.json``; the
    stage order, the action semantics, the observation layout and the multi-agent reward remain
    this build's own specification.

Two forms are provided:
  ``ScalarEnv``   per-entity Python loops, op-for-op like the reference (slow; also the
                  ``cpu_baseline`` of bench.py because that is how the reference runs);
  ``batch_*``     the same arithmetic vectorised over [E,U,B] for full-size checks.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Optional, Sequence

import numpy as np

EPSILON = 1e-16  # reference core/channels.py:8
NOOP_ACTION = 0  # reference core/base.py:29

PURPOSE_WAYPOINT, PURPOSE_INITPOS, PURPOSE_BSLAYOUT = 0, 1, 2


# --------------------------------------------------------------------------------------
# Parameters (defaults = reference core/base.py:102-137)
# --------------------------------------------------------------------------------------
@dataclass
class Params:
    width: float = 200.0
    height: float = 200.0
    ep_time: int = 20  # min(EP_MAX_TIME, max_departure), base.py:105,126,407-409
    bw: float = 9e6
    freq: float = 2500.0
    tx: float = 40.0
    bs_height: float = 50.0
    velocity: float = 1.5
    snr_tr: float = 2e-8
    noise: float = 1e-9
    ue_height: float = 1.6
    util_lower: float = -20.0
    util_upper: float = 20.0
    util_coeffs: Sequence[float] = (10.0, 0.0, 10.0)
    scheduler: str = "resource_fair"  # or "proportional_fair" / "rate_fair" (own specs below)
    # None: OkumuraHata (channels.py:131-146).  Custom Channel subclasses of the reference that only
    # override power_loss (channels.py:18-21), as the fixtures' generator defines them:
    #   ("pathloss", gamma)                      the README's example (README.md:108-121)
    #   ("two_slope", gamma1, gamma2, d_break)   a loss that is NOT affine in log-distance
    channel: Optional[Sequence] = None


# --------------------------------------------------------------------------------------
# Scalar chain, op-for-op (used for the LUT-free oracle values)
# --------------------------------------------------------------------------------------
def power_loss(p: Params, dist: float) -> float:
    """OkumuraHata.power_loss, reference core/channels.py:132-146 -- or one of the custom channels
    (subclasses overriding power_loss only) the fixture generator runs through the reference."""
    if p.channel is not None:
        kind = p.channel[0]
        with np.errstate(divide="ignore"):
            if kind == "pathloss":  # README.md:108-121: 10 * gamma * log10(4 pi d f)
                return 10 * p.channel[1] * np.log10(4 * np.pi * dist * p.freq)
            if kind == "two_slope":  # oracle/gen_golden.py:TwoSlope
                g1, g2, brk = p.channel[1:4]
                if dist <= brk:
                    return 10 * g1 * np.log10(4 * np.pi * dist * p.freq)
                return 10 * g1 * np.log10(4 * np.pi * brk * p.freq) + 10 * g2 * np.log10(dist / brk)
        raise ValueError(f"unknown channel {p.channel!r}")
    ch = 0.8 + (1.1 * np.log10(p.freq) - 0.7) * p.ue_height - 1.56 * np.log10(p.freq)
    tmp_1 = 69.55 - ch + 26.16 * np.log10(p.freq) - 13.82 * np.log10(p.bs_height)
    tmp_2 = 44.9 - 6.55 * np.log10(p.bs_height)
    return tmp_1 + tmp_2 * np.log10(dist + EPSILON)


def snr_of(p: Params, dist: float) -> float:
    """Channel.calculateSNR, reference core/channels.py:24-27."""
    loss = power_loss(p, dist)
    with np.errstate(over="ignore"):
        power = 10 ** ((p.tx - loss) / 10)
    return power / p.noise


def datarate_of(p: Params, snr: float) -> float:
    """Channel.datarate, reference core/channels.py:78-83."""
    if snr > p.snr_tr:
        return p.bw * np.log2(1 + snr)
    return 0.0


def utility_of(p: Params, rate: float) -> float:
    """BoundedLogUtility.calculateUtility, reference core/utilities.py:44-52."""
    w1, w2, w3 = p.util_coeffs
    if rate <= 0.0:
        return p.util_lower
    return float(np.clip(w1 * np.log(w2 + rate) / np.log(w3), p.util_lower, p.util_upper))


def scale_utility(p: Params, u: float) -> float:
    """BoundedLogUtility.scaleUtility, reference core/utilities.py:54-55."""
    return 2 * (u - p.util_lower) / (p.util_upper - p.util_lower) - 1


PF_SCALE = 2.0**20


def pf_total(rates) -> float:
    """ProportionalFair denominator.  The fork has no ProportionalFair (core/schedules.py holds
    ResourceFair and a broken RateFair only), so this is the build's specification (parity
    unpinned): every UE of a BS gets the fraction r_u / sum(r) of the BS's resources, i.e.
    ``share_u = r_u * r_u / total``.  To make ``total`` independent of the summation order (the
    reference iterates a Python set) it is accumulated in 2^-20 fixed point:
    ``total = float(sum(int(rint(r * 2^20)))) * 2^-20``."""
    return float(sum(int(np.rint(np.float64(r) * PF_SCALE)) for r in rates)) * (1.0 / PF_SCALE)


RF_SCALE = 2.0**50


def rate_fair_share(rates) -> float:
    """RateFair: every UE of the BS receives the same rate 1 / sum(1 / r_i).  The fork's RateFair
    (core/schedules.py:26-29) computes exactly this scalar but returns it instead of a list, so
    allocateDataRate2User (base.py:435) cannot use it; this is the repaired form (parity
    unpinned).  The sum of inverse rates is accumulated in 2^-50 fixed point so that it does not
    depend on the summation order: total = float(sum(int(rint(2^50 / r)))) * 2^-50."""
    tot = float(sum(int(np.rint(RF_SCALE / np.float64(r))) for r in rates)) * (1.0 / RF_SCALE)
    return 1.0 / tot


def int_point_dist(ax, ay, bx, by) -> float:
    """bs.point.distance(ue.point): both truncated to int (entities.py:24-26,52-54)."""
    return math.hypot(int(ax) - int(bx), int(ay) - int(by))


def move_one(pos, wp, velocity):
    """RandomWaypointMovement.move after the waypoint exists (movement.py:49-62).
    Returns (new_pos, arrived)."""
    position = np.array([pos[0], pos[1]])
    waypoint = np.array([wp[0], wp[1]])
    if np.linalg.norm(position - waypoint) <= velocity:
        return (int(wp[0]), int(wp[1])), True
    v = waypoint - position
    position = position + velocity * v / np.linalg.norm(v)
    position = np.round(position).astype(int)
    return (int(position[0]), int(position[1])), False


# --------------------------------------------------------------------------------------
# Philox4x32-10 (Salmon et al., SC'11) -- the counter-based generator of the CUDA path
# --------------------------------------------------------------------------------------
_M0, _M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
_W0, _W1 = np.uint32(0x9E3779B9), np.uint32(0xBB67AE85)


def philox4x32(c0, c1, c2, c3, k0, k1):
    """Vectorised Philox4x32-10; all inputs broadcastable uint32 arrays. Returns 4 uint32 arrays."""
    c0, c1, c2, c3 = (np.asarray(c, dtype=np.uint32) for c in (c0, c1, c2, c3))
    c0, c1, c2, c3 = np.broadcast_arrays(c0, c1, c2, c3)
    k0 = np.uint32(k0)
    k1 = np.uint32(k1)
    mask = np.uint64(0xFFFFFFFF)
    with np.errstate(over="ignore"):
        for _ in range(10):
            p0 = c0.astype(np.uint64) * _M0
            p1 = c2.astype(np.uint64) * _M1
            hi0, lo0 = (p0 >> np.uint64(32)).astype(np.uint32), (p0 & mask).astype(np.uint32)
            hi1, lo1 = (p1 >> np.uint64(32)).astype(np.uint32), (p1 & mask).astype(np.uint32)
            c0, c1, c2, c3 = hi1 ^ c1 ^ k0, lo1, hi0 ^ c3 ^ k1, lo0
            k0 = np.uint32((int(k0) + int(_W0)) & 0xFFFFFFFF)
            k1 = np.uint32((int(k1) + int(_W1)) & 0xFFFFFFFF)
    return c0, c1, c2, c3


def philox_point(seed: int, env_gid, ue, t, purpose: int, salt, width: float, height: float):
    """Uniform integer point: x = int(u0*W), y = int(u1*H), u = r * 2**-32 in FP64
    (the counter-based analogue of ``int(rng.uniform(0, W))``, movement.py:45-46,69-70).
    Counter = (env_gid, ue, t, purpose + 4*salt); key = (seed & 0xffffffff, seed >> 32)."""
    c3 = (np.asarray(salt, dtype=np.uint64) * np.uint64(4) + np.uint64(purpose)).astype(np.uint32)
    r0, r1, _, _ = philox4x32(env_gid, ue, t, c3, seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
    x = (r0.astype(np.float64) * 2.0**-32 * float(width)).astype(np.int64)
    y = (r1.astype(np.float64) * 2.0**-32 * float(height)).astype(np.int64)
    return x, y


def philox_bs_count(seed: int, env_gid, salt, nmin: int, nmax: int):
    """Number of BSs for the random-layout scenario (custom.py:70 ``random.randint(5,10)``):
    nmin + floor(u * (nmax-nmin+1)), drawn with ue index 0xFFFF and t = 0xFFFF."""
    c3 = (np.asarray(salt, dtype=np.uint64) * np.uint64(4) + np.uint64(PURPOSE_BSLAYOUT)).astype(np.uint32)
    r0, _, _, _ = philox4x32(env_gid, 0xFFFF, 0xFFFF, c3, seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
    return nmin + (r0.astype(np.float64) * 2.0**-32 * float(nmax - nmin + 1)).astype(np.int64)


# --------------------------------------------------------------------------------------
# ScalarEnv -- per-entity loops like the reference
# --------------------------------------------------------------------------------------
@dataclass
class ScalarEnv:
    """One environment, FORK and GYM steps.  Waypoints come from ``wp_source(ue, k)``
    (k = how many this UE has drawn so far) so that reference trajectories can be replayed."""

    p: Params
    bs_xy: list
    num_ues: int
    wp_source: Optional[object] = None
    bs_over: Optional[list] = None  # per-BS overrides of bw / freq / tx / bs_height (entities.py:6-22)
    ue_over: Optional[list] = None  # per-UE overrides of velocity / snr_tr / noise / ue_height (entities.py:32-57)
    pos: list = field(default_factory=list)
    wp: list = field(default_factory=list)
    wp_count: list = field(default_factory=list)
    conn: list = field(default_factory=list)  # GYM: set of bs per ue
    utilities: dict = field(default_factory=dict)
    t: int = 0

    def reset(self, init_pos):
        self.t = 0
        self.pos = [(int(x), int(y)) for x, y in init_pos]
        self.wp = [None] * self.num_ues
        self.wp_count = [0] * self.num_ues
        self.conn = [set() for _ in range(self.num_ues)]
        self.utilities = {}

    # -- stages ------------------------------------------------------------------------
    def _move_all(self):
        for u in range(self.num_ues):  # base.py:232-233
            if self.wp[u] is None:  # movement.py:44-47
                self.wp[u] = tuple(int(v) for v in self.wp_source(u, self.wp_count[u]))
                self.wp_count[u] += 1
            new, arrived = move_one(self.pos[u], self.wp[u], self.p_of(None, u).velocity)
            if arrived:
                self.wp[u] = None
            self.pos[u] = new

    def p_of(self, b, u=None) -> Params:
        """Parameters seen by the link BS b -- UE u (the reference keeps bw/freq/tx/height per BS and
        velocity/snr_threshold/noise/height per UE, entities.py:6-57)."""
        over = {}
        if b is not None and self.bs_over and self.bs_over[b]:
            over.update(self.bs_over[b])
        if u is not None and self.ue_over and self.ue_over[u]:
            over.update(self.ue_over[u])
        if not over:
            return self.p
        import dataclasses

        return dataclasses.replace(self.p, **over)

    def snr(self, b, u):
        bx, by = self.bs_xy[b]
        return snr_of(self.p_of(b, u), int_point_dist(bx, by, *self.pos[u]))

    def connectable(self, b, u):  # base.py:212-214
        return self.snr(b, u) > self.p_of(b, u).snr_tr

    def _allocate(self, bs_conns):
        """allocateDataRate2User for every BS (base.py:421-435) + user_total_datarates
        (413-418).  bs_conns[b] = list of ues.  Returns pair rates and per-UE totals."""
        pair = {}
        for b, ues in enumerate(bs_conns):
            snrs = [self.snr(b, u) for u in ues]
            max_alloc = [datarate_of(self.p_of(b, u), s) for u, s in zip(ues, snrs)]
            if self.p.scheduler == "proportional_fair":
                tot = pf_total(max_alloc)
                rates = [np.float64(r) * np.float64(r) / tot for r in max_alloc]
            elif self.p.scheduler == "rate_fair" and max_alloc:
                rates = [np.float64(rate_fair_share(max_alloc))] * len(max_alloc)
            else:
                rates = [r / len(max_alloc) for r in max_alloc]  # ResourceFair, schedules.py:20-22
            for u, r in zip(ues, rates):
                pair[(b, u)] = round(np.float64(r), 2)  # np.float64.__round__
        total = {}
        for (b, u), r in pair.items():  # bs-major insertion order
            total[u] = total.get(u, 0) + r
        return pair, total

    def _utilities(self, total):
        return [
            float(scale_utility(self.p, utility_of(self.p, total.get(u, 0.0))))
            for u in range(self.num_ues)
        ]

    # -- FORK step: move -> associate -> allocate -> utility (base.py:230-296) ----------
    def step_fork(self):
        B = len(self.bs_xy)
        self._move_all()
        assoc = [-1] * self.num_ues
        bs_conns = [[] for _ in range(B)]
        for u in range(self.num_ues):  # base.py:236-241
            avail = [b for b in range(B) if self.connectable(b, u)]
            if avail:
                ux, uy = self.pos[u]
                closest = min(
                    avail,
                    key=lambda b: np.linalg.norm(
                        [ux - int(self.bs_xy[b][0]), uy - int(self.bs_xy[b][1])]
                    ),
                )
                assoc[u] = closest
                bs_conns[closest].append(u)
        pair, total = self._allocate(bs_conns)
        util = self._utilities(total)
        rates = [float(total.get(u, 0.0)) for u in range(self.num_ues)]
        self.utilities = dict(enumerate(util))
        self.t += 1
        return {
            "pos": list(self.pos),
            "assoc": assoc,
            "pair_rates": pair,
            "rate": rates,
            "utility": util,
            "n_connections": sum(len(c) for c in bs_conns),  # metrics.py:5-9
            "n_connected": sum(1 for a in assoc if a >= 0),  # metrics.py:13-14
            "mean_utility": float(np.mean(util)) if util else self.p.util_lower,  # metrics.py:25-28
            "mean_datarate": float(np.mean(list(total.values()))) if total else 0.0,  # metrics.py:18-21
            "done": self.t >= self.p.ep_time,  # base.py:407-409
        }

    # -- GYM step (spec of record, parity unpinned) ------------------------------------
    def _bs_utilities(self, bs_conns):
        idle = scale_utility(self.p, self.p.util_lower)  # base.py:438-447
        out = []
        for ues in bs_conns:
            if ues:
                out.append(sum(self.utilities[u] for u in ues) / len(ues))
            else:
                out.append(idle)
        return out

    def _bs_conns(self):
        B = len(self.bs_xy)
        return [[u for u in range(self.num_ues) if b in self.conn[u]] for b in range(B)]

    def observe(self, handler="central", active=True):
        """Per-UE features.  central: [conn onehot(B), snr/max snr (B), utility(1)];
        ma adds [bcast(B), stations_connected(B)].  Inactive UEs -> zeros."""
        B = len(self.bs_xy)
        F = 2 * B + 1 if handler == "central" else 4 * B + 1
        obs = np.zeros((self.num_ues, F), dtype=np.float32)
        if not active:
            return obs
        idle = scale_utility(self.p, self.p.util_lower)
        bs_conns = self._bs_conns()
        bs_util = self._bs_utilities(bs_conns) if self.utilities else [idle] * B
        for u in range(self.num_ues):
            onehot = [1.0 if b in self.conn[u] else 0.0 for b in range(B)]
            snrs = [self.snr(b, u) for b in range(B)]
            mx = max(snrs)
            snrs = [s / mx for s in snrs]
            util = self.utilities.get(u, idle)
            row = onehot + snrs + [util]
            if handler != "central":
                ok = [self.connectable(b, u) for b in range(B)]
                bcast = [bs_util[b] if ok[b] else idle for b in range(B)]
                cnt = [float(len(bs_conns[b])) if ok[b] else 0.0 for b in range(B)]
                tot = max(1, sum(cnt))
                row = row + bcast + [c / tot for c in cnt]
            obs[u] = np.asarray(row, dtype=np.float32)
        return obs

    def step_gym(self, actions, handler="central"):
        B = len(self.bs_xy)
        # (1) update_connections: drop links now below threshold (base.py:221-227)
        for u in range(self.num_ues):
            self.conn[u] = {b for b in self.conn[u] if self.connectable(b, u)}
        # (2) apply actions: 0 = NOOP; a>0 toggles BS a-1 (connect only if connectable)
        for u, a in enumerate(actions):
            a = int(a)
            if a == NOOP_ACTION:
                continue
            b = a - 1
            if b in self.conn[u]:
                self.conn[u].discard(b)
            elif self.connectable(b, u):
                self.conn[u].add(b)
        # (3) allocate + (4) utility
        bs_conns = self._bs_conns()
        pair, total = self._allocate(bs_conns)
        util = self._utilities(total)
        self.utilities = dict(enumerate(util))
        # (5) reward
        if handler == "central":
            reward = float(np.mean(util))
        else:
            bs_util = self._bs_utilities(bs_conns)
            reward = []
            for u in range(self.num_ues):
                ok = [b for b in range(B) if self.connectable(b, u)]
                ngbr_u = sum(bs_util[b] for b in ok)
                ngbr_c = sum(len(bs_conns[b]) for b in ok)
                reward.append((ngbr_u + util[u]) / (ngbr_c + 1))
        info = {
            "conn": [sorted(c) for c in self.conn],
            "pair_rates": pair,
            "rate": [float(total.get(u, 0.0)) for u in range(self.num_ues)],
            "utility": util,
            "n_connections": sum(len(c) for c in bs_conns),
            "n_connected": sum(1 for c in self.conn if c),
            "mean_utility": float(np.mean(util)),
            "mean_datarate": float(np.mean(list(total.values()))) if total else 0.0,
            "bs_utility": [float(v) for v in self._bs_utilities(bs_conns)],  # base.py:438-447
        }
        # (6) move, clock, departures
        self._move_all()
        self.t += 1
        done = self.t >= self.p.ep_time
        if done:  # NoDeparture: everyone leaves at ep_time (arrival.py:32-36; base.py:283-291)
            self.conn = [set() for _ in range(self.num_ues)]
        obs = self.observe(handler, active=not done)
        info["pos"] = list(self.pos)
        return obs, reward, done, info


# --------------------------------------------------------------------------------------
# Vectorised form over [E,U,B] (same arithmetic; FP64)
# --------------------------------------------------------------------------------------
def batch_snr(p: Params, pos, bs_xy):
    """pos [E,U,2] int, bs_xy [E,B,2] or [B,2] -> snr [E,U,B] FP64 (channels.py:24-27,132-146)."""
    pos = np.asarray(pos, dtype=np.int64)
    bs = np.asarray(bs_xy, dtype=np.int64)
    if bs.ndim == 2:
        bs = bs[None]
    dx = pos[:, :, None, 0] - bs[:, None, :, 0]
    dy = pos[:, :, None, 1] - bs[:, None, :, 1]
    d2 = dx * dx + dy * dy
    dist = np.sqrt(d2.astype(np.float64))
    ch = 0.8 + (1.1 * np.log10(p.freq) - 0.7) * p.ue_height - 1.56 * np.log10(p.freq)
    tmp_1 = 69.55 - ch + 26.16 * np.log10(p.freq) - 13.82 * np.log10(p.bs_height)
    tmp_2 = 44.9 - 6.55 * np.log10(p.bs_height)
    loss = tmp_1 + tmp_2 * np.log10(dist + EPSILON)
    with np.errstate(over="ignore"):
        snr = 10 ** ((p.tx - loss) / 10) / p.noise
    return snr, d2


def batch_move(p: Params, pos, wp, new_wp):
    """pos, wp [E,U,2] int64 (wp[...,0] < 0 = none); new_wp [E,U,2] used where none.
    Returns pos', wp', drew[E,U] (movement.py:42-62)."""
    pos = np.asarray(pos, dtype=np.int64)
    wp = np.asarray(wp, dtype=np.int64).copy()
    drew = wp[..., 0] < 0
    wp[drew] = np.asarray(new_wp, dtype=np.int64)[drew]
    v = (wp - pos).astype(np.float64)
    norm = np.sqrt(v[..., 0] * v[..., 0] + v[..., 1] * v[..., 1])
    arrived = norm <= p.velocity
    with np.errstate(invalid="ignore", divide="ignore"):
        stepped = np.round(pos + p.velocity * v / norm[..., None])
    stepped = np.where(arrived[..., None], wp, stepped).astype(np.int64)
    wp_out = np.where(arrived[..., None], -1, wp)
    return stepped, wp_out, drew


def batch_allocate(p: Params, snr, conn):
    """snr [E,U,B], conn [E,U,B] bool -> pair rate [E,U,B] (rounded), total [E,U]
    (base.py:421-435, schedules.py:20-22, base.py:413-418)."""
    n = conn.sum(axis=1, keepdims=True)  # [E,1,B]
    with np.errstate(over="ignore"):
        raw = np.where(snr > p.snr_tr, p.bw * np.log2(1 + snr), 0.0)
    if p.scheduler == "proportional_fair":
        fixed = np.where(conn, np.rint(raw * PF_SCALE), 0.0).astype(np.int64)
        tot = fixed.sum(axis=1, keepdims=True).astype(np.float64) * (1.0 / PF_SCALE)  # [E,1,B]
        with np.errstate(divide="ignore", invalid="ignore"):
            share = np.where(conn, raw * raw / tot, 0.0)
    elif p.scheduler == "rate_fair":
        with np.errstate(divide="ignore", invalid="ignore"):
            inv = np.where(conn, np.rint(RF_SCALE / np.where(conn, raw, 1.0)), 0.0).astype(np.int64)
            tot = inv.sum(axis=1, keepdims=True).astype(np.float64) * (1.0 / RF_SCALE)  # [E,1,B]
            share = np.where(conn, 1.0 / tot, 0.0)
    else:
        with np.errstate(divide="ignore", invalid="ignore"):
            share = np.where(conn, raw / n, 0.0)
    pair = np.round(share, 2)
    total = np.zeros(pair.shape[:2])
    for b in range(pair.shape[2]):  # bs-major accumulation order
        total = total + pair[:, :, b]
    return pair, total


def batch_utility(p: Params, total):
    w1, w2, w3 = p.util_coeffs
    with np.errstate(divide="ignore", invalid="ignore"):
        u = np.clip(w1 * np.log(w2 + total) / np.log(w3), p.util_lower, p.util_upper)
    u = np.where(total <= 0.0, p.util_lower, u)
    return 2 * (u - p.util_lower) / (p.util_upper - p.util_lower) - 1


def batch_assoc_fork(p: Params, snr, d2, nbs=None):
    """Nearest eligible BS, first-min tie-break (base.py:236-241). -1 = none.
    nbs [E] optionally limits the number of valid BSs per env."""
    elig = snr > p.snr_tr
    if nbs is not None:
        elig = elig & (np.arange(snr.shape[2])[None, None, :] < np.asarray(nbs)[:, None, None])
    big = np.iinfo(np.int64).max
    key = np.where(elig, d2, big)
    idx = key.argmin(axis=2)
    return np.where(elig.any(axis=2), idx, -1), elig


def batch_step_fork(p: Params, pos, wp, new_wp, bs_xy, t, nbs=None):
    pos, wp, drew = batch_move(p, pos, wp, new_wp)
    snr, d2 = batch_snr(p, pos, bs_xy)
    assoc, elig = batch_assoc_fork(p, snr, d2, nbs)
    B = snr.shape[2]
    conn = assoc[:, :, None] == np.arange(B)[None, None, :]
    pair, total = batch_allocate(p, snr, conn)
    util = batch_utility(p, total)
    connected = assoc >= 0
    ncon = connected.sum(axis=1)
    with np.errstate(invalid="ignore", divide="ignore"):
        mean_rate = np.where(ncon > 0, (total * connected).sum(axis=1) / ncon, 0.0)
    t = np.asarray(t) + 1
    return {
        "pos": pos, "wp": wp, "drew": drew, "snr": snr, "d2": d2, "elig": elig,
        "assoc": assoc, "pair": pair, "rate": total, "utility": util,
        "n_connected": ncon, "mean_utility": util.mean(axis=1), "mean_datarate": mean_rate,
        "t": t, "done": t >= p.ep_time,
    }


def batch_bs_utility(p: Params, conn, util):
    """allStationUtilities (base.py:438-447): [E,B]."""
    idle = 2 * (p.util_lower - p.util_lower) / (p.util_upper - p.util_lower) - 1
    n = conn.sum(axis=1)  # [E,B]
    s = (conn * util[:, :, None]).sum(axis=1)
    with np.errstate(invalid="ignore", divide="ignore"):
        return np.where(n > 0, s / n, idle), n


def batch_observe(p: Params, pos, bs_xy, conn, util, handler="central", active=None):
    """[E,U,F] float32 features (see ScalarEnv.observe). util None -> idle everywhere."""
    snr, _ = batch_snr(p, pos, bs_xy)
    E, U, B = snr.shape
    idle = -1.0
    if util is None:
        util = np.full((E, U), idle)
        bs_util = np.full((E, B), idle)
        n = conn.sum(axis=1)
    else:
        bs_util, n = batch_bs_utility(p, conn, util)
    with np.errstate(invalid="ignore", over="ignore"):
        ratio = snr / snr.max(axis=2, keepdims=True)
    parts = [conn.astype(np.float64), ratio, util[:, :, None]]
    if handler != "central":
        ok = snr > p.snr_tr
        bcast = np.where(ok, bs_util[:, None, :], idle)
        cnt = np.where(ok, n[:, None, :].astype(np.float64), 0.0)
        tot = np.maximum(1.0, cnt.sum(axis=2, keepdims=True))
        parts += [bcast, cnt / tot]
    obs = np.concatenate(parts, axis=2).astype(np.float32)
    if active is not None:
        obs = np.where(np.asarray(active)[:, None, None], obs, np.float32(0))
    return obs


def batch_step_gym(p: Params, pos, wp, new_wp, bs_xy, conn, actions, t, handler="central"):
    """Vectorised GYM step (see ScalarEnv.step_gym).  conn [E,U,B] bool, actions [E,U] int."""
    snr, _ = batch_snr(p, pos, bs_xy)
    E, U, B = snr.shape
    ok = snr > p.snr_tr
    conn = conn & ok
    a = np.asarray(actions, dtype=np.int64)
    sel = (a[:, :, None] - 1) == np.arange(B)[None, None, :]
    conn = np.where(sel, np.where(conn, False, ok), conn)
    pair, total = batch_allocate(p, snr, conn)
    util = batch_utility(p, total)
    bs_util, n = batch_bs_utility(p, conn, util)
    if handler == "central":
        reward = util.mean(axis=1)
    else:
        ngbr_u = (ok * bs_util[:, None, :]).sum(axis=2)
        ngbr_c = (ok * n[:, None, :]).sum(axis=2)
        reward = (ngbr_u + util) / (ngbr_c + 1)
    connected = conn.any(axis=2)
    ncon = connected.sum(axis=1)
    with np.errstate(invalid="ignore", divide="ignore"):
        mean_rate = np.where(ncon > 0, (total * connected).sum(axis=1) / ncon, 0.0)
    pos2, wp2, drew = batch_move(p, pos, wp, new_wp)
    t = np.asarray(t) + 1
    done = t >= p.ep_time
    conn_out = conn & ~done[:, None, None]
    obs = batch_observe(p, pos2, bs_xy, conn_out, util, handler, active=~done)
    return {
        "pos": pos2, "wp": wp2, "drew": drew, "snr": snr, "conn": conn_out, "conn_pre": conn,
        "pair": pair, "rate": total, "utility": util, "reward": reward, "obs": obs,
        "n_connections": conn.sum(axis=(1, 2)), "n_connected": ncon,
        "mean_utility": util.mean(axis=1), "mean_datarate": mean_rate, "t": t, "done": done,
    }
