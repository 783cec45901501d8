"""TEST INFRASTRUCTURE ONLY -- drives the *real* reference from /root/reference (or its install
under ``oracle/_ref``, see ``oracle/build_ref.py``).

Nothing under ``oracle/`` is product code: only ``tests/``, ``__graft_entry__.smoke()``
and the ``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import it.

This module exists so that the numpy restatement in ``oracle/mbe_oracle.py`` can be
pinned against the reference's own ``MComCore.step`` (reference
``mobile_env/core/base.py:230-296``) and so that golden vectors can be generated
(``oracle/gen_golden.py`` -> ``tests/golden/*.json``).  ``/root/reference`` exists only in the build
container; the pip-installed copy ``oracle/_ref`` (git-ignored, built by ``build()``) travels to the
GPU box, where only bench.py's CPU legs use it (``oracle/cpu_baseline.py``) -- the ``-m gpu`` tests and
``smoke()`` compare against the oracle and the committed fixtures, never against this module.

The reference imports shapely / matplotlib / pygame / svgpath2mpl at module top
(``core/base.py:8-15``, ``core/util.py:3-4,24-28``, ``core/entities.py:3``); none is
installed here, so inert stand-ins are put into ``sys.modules`` first.  The only one
with arithmetic is ``shapely.geometry.Point.distance`` (requirements.txt:9,
shapely~=2.0.6): planar Euclidean distance of two points, restated with
``math.hypot`` (call site ``core/channels.py:134``).
"""
from __future__ import annotations

import math
import os
import sys
import types

_INSTALLED = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")


def _pick_root() -> str:
    env = os.environ.get("MBE_REFERENCE_ROOT")
    if env:
        return env
    if os.path.isdir("/root/reference/mobile_env/core"):
        return "/root/reference"
    return _INSTALLED


REFERENCE_ROOT = _pick_root()


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "mobile_env", "core"))


class _Point:
    """Stand-in for shapely.geometry.Point (x, y, distance only)."""

    def __init__(self, x, y):
        self.x, self.y = x, y

    def distance(self, other):
        return math.hypot(self.x - other.x, self.y - other.y)


class _Inert(types.ModuleType):
    """Attribute-returning module used for the rendering libraries (never rendered)."""

    def __getattr__(self, key):
        if key.startswith("__"):
            raise AttributeError(key)
        return _Inert(self.__name__ + "." + key)

    def __call__(self, *a, **k):
        return _Inert("call")

    def __isub__(self, other):
        return self

    def mean(self, *a, **k):
        return 0


def install_stubs() -> None:
    if "shapely" not in sys.modules:
        shp = types.ModuleType("shapely")
        geo = types.ModuleType("shapely.geometry")
        geo.Point = _Point
        shp.geometry = geo
        sys.modules["shapely"] = shp
        sys.modules["shapely.geometry"] = geo
    for name in (
        "matplotlib",
        "matplotlib.patheffects",
        "matplotlib.pyplot",
        "matplotlib.backends",
        "matplotlib.backends.backend_agg",
        "matplotlib.transforms",
        "pygame",
        "svgpath2mpl",
    ):
        if name not in sys.modules:
            sys.modules[name] = _Inert(name)
    if not hasattr(sys.modules["matplotlib"], "cm") or True:
        sys.modules["matplotlib"].cm = _Inert("matplotlib.cm")


def import_reference():
    """Returns the reference's modules (base, entities, custom) -- unmodified code."""
    if not reference_available():
        raise RuntimeError(f"reference not found under {REFERENCE_ROOT}")
    install_stubs()
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    from mobile_env.core import base, entities  # noqa: E402
    from mobile_env.scenarios import custom  # noqa: E402

    return base, entities, custom


def make_fixed_layout_env(bs_xy, num_ues, config=None, bs_params=None, ue_params=None, bs_over=None, ue_over=None):
    """Reference env with a fixed BS list (what MComCustom does at custom.py:40-62,
    minus the unseeded ``random`` BS generator) and JSON dumps disabled."""
    base, entities, _ = import_reference()

    class FixedLayoutEnv(base.MComCore):
        def reset(self, *, seed=None):
            super().reset(seed=seed)
            users = [ue for ue in self.userDict.values() if ue.startTime <= 0]
            self.activeUsers = sorted(users, key=lambda ue: ue.ue_id)
            self.users_dataRateList = {ue.ue_id: [] for ue in self.userDict.values()}
            self.users_trajectoryList = {ue.ue_id: [] for ue in self.userDict.values()}

        def save_layout_and_data_rates(self, epoch_number, curr_step):  # dumps off
            return

    cfg = FixedLayoutEnv.default_config()
    if config:
        from mobile_env.core.util import deep_dict_merge

        cfg = deep_dict_merge(cfg, config)
    bsp = dict(cfg["bs"])
    if bs_params:
        bsp.update(bs_params)
    uep = dict(cfg["ue"])
    if ue_params:
        uep.update(ue_params)
    stations = []
    for i, xy in enumerate(bs_xy):
        kw = dict(bsp)
        if bs_over and bs_over.get(i):
            kw.update(bs_over[i])  # keys of BaseStation.__init__: bw, freq, tx, height
        stations.append(entities.BaseStation(i, tuple(xy), **kw))
    users = []
    for i in range(num_ues):
        kw = dict(uep)
        if ue_over and ue_over.get(i):
            kw.update(ue_over[i])  # keys of UserEquipment.__init__: velocity, snr_tr, noise, height
        users.append(entities.UserEquipment(i, **kw))
    return FixedLayoutEnv(stations, users, config or {})


def custom_channels():
    """Channel subclasses of the REFERENCE's Channel that override power_loss only (the reference's
    plugin contract, channels.py:18-21).  PathLoss is the README's example verbatim (README.md:108-121,
    with the module path the fork really has); TwoSlope is not affine in log-distance."""
    import_reference()
    import numpy as np
    from mobile_env.core.channels import Channel

    class PathLoss(Channel):
        def __init__(self, gamma, **kwargs):
            super().__init__(**kwargs)
            # path loss exponent
            self.gamma = gamma

        def power_loss(self, bs, ue):
            """Computes power loss between BS and UE."""
            dist = bs.point.distance(ue.point)
            loss = 10 * self.gamma * np.log10(4 * np.pi * dist * bs.frequency)
            return loss

    class TwoSlope(Channel):
        def __init__(self, gamma1, gamma2, d_break, **kwargs):
            super().__init__(**kwargs)
            self.gamma1, self.gamma2, self.d_break = gamma1, gamma2, d_break

        def power_loss(self, bs, ue):
            dist = bs.point.distance(ue.point)
            if dist <= self.d_break:
                return 10 * self.gamma1 * np.log10(4 * np.pi * dist * bs.frequency)
            return (10 * self.gamma1 * np.log10(4 * np.pi * self.d_break * bs.frequency)
                    + 10 * self.gamma2 * np.log10(dist / self.d_break))

    return {"pathloss": PathLoss, "two_slope": TwoSlope}


def record_fork_episode(env, steps, init_pos=None):
    """Runs reset + ``steps`` reference steps, recording everything a parity test needs.

    Returns a dict of plain lists: per step positions, the waypoint each UE held when it
    moved (so the draw sequence can be injected), SNR matrix, connection index per UE,
    per-(bs,ue) rounded rates, per-UE total rate, scaled utility, monitor scalars, done.
    """
    env.reset()
    ues = [env.userDict[k] for k in sorted(env.userDict)]
    bss = [env.stationDict[k] for k in sorted(env.stationDict)]
    if init_pos is not None:
        for ue, (x, y) in zip(ues, init_pos):
            ue.x, ue.y = x, y
    rec = {
        "bs_xy": [[bs.x, bs.y] for bs in bss],
        "init_pos": [[int(ue.x), int(ue.y)] for ue in ues],
        "steps": [],
    }
    # wrap move() to log the waypoint each UE targets (movement.py:42-62)
    mv = env.movementModel
    orig_move = mv.move
    targets = {}

    def logged_move(ue):
        had = ue in mv.userMoveDirection
        out = orig_move(ue)
        # waypoint used this call: either still stored, or equals the snap result
        wp = mv.userMoveDirection.get(ue, out)
        targets[ue.ue_id] = (int(wp[0]), int(wp[1]), 0 if had else 1)
        return out

    mv.move = logged_move
    for s in range(steps):
        targets.clear()
        env.step(0, s)
        # association used by this step's allocation (keys of bs2ue_dataRates, base.py:435);
        # bs2ue_connections itself is emptied of leaving UEs at the last step (base.py:283-285)
        conn = [-1] * len(ues)
        for (bs, ue) in env.bs2ue_dataRates:
            conn[ue.ue_id] = bs.bs_id
        conn_after = [-1] * len(ues)
        for bs, cues in env.bs2ue_connections.items():
            for ue in cues:
                conn_after[ue.ue_id] = bs.bs_id
        snr = [[float(env.channelModel.calculateSNR(bs, ue)) for bs in bss] for ue in ues]
        pair_rates = sorted(
            [[ue.ue_id, bs.bs_id, float(r)] for (bs, ue), r in env.bs2ue_dataRates.items()]
        )
        info = env.monitor.info()
        rec["steps"].append(
            {
                "pos": [[int(ue.x), int(ue.y)] for ue in ues],
                "wp": [list(targets.get(ue.ue_id, (-1, -1, 0))) for ue in ues],
                "snr": snr,
                "conn": conn,
                "conn_after": conn_after,
                "pair_rates": pair_rates,
                "rate": [float(env.allUserDataRates.get(ue, 0.0)) for ue in ues],
                "utility": [float(env.ue_utilities.get(ue, float("nan"))) for ue in ues],
                "n_connections": int(info["number connections"]),
                "n_connected": int(info["number connected"]),
                "mean_utility": float(info["mean utility"]),
                "mean_datarate": float(info["mean datarate"]),
                "time": float(env.time),
                "done": bool(env.time_is_up),
            }
        )
    mv.move = orig_move
    return rec


def record_gym_pieces_episode(env, actions, init_pos=None):
    """GYM-order episode assembled from the reference's OWN primitives.

    The fork has no GYM step, so the ORDER below (update_connections -> action toggle -> allocate ->
    utility -> BS utilities -> move -> clock) is this build's specification; but every stage is
    executed by the unmodified reference code on its own objects:

      * ``MComCore.update_connections``            base.py:221-227
      * ``MComCore.check_connectivity``            base.py:212-214   (gate of a connect action)
      * ``MComCore.allocateDataRate2User``         base.py:421-435   (scheduler share + round)
      * ``MComCore.user_total_datarates``          base.py:413-418   (a UE on SEVERAL BSs: the
                                                   multi-connection sum the FORK step never exercises)
      * ``utilityModel.calculateUtility/scaleUtility``  base.py:253-258, utilities.py:44-55
      * ``MComCore.allStationUtilities``           base.py:438-447
      * ``movementModel.move``                     movement.py:42-62
      * clock / leaving UEs                        base.py:280-291

    ``actions``: [T][U] ints, 0 = NOOP (base.py:29), a > 0 toggles BS a-1.  Returns plain lists like
    ``record_fork_episode`` (waypoint triples for injection included)."""
    env.reset()
    ues = [env.userDict[k] for k in sorted(env.userDict)]
    bss = [env.stationDict[k] for k in sorted(env.stationDict)]
    if init_pos is not None:
        for ue, (x, y) in zip(ues, init_pos):
            ue.x, ue.y = x, y
    rec = {"bs_xy": [[bs.x, bs.y] for bs in bss], "init_pos": [[int(ue.x), int(ue.y)] for ue in ues],
           "actions": [list(map(int, a)) for a in actions], "steps": []}
    mv = env.movementModel
    orig_move = mv.move
    targets = {}

    def logged_move(ue):
        had = ue in mv.userMoveDirection
        out = orig_move(ue)
        wp = mv.userMoveDirection.get(ue, out)
        targets[ue.ue_id] = (int(wp[0]), int(wp[1]), 0 if had else 1)
        return out

    mv.move = logged_move
    for acts in actions:
        targets.clear()
        env.update_connections()
        for ue, a in zip(ues, acts):
            if a == 0 or ue not in env.activeUsers:
                continue
            bs = bss[a - 1]
            if ue in env.bs2ue_connections[bs]:
                env.bs2ue_connections[bs].remove(ue)
            elif env.check_connectivity(bs, ue):
                env.bs2ue_connections[bs].add(ue)
        env.bs2ue_dataRates = {}
        for bs in bss:
            env.bs2ue_dataRates.update(env.allocateDataRate2User(bs))
        env.allUserDataRates = env.user_total_datarates(env.bs2ue_dataRates)
        env.ue_utilities = {
            ue: env.utilityModel.scaleUtility(env.utilityModel.calculateUtility(env.allUserDataRates.get(ue, 0.0)))
            for ue in env.activeUsers
        }
        bs_util = env.allStationUtilities()
        conn = [sorted(bs.bs_id for bs in bss if ue in env.bs2ue_connections[bs]) for ue in ues]
        step = {
            "conn": conn,
            "pair_rates": sorted([ue.ue_id, bs.bs_id, float(r)] for (bs, ue), r in env.bs2ue_dataRates.items()),
            "rate": [float(env.allUserDataRates.get(ue, 0.0)) for ue in ues],
            "utility": [float(env.ue_utilities.get(ue, float("nan"))) for ue in ues],
            "bs_utility": [float(bs_util[bs]) for bs in bss],
            "connectable": [[bool(env.check_connectivity(bs, ue)) for bs in bss] for ue in ues],
        }
        for ue in env.activeUsers:
            ue.x, ue.y = mv.move(ue)
        env.time += 1
        leaving = set(ue for ue in env.activeUsers if ue.exitTime <= env.time)
        for bs, cues in env.bs2ue_connections.items():
            env.bs2ue_connections[bs] = cues - leaving
        env.activeUsers = sorted(
            [ue for ue in env.userDict.values() if ue.exitTime > env.time >= ue.startTime], key=lambda ue: ue.ue_id)
        step.update({
            "pos": [[int(ue.x), int(ue.y)] for ue in ues],
            "wp": [list(targets.get(ue.ue_id, (-1, -1, 0))) for ue in ues],
            "conn_after": [sorted(bs.bs_id for bs in bss if ue in env.bs2ue_connections[bs]) for ue in ues],
            "done": bool(env.time_is_up),
        })
        rec["steps"].append(step)
    mv.move = orig_move
    return rec
