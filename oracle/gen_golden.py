"""TEST INFRASTRUCTURE ONLY -- generates tests/golden/*.json by running the UNMODIFIED
reference (``/root/reference``, via ``oracle/ref_harness.py``).  Run in the build container:

    python oracle/gen_golden.py

Each case records a full FORK-mode episode of ``MComCore.step`` (reference
``mobile_env/core/base.py:230-296``): positions, waypoint targets, SNR matrix,
association, rounded pair rates, utilities, monitor scalars and ``time_is_up``.
The notebook known-answer vectors (``mobile_env/GNN/GNN.ipynb`` cell 3 output raw 86-93,
cell 17 output raw 723-1050) are asserted while generating so the fixtures are pinned
to what the reference's authors saw.
"""
from __future__ import annotations

import json

import numpy as np
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))

from oracle import ref_harness as rh  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")

KAT1_BS = [(181, 153), (75, 68), (128, 32), (74, 108), (176, 115)]
KAT1_POS = [(81, 109), (142, 187), (161, 86), (156, 91), (70, 21), (10, 48), (177, 108)]
KAT1_RATES = {(1, 0): 0.84, (5, 1): 0.17, (4, 1): 0.57, (0, 3): 700.2, (2, 4): 1.33,
              (6, 4): 233.4, (3, 4): 1.55}
KAT2_BS = [(194, 153), (150, 104), (27, 70), (28, 186), (26, 127), (69, 172), (23, 140),
           (189, 68), (63, 183), (31, 48)]
KAT2_POS = [(65, 33), (54, 55), (33, 150), (97, 25), (36, 124), (54, 144), (43, 129)]
# (ue, graph node, rate); graph node = 7 + bs_id
KAT2_RATES = {(0, 9): 0.86, (1, 9): 3.74, (2, 6): 33.7, (3, 9): 0.1, (4, 4): 93.91,
              (5, 6): 2.31, (6, 4): 17.68}

KAT_CFG = {"ue": {"velocity": 10, "height": 1.8}, "bs": {"tx": 30}}


def ep_cfg(n, extra=None):
    cfg = {"EP_MAX_TIME": n, "arrival_params": {"ep_time": n}}
    if extra:
        cfg.update(extra)
    return cfg


CASES = {
    # name: (bs_xy, num_ues, config, steps)
    "kat1": (KAT1_BS, 7, KAT_CFG, 20),
    "kat2": (KAT2_BS, 7, KAT_CFG, 20),
    "default_v10": ([(40, 150), (100, 100), (160, 40), (30, 30), (170, 170), (100, 20), (20, 100), (150, 110)],
                    7, {"ue": {"velocity": 10}}, 20),
    "small_v1p5": ([(110, 130), (65, 80), (120, 30)], 5, ep_cfg(60), 60),
    "medium_v1p5": ([(50, 50), (150, 50), (50, 150), (150, 150)], 15, ep_cfg(40), 40),
    "medium_v3": ([(50, 50), (150, 50), (50, 150), (150, 150)], 15, ep_cfg(40, {"ue": {"velocity": 3}}), 40),
    "large_v5": ([(20 + 45 * (i % 4), 25 + 50 * (i // 4)) for i in range(13)], 30,
                 ep_cfg(30, {"ue": {"velocity": 5}}), 30),
    # UE0 of the seed-2028 trajectory stands exactly on a BS at step 0 (d = 0 => log10(1e-16))
    "ue_on_bs": ([(81, 109), (142, 187), (10, 10)], 7, {"ue": {"velocity": 10}}, 20),
    # weak transmitters: nobody in range for most steps
    "out_of_range": ([(0, 0), (199, 199)], 7, {"ue": {"velocity": 10}, "bs": {"tx": 5}}, 20),
    # everyone on one BS
    "single_bs": ([(100, 100)], 15, ep_cfg(25, {"ue": {"velocity": 7}, "bs": {"tx": 46}}), 25),
    # distance ties: BSs mirrored around the map centre line, first-min tie-break matters
    "tie_layout": ([(100, 60), (100, 140), (60, 100), (140, 100), (100, 60)], 15,
                   ep_cfg(30, {"ue": {"velocity": 4}}), 30),
    # non-square map, fast UEs: waypoints are drawn per axis (movement.py:45-46)
    "nonsquare_fast": ([(30, 40), (150, 60), (270, 30), (210, 100)], 9,
                       ep_cfg(25, {"width": 300, "height": 120, "ue": {"velocity": 25},
                                   "movement_params": {"width": 300, "height": 120}}), 25),
    # another carrier and mast height: different folded constants and cut-off distance
    "low_freq": ([(40, 40), (160, 40), (100, 160)], 8,
                 ep_cfg(20, {"bs": {"freq": 900, "height": 30, "tx": 33, "bw": 5e6}, "ue": {"velocity": 6, "height": 2.0}}), 20),
    # a long episode at the default crawl speed (v = 1.5: many exact .5 rounding ties)
    "long_crawl": ([(60, 60), (140, 60), (60, 140), (140, 140)], 6, ep_cfg(100), 100),
}

# a wide shape (more than 32 UEs and more than 32 BSs): the block-per-env kernel's territory
CASES["wide_40x35"] = ([((i * 37 + 11) % 200, (i * 53 + 29) % 200) for i in range(35)], 40,
                       ep_cfg(12, {"ue": {"velocity": 9}}), 12)

# movement corner cases: v = 2.5 gives exact .5 rounding ties on axis-aligned legs (np.round half-even,
# movement.py:60); v = 0.4 rounds most steps to no move at all; v = 150 snaps to every waypoint at once
CASES["v2p5_ties"] = ([(50, 50), (150, 50), (100, 150)], 10, ep_cfg(60, {"ue": {"velocity": 2.5}}), 60)
CASES["v0p4_crawl"] = ([(50, 50), (150, 50), (100, 150)], 6, ep_cfg(40, {"ue": {"velocity": 0.4}}), 40)
CASES["v150_teleport"] = ([(50, 50), (150, 50), (100, 150), (20, 180)], 8, ep_cfg(30, {"ue": {"velocity": 150}}), 30)
# non-default utility curve and receiver: w1*log(w2 + r)/log(w3) clipped to [-5, 25] (utilities.py:44-55)
CASES["utility_custom"] = ([(60, 60), (140, 60), (60, 140), (140, 140)], 9,
                           ep_cfg(25, {"ue": {"velocity": 6, "snr_tr": 1e-7, "noise": 5e-10},
                                       "utility_params": {"lower": -5, "upper": 25, "coeffs": (3, 1, 2)}}), 25)

# per-BS radio overrides (BaseStation keyword names): the reference keeps bw/freq/tx/height per BS
BS_OVERRIDES = {
    "two_classes": {1: {"tx": 30}, 3: {"tx": 30, "bw": 18e6}},
}
CASES["two_classes"] = ([(50, 50), (150, 50), (50, 150), (150, 150), (100, 100)], 12,
                        ep_cfg(25, {"ue": {"velocity": 8}}), 25)


# ---- the plugin surface: per-UE parameters (entities.py:32-57) and Channel subclasses that override
# power_loss only (channels.py:18-21; README.md:108-121).  name: (bs_xy, nue, cfg, steps, bs_over, ue_over, channel)
PLUGIN_CASES = {
    "two_ue_classes": ([(50, 50), (150, 50), (50, 150), (150, 150)], 10, ep_cfg(30, {"ue": {"velocity": 6}}), 30, None,
                       {1: {"velocity": 2.5}, 4: {"velocity": 2.5}, 7: {"velocity": 2.5, "snr_tr": 1e-6},
                        8: {"snr_tr": 1e-6}}, None),
    "ue_bs_classes": ([(40, 40), (160, 60), (100, 160), (30, 150), (100, 90)], 12, ep_cfg(25, {"ue": {"velocity": 8}}), 25,
                      {1: {"tx": 30}, 3: {"tx": 30, "bw": 18e6}},
                      {0: {"noise": 4e-9}, 3: {"height": 2.2, "velocity": 3}, 5: {"noise": 4e-9},
                       9: {"height": 2.2, "velocity": 3}, 10: {"snr_tr": 5e-8, "velocity": 15}}, None),
    "pathloss_readme": ([(50, 50), (150, 50), (100, 150)], 9, ep_cfg(25, {"ue": {"velocity": 7}}), 25, None, None,
                        ("pathloss", 2.0)),
    "pathloss_short_range": ([(50, 50), (150, 50), (100, 150), (100, 100)], 12,
                             ep_cfg(30, {"ue": {"velocity": 9, "snr_tr": 1e-3}}), 30, None, None, ("pathloss", 2.6)),
    "two_slope": ([(60, 60), (140, 60), (60, 140), (140, 140)], 10, ep_cfg(30, {"ue": {"velocity": 8, "snr_tr": 0.1}}), 30,
                  None, {2: {"snr_tr": 1.0}, 6: {"snr_tr": 1.0}}, ("two_slope", 2.0, 4.5, 25.0)),
}
PLUGIN_GYM = {  # GYM-order episodes by the reference's primitives with the same plugins
    "two_ue_classes": ("two_ue_classes", 21),
    "two_slope": ("two_slope", 22),
}


def plugin_config(cfg, channel):
    """cfg + the reference-side channel class / params for a ("name", *args) channel spec."""
    if channel is None:
        return cfg
    cls = rh.custom_channels()[channel[0]]
    names = {"pathloss": ("gamma",), "two_slope": ("gamma1", "gamma2", "d_break")}[channel[0]]
    out = dict(cfg)
    out["channel"] = cls
    out["channel_params"] = dict(zip(names, channel[1:]))
    return out


def record_plugin_case(bs_xy, nue, cfg, steps, bs_over, ue_over, channel, actions=None):
    env = rh.make_fixed_layout_env(bs_xy, nue, config=plugin_config(cfg, channel), bs_over=bs_over, ue_over=ue_over)
    rec = rh.record_fork_episode(env, steps) if actions is None else rh.record_gym_pieces_episode(env, actions)
    shell = record_case(bs_xy, nue, cfg, 1, bs_over)
    rec["params"] = shell["params"]
    if bs_over:
        rec["bs_over"] = shell["bs_over"]
    if ue_over:
        ren = {"velocity": "velocity", "snr_tr": "snr_tr", "noise": "noise", "height": "ue_height"}
        rec["ue_over"] = [{ren[k]: v for k, v in ue_over.get(i, {}).items()} for i in range(nue)]
    if channel:
        rec["params"]["channel"] = list(channel)
    return rec


def plugin_golden():
    """-> tests/golden/fork_<name>.json and gymref_<name>.json for the plugin-surface cases."""
    for name, (bs_xy, nue, cfg, steps, bs_over, ue_over, channel) in PLUGIN_CASES.items():
        rec = record_plugin_case(bs_xy, nue, cfg, steps, bs_over, ue_over, channel)
        path = os.path.join(OUT, f"fork_{name}.json")
        with open(path, "w") as f:
            json.dump(rec, f, separators=(",", ":"))
        nconn = sum(st["n_connected"] for st in rec["steps"])
        print(os.path.basename(path), steps, "steps,", nconn, "connected UE-steps,", os.path.getsize(path), "bytes")
    for name, (case, aseed) in PLUGIN_GYM.items():
        bs_xy, nue, cfg, steps, bs_over, ue_over, channel = PLUGIN_CASES[case]
        rng = np.random.default_rng(aseed)
        actions = [[int(a) for a in rng.integers(0, len(bs_xy) + 1, size=nue)] for _ in range(steps)]
        rec = record_plugin_case(bs_xy, nue, cfg, steps, bs_over, ue_over, channel, actions)
        multi = sum(len(c) > 1 for st in rec["steps"] for c in st["conn"])
        assert multi > 0, "the fixture must exercise UEs connected to several BSs"
        path = os.path.join(OUT, f"gymref_{name}.json")
        with open(path, "w") as f:
            json.dump(rec, f, separators=(",", ":"))
        print(os.path.basename(path), steps, "steps,", multi, "multi-connection UE-steps,", os.path.getsize(path), "bytes")


def random_case(seed):
    """A random scenario (layout, UE count, speed, radio, map, utility curve, sometimes per-BS
    overrides): the generator behind the `rand_*` fixtures and the live cross-check in
    tests/test_oracle_golden.py."""
    rng = np.random.default_rng(1000 + seed)
    W, H = int(rng.integers(60, 400)), int(rng.integers(60, 400))
    nbs, nue, steps = int(rng.integers(1, 11)), int(rng.integers(1, 21)), int(rng.integers(5, 40))
    bs_xy = [(int(rng.integers(0, W)), int(rng.integers(0, H))) for _ in range(nbs)]
    cfg = ep_cfg(steps, {
        "width": W, "height": H, "movement_params": {"width": W, "height": H},
        "bs": {"tx": float(rng.choice([20, 30, 40, 46])), "freq": float(rng.choice([900, 1800, 2500, 3500])),
               "height": float(rng.choice([25, 50, 80])), "bw": float(rng.choice([5e6, 9e6, 20e6]))},
        "ue": {"velocity": float(rng.choice([0.7, 1.5, 2.5, 3, 7.3, 10, 33])), "snr_tr": float(rng.choice([2e-8, 1e-7])),
               "noise": float(rng.choice([1e-9, 4e-10])), "height": float(rng.choice([1.5, 1.6, 2.0]))},
        "utility_params": {"lower": int(rng.choice([-20, -5])), "upper": int(rng.choice([20, 30])),
                           "coeffs": tuple(int(v) for v in rng.choice([[10, 0, 10], [3, 1, 2], [5, 2, 4]]))},
    })
    over = None
    if nbs >= 2 and seed % 3 == 0:
        over = {int(rng.integers(0, nbs)): {"tx": 25.0}, int(rng.integers(0, nbs)): {"bw": 15e6, "freq": 2000.0}}
    return bs_xy, nue, cfg, steps, over


def random_golden(seeds=range(6)):
    """Commits a few of the random scenarios as fixtures so the GPU box (no reference) replays them."""
    for seed in seeds:
        bs_xy, nue, cfg, steps, over = random_case(seed)
        rec = record_case(bs_xy, nue, cfg, steps, over)
        path = os.path.join(OUT, f"fork_rand_{seed:02d}.json")
        with open(path, "w") as f:
            json.dump(rec, f, separators=(",", ":"))
        print(os.path.basename(path), len(bs_xy), "BS", nue, "UE", steps, "steps", os.path.getsize(path), "bytes")


GYM_PIECES = {
    # name: (bs_xy, num_ues, config, steps, per-BS overrides, action seed)
    "medium": ([(50, 50), (150, 50), (50, 150), (150, 150)], 15, ep_cfg(20), 20, None, 11),
    "small_fast": ([(110, 130), (65, 80), (120, 30)], 5, ep_cfg(30, {"ue": {"velocity": 9}}), 30, None, 12),
    "custom_10bs": ([(40, 150), (100, 100), (160, 40), (30, 30), (170, 170), (100, 20), (20, 100), (150, 110),
                     (90, 180), (60, 60)], 7, ep_cfg(20, {"ue": {"velocity": 10}}), 20, None, 13),
    "two_classes": ([(50, 50), (150, 50), (50, 150), (150, 150), (100, 100)], 12,
                    ep_cfg(25, {"ue": {"velocity": 8}}), 25, {1: {"tx": 30}, 3: {"tx": 30, "bw": 18e6}}, 14),
}


def gym_pieces_golden():
    """GYM-ORDER episodes executed by the reference's own primitives (ref_harness.
    record_gym_pieces_episode): pins update_connections, the multi-connection allocation / per-UE
    sum, utilities and allStationUtilities of the GYM step; the stage order and the action semantics
    stay this build's specification.  -> tests/golden/gymref_*.json"""
    for name, (bs_xy, nue, cfg, steps, over, aseed) in GYM_PIECES.items():
        rng = np.random.default_rng(aseed)
        # mostly connect requests early on, then a mix with NOOPs and disconnects
        actions = [[int(a) for a in rng.integers(0, len(bs_xy) + 1, size=nue)] for _ in range(steps)]
        env = rh.make_fixed_layout_env(bs_xy, nue, config=cfg, bs_over=over)
        rec = rh.record_gym_pieces_episode(env, actions)
        shell = record_case(bs_xy, nue, cfg, 1, over)  # params / bs_over in the fixtures' format
        rec["params"] = shell["params"]
        if over:
            rec["bs_over"] = shell["bs_over"]
        multi = sum(len(c) > 1 for st in rec["steps"] for c in st["conn"])
        assert multi > 0, "the fixture must exercise UEs connected to several BSs"
        path = os.path.join(OUT, f"gymref_{name}.json")
        with open(path, "w") as f:
            json.dump(rec, f, separators=(",", ":"))
        print(os.path.basename(path), steps, "steps,", multi, "multi-connection UE-steps,", os.path.getsize(path), "bytes")


def record_case(bs_xy, nue, cfg, steps, over=None):
    """One episode of the unmodified reference on a fixed layout -> the record the golden files hold
    (also used live by tests/test_oracle_golden.py when /root/reference is present)."""
    env = rh.make_fixed_layout_env(bs_xy, nue, config=cfg, bs_over=over)
    rec = rh.record_fork_episode(env, steps)
    if over:  # oracle-side names (Params fields)
        ren = {"tx": "tx", "bw": "bw", "freq": "freq", "height": "bs_height"}
        rec["bs_over"] = [{ren[k]: v for k, v in over.get(i, {}).items()} for i in range(len(bs_xy))]
    p = env.default_config()
    from mobile_env.core.util import deep_dict_merge

    p = deep_dict_merge(p, cfg)
    rec["params"] = {
        "width": p["width"], "height": p["height"],
        "ep_time": min(p["EP_MAX_TIME"], p["arrival_params"]["ep_time"]),
        "bw": p["bs"]["bw"], "freq": p["bs"]["freq"], "tx": p["bs"]["tx"], "bs_height": p["bs"]["height"],
        "velocity": p["ue"]["velocity"], "snr_tr": p["ue"]["snr_tr"], "noise": p["ue"]["noise"],
        "ue_height": p["ue"]["height"],
        "util_lower": p["utility_params"]["lower"], "util_upper": p["utility_params"]["upper"],
        "util_coeffs": list(p["utility_params"]["coeffs"]),
    }
    return rec


def main():
    # NOTE: re-running changes `mean_datarate` of existing files in the last ulp (the reference sums a
    # dict whose iteration order depends on object hashes); tests compare that field with a tolerance.
    os.makedirs(OUT, exist_ok=True)
    for name, (bs_xy, nue, cfg, steps) in CASES.items():
        rec = record_case(bs_xy, nue, cfg, steps, BS_OVERRIDES.get(name))
        if name == "kat1":
            s0 = rec["steps"][0]
            assert [tuple(q) for q in s0["pos"]] == KAT1_POS, s0["pos"]
            got = {(u, b): r for u, b, r in s0["pair_rates"]}
            assert got == KAT1_RATES, got
        if name == "kat2":
            s19 = rec["steps"][19]
            assert [tuple(q) for q in s19["pos"]] == KAT2_POS, s19["pos"]
            got = {(u, b): r for u, b, r in s19["pair_rates"]}
            assert got == KAT2_RATES, got
        if name.startswith("wide_"):
            for st in rec["steps"]:
                st["snr"] = None  # 40 x 35 doubles per step: not needed, the rates pin the chain
        path = os.path.join(OUT, f"fork_{name}.json")
        with open(path, "w") as f:
            json.dump(rec, f, separators=(",", ":"))
        print(name, "steps", steps, "bytes", os.path.getsize(path))


def dump_golden(name="kat1", steps=20):
    """Runs the reference with its JSON/CSV dumps ENABLED (base.py:261,298-404; custom.py:79-85)
    in a scratch directory and copies the files into tests/golden/dumps/<name>/ -- the byte-level
    target of mobile_env_gan_b200/export.py."""
    import shutil
    import tempfile

    base, entities, custom = rh.import_reference()
    bs_xy, nue, cfg, _ = CASES[name]

    class DumpingEnv(base.MComCore):
        def reset(self, *, seed=None):
            super().reset(seed=seed)
            users = [ue for ue in self.userDict.values() if ue.startTime <= 0]
            self.activeUsers = sorted(users, key=lambda ue: ue.ue_id)
            self.users_dataRateList = {ue.ue_id: [] for ue in self.userDict.values()}
            self.users_trajectoryList = {ue.ue_id: [] for ue in self.userDict.values()}

    full = DumpingEnv.default_config()
    from mobile_env.core.util import deep_dict_merge

    full = deep_dict_merge(full, cfg)
    stations = [entities.BaseStation(i, tuple(xy), **full["bs"]) for i, xy in enumerate(bs_xy)]
    users = [entities.UserEquipment(i, **full["ue"]) for i in range(nue)]
    env = DumpingEnv(stations, users, cfg)
    tmp = tempfile.mkdtemp()
    run = os.path.join(tmp, "run")
    os.makedirs(run)
    cwd = os.getcwd()
    os.chdir(run)
    try:
        env.reset()
        custom.MComCustom.save_base_station_positions(env, 0)
        for s in range(steps):
            env.step(0, s)
        env.save_epoch_data(0)
    finally:
        os.chdir(cwd)
    dst = os.path.join(OUT, "dumps", name)
    shutil.rmtree(dst, ignore_errors=True)
    for sub in ("collectData", "collectData2"):
        shutil.copytree(os.path.join(tmp, sub), os.path.join(dst, sub))
    shutil.rmtree(tmp)
    n = sum(len(f) for _, _, f in os.walk(dst))
    print("dumps", name, n, "files")


def custom_epochs_golden(epochs=4, steps=20, seed=123):
    """The fork's own scenario exactly as shipped (MComCustom, custom.py:12-85): per epoch a fresh
    BS layout from the global ``random`` (seeded here), the same UE trajectory every epoch
    (movement reset_rng_episode=True, base.py:130-134), dumps disabled."""
    import random

    base, entities, custom = rh.import_reference()

    class Quiet(custom.MComCustom):
        def save_layout_and_data_rates(self, epoch_number, curr_step):
            return

    random.seed(seed)
    env = Quiet()
    out = {"epochs": []}
    mv = env.movementModel
    for ep in range(epochs):
        env.reset()
        ues = [env.userDict[k] for k in sorted(env.userDict)]
        bss = [env.stationDict[k] for k in sorted(env.stationDict)]
        rec = {"bs_xy": [[bs.x, bs.y] for bs in bss], "init_pos": [[int(u.x), int(u.y)] for u in ues], "steps": []}
        orig = mv.move
        targets = {}

        def logged(ue, orig=orig):
            had = ue in mv.userMoveDirection
            res = orig(ue)
            wp = mv.userMoveDirection.get(ue, res)
            targets[ue.ue_id] = (int(wp[0]), int(wp[1]), 0 if had else 1)
            return res

        mv.move = logged
        for s in range(steps):
            targets.clear()
            env.step(ep, s)
            conn = [-1] * len(ues)
            for (bs, ue) in env.bs2ue_dataRates:
                conn[ue.ue_id] = bs.bs_id
            rec["steps"].append({
                "pos": [[int(u.x), int(u.y)] for u in ues],
                "wp": [list(targets[u.ue_id]) for u in ues],
                "conn": conn,
                "rate": [float(env.allUserDataRates.get(u, 0.0)) for u in ues],
                "utility": [float(env.ue_utilities[u]) for u in ues],
                "done": bool(env.time_is_up),
            })
        mv.move = orig
        out["epochs"].append(rec)
    with open(os.path.join(OUT, "custom_epochs.json"), "w") as f:
        json.dump(out, f, separators=(",", ":"))
    print("custom_epochs", epochs, "epochs", [len(e["bs_xy"]) for e in out["epochs"]], "BSs")


def custom_dump_golden(epochs=2, steps=20, seed=123):
    """The fork's collect loop exactly as shipped (collectData2.ipynb: MComCustom, reset ->
    save_base_station_positions -> 20 x step (dumping inside) -> save_epoch_data) with the global
    ``random`` seeded like custom_epochs_golden, so epoch e here IS epoch e of custom_epochs.json.
    Files -> tests/golden/dumps/custom/ (5..10 BSs per epoch: the station files of varying length)."""
    import random
    import shutil
    import tempfile

    base, entities, custom = rh.import_reference()
    random.seed(seed)
    env = custom.MComCustom()
    tmp = tempfile.mkdtemp()
    run = os.path.join(tmp, "run")
    os.makedirs(run)
    cwd = os.getcwd()
    os.chdir(run)
    try:
        for ep in range(epochs):
            env.reset()
            env.save_base_station_positions(ep)
            for s in range(steps):
                env.step(ep, s)
            env.save_epoch_data(ep)
    finally:
        os.chdir(cwd)
    dst = os.path.join(OUT, "dumps", "custom")
    shutil.rmtree(dst, ignore_errors=True)
    for sub in ("collectData", "collectData2"):
        shutil.copytree(os.path.join(tmp, sub), os.path.join(dst, sub))
    shutil.rmtree(tmp)
    print("dumps custom", sum(len(f) for _, _, f in os.walk(dst)), "files")


def isoline_golden():
    """Coverage outlines from the reference's ``Channel.isoline`` (channels.py:30-75) for the default
    BS / UE parameters (base.py:117-123); rays that raise in the reference are recorded by exception
    name.  Target of ``mobile_env_gan_b200.core.channels.Channel.isoline``."""
    rh.import_reference()
    from mobile_env.core.channels import OkumuraHata
    from mobile_env.core.entities import BaseStation

    ue_cfg = dict(velocity=1.5, snr_tr=2e-8, noise=1e-9, height=1.6)
    cases = []
    for pos in [(110, 130), (65, 80), (20, 190), (199, 1), (100.5, 77.25), (0, 0)]:
        for thr in (0.0, 5.0, 30.0, 1e9):
            for num in (32, 17):
                case = {"pos": list(pos), "dthresh": thr, "num": num, "bounds": [200, 200]}
                try:
                    with np.errstate(all="ignore"):
                        xs, ys = OkumuraHata().isoline(BaseStation(0, pos, 9e6, 2500, 40, 50), ue_cfg, (200, 200), thr, num)
                    case["xs"], case["ys"] = [float(v) for v in xs], [float(v) for v in ys]
                except Exception as exc:  # noqa: BLE001 - the exception type is the recorded behaviour
                    case["raises"] = type(exc).__name__
                cases.append(case)
    with open(os.path.join(OUT, "isoline.json"), "w") as f:
        json.dump({"ue_config": ue_cfg, "bs": {"bw": 9e6, "freq": 2500, "tx": 40, "height": 50}, "cases": cases}, f)
    print("isoline:", len(cases), "cases,", sum("raises" in c for c in cases), "raising")


if __name__ == "__main__":
    main()
    dump_golden("kat1")
    custom_epochs_golden()
    custom_dump_golden()
    isoline_golden()
    random_golden()
    gym_pieces_golden()
