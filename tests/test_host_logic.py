"""CPU-side checks: host folding vs the oracle's scalar chain, config/plugin surface, the
C-ABI library (loads, exports every symbol of include/mbe.h; no compute without a GPU) and the
multi-rank sharding logic on gloo (world_size 2)."""
import ctypes
import math
import os
import re

import numpy as np
import pytest
import torch

from oracle import mbe_oracle as orc
from mobile_env_gan_b200 import _lib
from mobile_env_gan_b200.core.base import MComCore
from mobile_env_gan_b200.core.channels import LogDistance, OkumuraHata
from mobile_env_gan_b200.core.entities import BaseStation, UserEquipment
from mobile_env_gan_b200.core.movement import Movement, RandomWaypointMovement
from mobile_env_gan_b200.core.util import deep_dict_merge
from mobile_env_gan_b200.sharding import shard_envs

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def entities(tx=40.0, h=1.6, v=1.5):
    bs = BaseStation(0, (10, 20), 9e6, 2500, tx, 50)
    ue = UserEquipment(0, v, 2e-8, 1e-9, h)
    return bs, ue


@pytest.mark.parametrize("tx,h", [(40.0, 1.6), (30.0, 1.8), (30.0, 1.5), (46.0, 1.6), (5.0, 1.6)])
def test_fold_matches_oracle_scalar_chain(tx, h):
    bs, ue = entities(tx, h)
    f = OkumuraHata().fold(bs, ue, 200 * 200 * 2)
    p = orc.Params(tx=tx, ue_height=h)
    d2max = f["d2max"]
    if d2max >= 0:
        assert orc.snr_of(p, math.sqrt(d2max)) > p.snr_tr
    assert not (orc.snr_of(p, math.sqrt(d2max + 1)) > p.snr_tr)
    # the rate table is the reference's Channel.datarate, bit for bit
    for d2 in list(range(0, min(d2max, 50) + 1)) + list(range(max(d2max - 50, 0), d2max + 1)):
        assert f["rate_lut"][d2] == orc.datarate_of(p, orc.snr_of(p, math.sqrt(d2)))
    # log-domain constants reproduce the FP64 SNR
    for d2 in (1, 2, 100, 5000, 19362, 79999):
        want = math.log2(orc.snr_of(p, math.sqrt(d2)))
        assert f["l0"] - f["k"] * math.log2(d2) == pytest.approx(want, abs=1e-9)
    assert f["l_zero"] == pytest.approx(math.log2(orc.snr_of(p, 0.0)), abs=1e-9)


def test_default_cutoff_is_the_surveyed_one():
    bs, ue = entities()
    assert OkumuraHata().fold(bs, ue, 80002)["d2max"] == 19362


def test_channel_scalar_surface_matches_reference_formulas():
    bs, ue = entities(30.0, 1.8)
    ue.x, ue.y = 81.9, 109.2  # truncated to (81, 109) like entities.py:52-54
    ch = OkumuraHata()
    p = orc.Params(tx=30.0, ue_height=1.8)
    d = orc.int_point_dist(10, 20, 81.9, 109.2)
    assert ch.power_loss(bs, ue) == orc.power_loss(p, d)
    assert ch.calculateSNR(bs, ue) == orc.snr_of(p, d)
    assert ch.datarate(bs, ue, 1e-6) == orc.datarate_of(p, 1e-6) and ch.datarate(bs, ue, 1e-9) == 0.0
    assert LogDistance(a=40, c=30).fold(bs, ue, 80000)["k"] == pytest.approx(1.5)


@pytest.mark.parametrize("v,want", [(10, 100), (1.5, 2), (0.5, 0), (1.0, 1), (7.5, 56), (2 ** 0.5, 2)])
def test_move_threshold(v, want):
    m = RandomWaypointMovement(width=200, height=200, seed=1, reset_rng_episode=True)
    got = m.device_params(v)["move_d2max"]
    assert got == want
    assert math.sqrt(got) <= v < math.sqrt(got + 1) or (got == 0 and v < 1)


def test_config_merge_seeding_and_plan():
    cfg = MComCore.default_config()
    assert cfg["bs"] == {"bw": 9e6, "freq": 2500, "tx": 40, "height": 50}  # base.py:117
    assert cfg["ue"] == {"velocity": 1.5, "snr_tr": 2e-8, "noise": 1e-9, "height": 1.6}
    cfg = deep_dict_merge(cfg, {"ue": {"velocity": 10}, "num_envs": 8, "mode": "gym", "handler": "ma"})
    cfg = MComCore.seeding(cfg)
    assert cfg["movement_params"]["seed"] == 2028 and cfg["arrival_params"]["seed"] == 2025  # base.py:155-170
    stations = [BaseStation(i, (10 * i, 5), **cfg["bs"]) for i in range(4)]
    stations[2].tx_power = 30  # second radio class
    users = [UserEquipment(i, **cfg["ue"]) for i in range(15)]
    plan = MComCore.build_plan(stations, users, cfg)
    assert (plan.num_envs, plan.num_ues, plan.num_bs, plan.feature_size) == (8, 15, 4, 17)
    assert plan.seed == 2028 and plan.ep_time == 20 and plan.move_d2max == 100
    assert plan.bs_class.tolist() == [0, 0, 1, 0] and len(plan.classes) == 2
    assert plan.classes[0]["d2max"] == 19362 and plan.classes[1]["d2max"] < 19362
    assert plan.bs_xy.tolist() == [[0, 5], [10, 5], [20, 5], [30, 5]]


def test_unsupported_plugins_fail_loudly():
    from mobile_env_gan_b200.core.schedules import Scheduler

    class Lottery(Scheduler):
        def share(self, bs, rates):
            return rates

    cfg = MComCore.seeding(deep_dict_merge(MComCore.default_config(), {"scheduler": Lottery}))
    users = [UserEquipment(0, **cfg["ue"])]
    stations = [BaseStation(0, (1, 1), **cfg["bs"])]
    with pytest.raises(NotImplementedError):
        MComCore.build_plan(stations, users, cfg)

    class Teleport(Movement):
        def move(self, ue):
            return 0, 0

    cfg = MComCore.seeding(deep_dict_merge(MComCore.default_config(), {"movement": Teleport}))
    with pytest.raises(NotImplementedError):
        MComCore.build_plan(stations, users, cfg)
    # heterogeneous UEs (entities.py:32-57 keeps velocity / snr_tr / noise / height per UE) become UE classes
    users2 = users + [UserEquipment(1, velocity=3, snr_tr=2e-8, noise=1e-9, height=1.6),
                      UserEquipment(2, velocity=3, snr_tr=1e-7, noise=1e-9, height=1.6),
                      UserEquipment(3, **cfg["ue"])]
    cfg = MComCore.seeding(MComCore.default_config())
    plan = MComCore.build_plan(stations, users2, cfg)
    assert plan.ue_class.tolist() == [0, 1, 2, 0] and len(plan.ue_classes) == 3 and len(plan.classes) == 3
    assert [c["velocity"] for c in plan.ue_classes] == [1.5, 3.0, 3.0]
    assert plan.classes[0] is plan.classes[1]  # same radio parameters, other speed: one fold
    assert plan.classes[2]["d2max"] < plan.classes[0]["d2max"]  # higher threshold: shorter range
    many = [UserEquipment(i, velocity=1.0 + i, snr_tr=2e-8, noise=1e-9, height=1.6) for i in range(9)]
    with pytest.raises(NotImplementedError):
        MComCore.build_plan(stations, many, cfg)


def test_library_exports_every_declared_symbol():
    """include/mbe.h is the contract: each prototype must be exported by libmbe.so."""
    header = open(os.path.join(ROOT, "include", "mbe.h")).read()
    declared = set(re.findall(r"\b(mbe_[a-z_]+)\s*\(", header))
    declared -= {"mbe_create"} - {"mbe_create"}  # keep all
    assert {name for name, _, _ in _lib.SYMBOLS} == declared
    lib = _lib.load()
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.mbe_abi_version() == _lib.MBE_ABI_VERSION
    assert b"sm_100a" in lib.mbe_build_info()
    # struct sizes the binding assumes (LP64)
    assert ctypes.sizeof(_lib.BsClass) == 48
    assert ctypes.sizeof(_lib.Buffers) == 18 * 8 + 8
    # ... and the library reports the same layout (checked again by _lib.load on every import)
    for which, struct in enumerate((_lib.Config, _lib.Buffers, _lib.BsClass, _lib.UeClass, _lib.RolloutOut)):
        assert lib.mbe_struct_size(which) == ctypes.sizeof(struct), struct.__name__
    assert lib.mbe_struct_size(99) == -1


def test_library_rejects_bad_configs_without_a_gpu():
    lib = _lib.load()
    cfg = _lib.Config()
    handle = ctypes.c_void_p()
    assert lib.mbe_create(ctypes.byref(cfg), ctypes.byref(handle)) != 0
    assert b"abi_version" in lib.mbe_last_error()
    cfg.abi_version = _lib.MBE_ABI_VERSION
    cfg.num_envs, cfg.num_ues, cfg.num_bs = 4, 5000, 4
    assert lib.mbe_create(ctypes.byref(cfg), ctypes.byref(handle)) != 0
    assert b"num_ues" in lib.mbe_last_error()
    assert lib.mbe_step(None, None) != 0


@pytest.mark.skipif(torch.cuda.is_available(), reason="CPU-only behaviour")
def test_no_cpu_fallback():
    import mobile_env_gan_b200 as mbe

    with pytest.raises(_lib.MbeError):
        mbe.make("mobile-small-central-v0", num_envs=2)


def test_shard_envs_partitions_exactly():
    for total, world in [(65536, 8), (1000, 3), (7, 8), (262144, 4)]:
        spans = [shard_envs(total, r, world) for r in range(world)]
        assert spans[0][0] == 0 and sum(c for _, c in spans) == total
        for (o0, c0), (o1, _) in zip(spans, spans[1:]):
            assert o0 + c0 == o1
        assert max(c for _, c in spans) - min(c for _, c in spans) <= 1


def _gloo_worker(rank, world, port, total, q):
    import torch.distributed as dist

    from mobile_env_gan_b200.sharding import gather_episode_stats, sharded_config

    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    cfg = sharded_config(total)
    ids = torch.arange(cfg["env_offset"], cfg["env_offset"] + cfg["num_envs"], dtype=torch.float32)
    stats = torch.stack([ids, ids * 2], dim=1)
    full = gather_episode_stats(stats, total)
    q.put((rank, cfg, full.tolist()))
    dist.barrier()
    dist.destroy_process_group()


def test_gloo_two_ranks_shard_and_gather():
    import torch.multiprocessing as mp

    total, world = 11, 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_gloo_worker, args=(r, world, port, total, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    want = [[float(i), float(2 * i)] for i in range(total)]
    offsets = sorted(cfg["env_offset"] for _, cfg, _ in got)
    assert offsets == [0, 6]
    for _, _, full in got:
        assert full == want


def test_handler_spaces_without_a_gpu():
    """Spaces only depend on the scenario sizes (upstream shapes: MultiDiscrete([B+1]*U), Box(U*(2B+1)))."""
    from mobile_env_gan_b200.handlers import MComCentralHandler, MComMAHandler

    class Sizes:
        NUM_STATIONS, NUM_USERS, userDict = 4, 15, {i: None for i in range(15)}

    a = MComCentralHandler.action_space(Sizes)
    o = MComCentralHandler.observation_space(Sizes)
    assert a.shape == (15,) and int(a.nvec[0]) == 5 and o.shape == (15 * 9,)
    am, om = MComMAHandler.action_space(Sizes), MComMAHandler.observation_space(Sizes)
    assert len(am.spaces) == 15 and am.spaces[0].n == 5 and om.spaces[3].shape == (17,)
    cfg = MComCore.seeding(deep_dict_merge(MComCore.default_config(), {"mode": "gym", "handler": MComMAHandler}))
    stations = [BaseStation(0, (1, 1), **cfg["bs"])]
    users = [UserEquipment(0, **cfg["ue"])]
    assert MComCore.build_plan(stations, users, cfg).handler == _lib.HANDLER_MA


def test_isoline_matches_reference_golden():
    """Channel.isoline / boundary_collison (reference channels.py:30-75, 86-127) against outlines the
    reference produced in the build container (oracle/gen_golden.py:isoline_golden), including the
    rays on which the reference raises."""
    import json
    import os

    import numpy as np

    from mobile_env_gan_b200.core.channels import OkumuraHata
    from mobile_env_gan_b200.core.entities import BaseStation

    with open(os.path.join(os.path.dirname(__file__), "golden", "isoline.json")) as f:
        gold = json.load(f)
    ok = 0
    for case in gold["cases"]:
        bs = BaseStation(0, tuple(case["pos"]), **gold["bs"])
        call = lambda: OkumuraHata().isoline(bs, gold["ue_config"], tuple(case["bounds"]), case["dthresh"], case["num"])
        with np.errstate(all="ignore"):
            if "raises" in case:
                try:
                    call()
                except Exception as exc:  # noqa: BLE001
                    assert type(exc).__name__ == case["raises"]
                else:
                    raise AssertionError(f"expected {case['raises']} for {case}")
            else:
                xs, ys = call()
                assert list(map(float, xs)) == case["xs"] and list(map(float, ys)) == case["ys"]
                ok += 1
    assert ok >= 16


def test_monitor_frames_match_reference_monitor():
    """Monitor.load_results / info (reference logging.py:6-87): same frames as the reference's Monitor
    fed the same per-step values (env 3 of a batch of 5)."""
    import pandas as pd
    import torch

    from oracle import ref_harness

    if not ref_harness.reference_available():
        pytest.skip("the reference is only importable in the build container")
    ref_harness.import_reference()
    from mobile_env.core.logging import Monitor as RefMonitor

    from mobile_env_gan_b200.core.logging import Monitor

    E, U, B, T, env = 5, 4, 3, 6, 3
    gen = torch.Generator().manual_seed(0)
    scal = [torch.rand(E, generator=gen, dtype=torch.float64) for _ in range(T)]
    ue = [torch.rand(E, U, generator=gen, dtype=torch.float64) for _ in range(T)]
    bs = [torch.rand(E, B, generator=gen, dtype=torch.float64) for _ in range(T)]
    clock = {"t": 0}
    mine = Monitor({"mean utility": lambda sim: scal[clock["t"]]}, {"rate": lambda sim: ue[clock["t"]],
                                                                    "qoe": lambda sim: 2 * ue[clock["t"]]},
                   {"load": lambda sim: bs[clock["t"]]})
    ref = RefMonitor({"mean utility": lambda sim: float(scal[clock["t"]][env])},
                     {"rate": lambda sim: dict(enumerate(ue[clock["t"]][env].tolist())),
                      "qoe": lambda sim: dict(enumerate((2 * ue[clock["t"]][env]).tolist()))},
                     {"load": lambda sim: dict(enumerate(bs[clock["t"]][env].tolist()))})
    mine.reset(), ref.reset()
    assert mine.info(env) == {} and ref.info() == {}
    for t in range(T):
        clock["t"] = t
        mine.update(None), ref.update(None)
    got, want = mine.load_results(env), ref.load_results()
    for g, w in zip(got, want):
        pd.testing.assert_frame_equal(g, w[sorted(w.columns)] if w.columns.name == "Metric" else w, check_dtype=False)
    gi, wi = mine.info(env), ref.info()
    assert gi["mean utility"] == wi["mean utility"]
    assert gi["rate"] == [wi["rate"][i] for i in range(U)] and gi["load"] == [wi["load"][i] for i in range(B)]


def test_shared_trajectory_follows_the_forks_movement_flag():
    """MComCustom mirrors the fork: movement_params.reset_rng_episode=True (the reference default,
    base.py:130-134) makes every epoch replay one UE trajectory, so with env index = epoch number all
    envs share the trajectory draws; the Gymnasium-shaped scenarios keep independent envs."""
    from mobile_env_gan_b200.scenarios.custom import MComCustom

    def plan_of(cls, extra, stations=()):
        cfg = cls.seeding(deep_dict_merge(cls.default_config(), extra))
        users = [UserEquipment(i, **cfg["ue"]) for i in range(7)]
        return cls.build_plan(list(stations), users, cfg)

    p = plan_of(MComCustom, {"num_envs": 64})
    assert p.shared_trajectory and p.reset_rng_episode and p.bs_random == (5, 10) and p.num_bs == 10
    assert not plan_of(MComCustom, {"num_envs": 64, "movement_params": {"reset_rng_episode": False}}).shared_trajectory
    assert not plan_of(MComCustom, {"num_envs": 64, "shared_trajectory": False}).shared_trajectory
    bs = [BaseStation(0, (10, 10), **MComCore.default_config()["bs"])]
    assert not plan_of(MComCore, {"num_envs": 64}, bs).shared_trajectory
    assert plan_of(MComCore, {"num_envs": 64, "shared_trajectory": True}, bs).shared_trajectory


def test_env_view_queries_match_the_reference_methods():
    """EnvView.check_connectivity / available_connections / allocateDataRate2User / user_total_datarates /
    allStationUtilities / update_connections (reference base.py:212-227, 413-447) on a snapshot, against
    the reference's own methods on the same state (a GYM-order episode with UEs on several BSs)."""
    import types

    import torch

    from oracle import ref_harness

    if not ref_harness.reference_available():
        pytest.skip("the reference is only importable in the build container")
    from mobile_env_gan_b200.core.schedules import ResourceFair
    from mobile_env_gan_b200.core.utilities import BoundedLogUtility
    from mobile_env_gan_b200.core.views import EnvView

    bs_xy, U = [(50, 50), (150, 50), (50, 150), (150, 150)], 9
    ref = ref_harness.make_fixed_layout_env(bs_xy, U, config={"ue": {"velocity": 6}})
    rng = np.random.default_rng(4)
    actions = rng.integers(0, 5, size=(6, U)).tolist()
    ref_harness.record_gym_pieces_episode(ref, actions)  # leaves the reference env in its final state
    r_bss = [ref.stationDict[k] for k in sorted(ref.stationDict)]
    r_ues = [ref.userDict[k] for k in sorted(ref.userDict)]
    # the same state as a batched-env snapshot (CPU tensors are enough for the view)
    cfg = MComCore.seeding(deep_dict_merge(MComCore.default_config(), {"ue": {"velocity": 6}, "mode": "gym"}))
    stations = [BaseStation(i, xy, **cfg["bs"]) for i, xy in enumerate(bs_xy)]
    users = [UserEquipment(i, **cfg["ue"]) for i in range(U)]
    conn = torch.tensor([[sum(1 << b.bs_id for b in r_bss if ue in ref.bs2ue_connections[b]) for ue in r_ues]])
    stub = types.SimpleNamespace(
        plan=MComCore.build_plan(stations, users, cfg), config=cfg,
        channelModel=OkumuraHata(), schedulerModel=ResourceFair(),
        utilityModel=BoundedLogUtility(**cfg["utility_params"]),
        pos=torch.tensor([[[int(ue.x), int(ue.y)] for ue in r_ues]]), bs_xy=torch.tensor(bs_xy), nbs=None,
        stationDict={b.bs_id: b for b in stations}, userDict={u.ue_id: u for u in users},
        t=torch.tensor([int(ref.time)]), rate=torch.tensor([[float(ref.allUserDataRates.get(ue, 0.0)) for ue in r_ues]]),
        utility_scaled=torch.tensor([[float(ref.ue_utilities[ue]) for ue in r_ues]], dtype=torch.float64),
        assoc=None, conn=conn)
    view = EnvView(stub, 0)
    v_bss, v_ues = [view.stationDict[i] for i in range(4)], [view.userDict[i] for i in range(U)]
    ids = lambda s: sorted(x.bs_id if hasattr(x, "bs_id") else x.ue_id for x in s)  # noqa: E731
    multi = 0
    for vu, ru in zip(v_ues, r_ues):
        assert ids(view.available_connections(vu)) == ids(ref.available_connections(ru))
        for vb, rb in zip(v_bss, r_bss):
            assert view.check_connectivity(vb, vu) == ref.check_connectivity(rb, ru)
        multi += sum(vu in view.bs2ue_connections[vb] for vb in v_bss) > 1
    assert multi > 0
    all_rates = {}
    for vb, rb in zip(v_bss, r_bss):
        got = {(b.bs_id, u.ue_id): r for (b, u), r in view.allocateDataRate2User(vb).items()}
        want = {(b.bs_id, u.ue_id): r for (b, u), r in ref.allocateDataRate2User(rb).items()}
        assert got == want
        all_rates.update(view.allocateDataRate2User(vb))
    got_tot = {u.ue_id: r for u, r in view.user_total_datarates(all_rates).items()}
    ref_rates = {}
    for rb in r_bss:
        ref_rates.update(ref.allocateDataRate2User(rb))
    assert got_tot == {u.ue_id: r for u, r in ref.user_total_datarates(ref_rates).items()}
    got_u = {b.bs_id: v for b, v in view.allStationUtilities().items()}
    want_u = {b.bs_id: v for b, v in ref.allStationUtilities().items()}
    assert got_u.keys() == want_u.keys() and all(got_u[k] == pytest.approx(want_u[k], rel=1e-12) for k in got_u)
    # move everybody far away: update_connections drops every link on both sides
    for vu, ru in zip(v_ues, r_ues):
        vu.x = ru.x = 5000
    view.update_connections(), ref.update_connections()
    assert all(len(view.bs2ue_connections[b]) == 0 for b in v_bss) and all(len(ref.bs2ue_connections[b]) == 0 for b in r_bss)


def test_reference_install_runs_the_unmodified_step(tmp_path):
    """oracle/build_ref.py installs the reference with its own setup.py into oracle/_ref (what the GPU
    box's CPU baseline executes); its files are the reference's, byte for byte, and the bench worker
    steps it (MComCore.step, base.py:230-296) on a scenario layout and as MComCustom with dumps on."""
    import filecmp

    from oracle import build_ref, cpu_baseline

    if not os.path.isdir(os.path.join(build_ref.REF_SRC, "mobile_env", "core")):
        pytest.skip("the reference sources only exist in the build container")
    out = build_ref.build_ref()
    assert out and build_ref.installed()
    for name in ("base.py", "channels.py", "movement.py", "schedules.py", "utilities.py", "entities.py"):
        assert filecmp.cmp(os.path.join(build_ref.REF_SRC, "mobile_env", "core", name),
                           os.path.join(out, "mobile_env", "core", name), shallow=False), name
    steps, wall = cpu_baseline._ref_worker(("mobile-medium-central-v0", 0.05, 3, False))
    assert steps >= 8 and wall > 0
    cwd = os.getcwd()
    try:
        steps, _ = cpu_baseline._ref_worker(("mobile-custom-v0", 0.05, 3, True))
    finally:
        os.chdir(cwd)
    assert steps >= 8


def test_scalar_movement_plugin_methods_match_the_reference_class():
    """Movement.reset / initial_position / move -- the per-entity plugin methods the reference env calls
    (base.py:189, 199, 233; movement.py:16-18, 42-72): same PCG64 stream, same positions, same bookkeeping
    dictionaries as the reference's RandomWaypointMovement for three UEs over 120 moves and a reset."""
    from oracle import ref_harness

    if not ref_harness.reference_available():
        pytest.skip("the reference is only importable in the build container")
    ref_harness.import_reference()
    from mobile_env.core.entities import UserEquipment as RefUE
    from mobile_env.core.movement import RandomWaypointMovement as RefMove

    from mobile_env_gan_b200.core.entities import UserEquipment
    from mobile_env_gan_b200.core.movement import RandomWaypointMovement

    for reset_rng in (True, False):
        kw = dict(width=200, height=160, seed=2028, reset_rng_episode=reset_rng)
        mine, ref = RandomWaypointMovement(**kw), RefMove(**kw)
        ues = [UserEquipment(i, velocity=v, snr_tr=2e-8, noise=1e-9, height=1.5) for i, v in enumerate((1.5, 10, 37.5))]
        rues = [RefUE(i, velocity=v, snr_tr=2e-8, noise=1e-9, height=1.5) for i, v in enumerate((1.5, 10, 37.5))]
        for episode in range(2):
            mine.reset(), ref.reset()
            for a, b in zip(ues, rues):
                a.x, a.y = mine.initial_position(a)
                b.x, b.y = ref.initial_position(b)
                assert (a.x, a.y) == (b.x, b.y) and mine.initial_position(a) == (a.x, a.y)
            for _ in range(60):
                for a, b in zip(ues, rues):
                    a.x, a.y = mine.move(a)
                    b.x, b.y = ref.move(b)
                    assert (int(a.x), int(a.y)) == (int(b.x), int(b.y))
                assert {u.ue_id: w for u, w in mine.userMoveDirection.items()} == \
                       {u.ue_id: w for u, w in ref.userMoveDirection.items()}
