"""Dump formats: mobile_env_gan_b200/export.py against files written by the
unmodified reference (tests/golden/dumps/kat1, produced by oracle/gen_golden.py:dump_golden with the
reference's own save_layout_and_data_rates / save_epoch_data / save_base_station_positions)."""
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN_DIR, golden_waypoints, load_golden
from oracle import mbe_oracle as orc

DUMPS = os.path.join(GOLDEN_DIR, "dumps", "kat1")


def read(rel, root=DUMPS):
    with open(os.path.join(root, rel)) as f:
        return f.read()


def check_against_reference_files(files, root=DUMPS):
    """files: {relative path: text}; root: the directory of files the reference wrote (default: epoch 0
    of the kat1 episode)."""
    n = 0
    for rel, text in files.items():
        ref = read(rel, root)
        if os.sep + "DataRate" + os.sep + "data_rates_" in rel:
            # the reference lists the UEs of a BS in Python-set order; compare order-insensitively
            key = lambda d: (d["bs_id"], d["ue_id"])  # noqa: E731
            assert sorted(json.loads(text), key=key) == sorted(json.loads(ref), key=key), rel
            assert text.count("\n") == ref.count("\n")
        else:
            assert text == ref, rel
        n += 1
    return n


def replay_kat1():
    rec = load_golden("kat1")
    p = orc.Params(**rec["params"])
    seq = golden_waypoints(rec)
    env = orc.ScalarEnv(p, rec["bs_xy"], len(rec["init_pos"]), wp_source=lambda u, k: seq[u][k])
    env.reset(rec["init_pos"])
    steps = []
    for _ in rec["steps"]:
        out = env.step_fork()
        steps.append({"pos": out["pos"], "arrived": [w is None for w in env.wp], "assoc": out["assoc"], "rate": out["rate"]})
    return rec, p, steps


def test_formatters_reproduce_reference_dump_files_byte_for_byte():
    from mobile_env_gan_b200.export import format_epoch_files, format_step_files

    rec, p, steps = replay_kat1()
    util = (p.util_lower, p.util_upper, tuple(p.util_coeffs))
    util = (int(util[0]), int(util[1]), tuple(int(c) for c in util[2]))  # reference config holds ints (base.py:136)
    n = 0
    for s, st in enumerate(steps):
        n += check_against_reference_files(format_step_files(0, s, rec["bs_xy"], st["pos"], st["assoc"], st["rate"], util))
    n += check_against_reference_files(format_epoch_files(
        0, rec["bs_xy"], [st["pos"] for st in steps], [st["arrived"] for st in steps],
        [st["assoc"] for st in steps], [st["rate"] for st in steps], util))
    assert n == 84  # every file the reference wrote for the episode


@pytest.mark.gpu
@pytest.mark.parametrize("how", ["steps", "rollout"])
def test_gpu_writer_reproduces_reference_dump_files(tmp_path, how):
    import torch

    from mobile_env_gan_b200.core.base import MComCore
    from mobile_env_gan_b200.core.entities import BaseStation, UserEquipment
    from mobile_env_gan_b200.core.util import deep_dict_merge
    from mobile_env_gan_b200.export import ReferenceDumpWriter

    rec = load_golden("kat1")
    pr = rec["params"]
    E, U = 6, len(rec["init_pos"])
    config = {"num_envs": E, "mode": "fork", "bs": {"tx": pr["tx"]}, "ue": {"velocity": pr["velocity"], "height": pr["ue_height"]}}
    cfg = deep_dict_merge(MComCore.default_config(), config)
    stations = [BaseStation(i, tuple(xy), **cfg["bs"]) for i, xy in enumerate(rec["bs_xy"])]
    users = [UserEquipment(i, **cfg["ue"]) for i in range(U)]
    env = MComCore(stations, users, config)
    seq = golden_waypoints(rec)
    K = max(len(s) for s in seq)
    wp = np.zeros((E, U, K, 2), dtype=np.int16)
    for u, s in enumerate(seq):
        for k, w in enumerate(s):
            wp[:, u, k] = w
    env.reset()
    env.inject_waypoints(wp)
    env.set_positions(np.broadcast_to(np.array(rec["init_pos"]), (E, U, 2)).copy())
    writer = ReferenceDumpWriter(env, str(tmp_path), envs=[0, 5])  # env index plays the epoch number
    writer.begin_episode()
    if how == "steps":
        for s in range(len(rec["steps"])):
            env.step(0, s)
            writer.after_step(s)
    else:  # the whole episode through mbe_rollout, files from its per-step series
        writer.write_rollout(env.rollout(len(rec["steps"]), record=("pos", "wp", "assoc", "rate")))
    writer.end_episode()
    writer.close()
    files = {}
    for dirpath, _, names in os.walk(tmp_path):
        for name in names:
            full = os.path.join(dirpath, name)
            files[os.path.relpath(full, tmp_path)] = open(full).read()
    assert len(files) == 2 * 84
    import re

    def epoch_of(rel):
        base = os.path.basename(rel)
        m = re.search(r"_(\d+)_(\d+)\.json$", base) if rel.startswith("collectData" + os.sep) else None
        if m:
            return int(m.group(1))
        return int(re.search(r"_(\d+)\.(csv|json)$", base).group(1))

    def with_epoch(rel, e):
        base = os.path.basename(rel)
        if rel.startswith("collectData" + os.sep):
            base = re.sub(r"_(\d+)_(\d+)\.json$", lambda m: f"_{e}_{m.group(2)}.json", base)
        else:
            base = re.sub(r"_(\d+)\.(csv|json)$", lambda m: f"_{e}.{m.group(2)}", base)
        return os.path.join(os.path.dirname(rel), base)

    epoch0 = {rel: text for rel, text in files.items() if epoch_of(rel) == 0}
    assert check_against_reference_files(epoch0) == 84
    # env 5 ran the same episode: identical contents under its own epoch number
    for rel, text in epoch0.items():
        assert files[with_epoch(rel, 5)] == text


@pytest.mark.gpu
def test_gpu_writer_rollout_equals_stepping_on_custom_scenario(tmp_path):
    """The fork's own scenario: dump files written from ONE fused-episode launch are the files the
    per-step path writes (which the test above pins to the reference's own files)."""
    from mobile_env_gan_b200.export import ReferenceDumpWriter
    from mobile_env_gan_b200.scenarios.custom import MComCustom

    def run(root, fused):
        env = MComCustom(config={"num_envs": 64})
        env.reset()
        writer = ReferenceDumpWriter(env, str(root), envs=[0, 33, 63])
        writer.begin_episode()
        if fused:
            before = env.launch_count
            writer.write_rollout(env.rollout(20, record=("pos", "wp", "assoc", "rate")))
            assert env.launch_count - before == 1
        else:
            for s in range(20):
                env.step(0, s)
                writer.after_step(s)
        writer.end_episode()
        writer.close()
        out = {}
        for dirpath, _, names in os.walk(root):
            for name in names:
                full = os.path.join(dirpath, name)
                out[os.path.relpath(full, root)] = open(full).read()
        return out

    a, b = run(tmp_path / "steps", False), run(tmp_path / "fused", True)
    assert len(a) == 3 * 84 and a == b


def test_env_view_writes_the_reference_step_files(tmp_path):
    """EnvView.save_layout_and_data_rates (the reference's method name, base.py:298-349) on snapshots of
    the kat1 episode: the same bytes as the files the reference wrote."""
    import types

    import torch

    from mobile_env_gan_b200.core.base import MComCore
    from mobile_env_gan_b200.core.channels import OkumuraHata
    from mobile_env_gan_b200.core.entities import BaseStation, UserEquipment
    from mobile_env_gan_b200.core.schedules import ResourceFair
    from mobile_env_gan_b200.core.util import deep_dict_merge
    from mobile_env_gan_b200.core.utilities import BoundedLogUtility
    from mobile_env_gan_b200.core.views import EnvView

    rec, p, steps = replay_kat1()
    pr = rec["params"]
    cfg = MComCore.seeding(deep_dict_merge(MComCore.default_config(), {
        "bs": {"tx": pr["tx"]}, "ue": {"velocity": pr["velocity"], "height": pr["ue_height"]}}))
    stations = [BaseStation(i, tuple(xy), **cfg["bs"]) for i, xy in enumerate(rec["bs_xy"])]
    users = [UserEquipment(i, **cfg["ue"]) for i in range(len(rec["init_pos"]))]
    plan = MComCore.build_plan(stations, users, cfg)
    n = 0
    for k in (0, 7, 19):
        st = steps[k]
        stub = types.SimpleNamespace(
            plan=plan, config=cfg, channelModel=OkumuraHata(), schedulerModel=ResourceFair(),
            utilityModel=BoundedLogUtility(**cfg["utility_params"]),
            pos=torch.tensor([st["pos"]]), bs_xy=torch.tensor(rec["bs_xy"]), nbs=None,
            stationDict={b.bs_id: b for b in stations}, userDict={u.ue_id: u for u in users},
            t=torch.tensor([k + 1]), rate=torch.tensor([st["rate"]], dtype=torch.float64),
            utility_scaled=torch.zeros(1, len(users), dtype=torch.float64),
            assoc=torch.tensor([st["assoc"]]), conn=None)
        written = EnvView(stub, 0).save_layout_and_data_rates(0, k, root=str(tmp_path))
        n += check_against_reference_files({rel: open(os.path.join(tmp_path, rel)).read() for rel in written})
    assert n == 12


def test_formatters_reproduce_the_shipped_collect_loop_files():
    """The fork's collect loop as shipped (MComCustom: a fresh 5..10-BS layout per epoch, dumps written
    from inside step / save_base_station_positions / save_epoch_data; tests/golden/dumps/custom from
    oracle/gen_golden.py:custom_dump_golden) for two epochs, from the oracle's replay of
    custom_epochs.json: every file byte for byte (168 files)."""
    from mobile_env_gan_b200.export import format_epoch_files, format_step_files

    root = os.path.join(GOLDEN_DIR, "dumps", "custom")
    with open(os.path.join(GOLDEN_DIR, "custom_epochs.json")) as f:
        epochs = json.load(f)["epochs"]
    p = orc.Params(velocity=10.0)  # MComCustom: custom.py:17
    util = (int(p.util_lower), int(p.util_upper), tuple(int(c) for c in p.util_coeffs))
    n = 0
    for e in (0, 1):
        rec = epochs[e]
        seq = golden_waypoints(rec)
        env = orc.ScalarEnv(p, rec["bs_xy"], len(rec["init_pos"]), wp_source=lambda u, k: seq[u][k])
        env.reset(rec["init_pos"])
        steps = []
        for s, g in enumerate(rec["steps"]):
            out = env.step_fork()
            assert [list(q) for q in out["pos"]] == g["pos"] and out["assoc"] == g["conn"]
            steps.append({"pos": out["pos"], "arrived": [w is None for w in env.wp], "assoc": out["assoc"], "rate": out["rate"]})
            n += check_against_reference_files(format_step_files(e, s, rec["bs_xy"], out["pos"], out["assoc"], out["rate"], util), root)
        n += check_against_reference_files(format_epoch_files(
            e, rec["bs_xy"], [st["pos"] for st in steps], [st["arrived"] for st in steps],
            [st["assoc"] for st in steps], [st["rate"] for st in steps], util), root)
    assert n == 168


@pytest.mark.gpu
def test_collect_notebook_loop_runs_verbatim(tmp_path, monkeypatch):
    """collectData2.ipynb cells 2-4 as written in the reference (env = MComCustom(render_mode="rgb_array");
    per epoch reset / save_base_station_positions / 20 x step(epoch, step) / save_epoch_data): the env-level
    methods of base.py:261,298-404 and custom.py:79-85 write into ../collectData and ../collectData2
    relative to the working directory, and the files are the ones export.ReferenceDumpWriter writes for an
    identical env (which other tests pin to the reference's own files)."""
    from mobile_env_gan_b200.export import ReferenceDumpWriter
    from mobile_env_gan_b200.scenarios.custom import MComCustom

    work = tmp_path / "run" / "notebooks"
    work.mkdir(parents=True)
    monkeypatch.chdir(work)

    env = MComCustom(render_mode="rgb_array")
    iteration_number = 2
    step_number = 20
    env.reset()
    for epoch_number in range(iteration_number):
        env.reset()
        env.save_base_station_positions(epoch_number)

        for curr_step in range(step_number):
            env.step(epoch_number, curr_step)

        env.save_epoch_data(epoch_number)

    # the reference's bookkeeping attributes (base.py:264-269) and the monitor call of base.py:272
    assert sorted(env.users_dataRateList) == list(range(7)) and len(env.users_dataRateList[0]) == step_number
    assert len(env.users_trajectoryList[3]) == step_number and len(env.userQoEList[6]) == step_number
    assert len(env.monitor.scalar_results["mean utility"]) == step_number
    assert set(env.monitor.info()) >= {"number connections", "number connected", "mean utility", "mean datarate"}

    def tree(root):
        out = {}
        for dirpath, _, names in os.walk(root):
            for name in names:
                full = os.path.join(dirpath, name)
                out[os.path.relpath(full, root)] = open(full).read()
        return out

    got = tree(tmp_path / "run")
    got = {rel: text for rel, text in got.items() if rel.startswith("collectData")}
    assert len(got) == iteration_number * 84

    twin = MComCustom(config={"dumps": False})
    twin.reset()
    for epoch_number in range(iteration_number):
        twin.reset()
        writer = ReferenceDumpWriter(twin, str(tmp_path / "twin"), envs=[0], epoch_offset=epoch_number)
        writer.begin_episode()
        for curr_step in range(step_number):
            twin.step(epoch_number, curr_step)
            writer.after_step(curr_step)
        writer.end_episode()
        writer.close()
    assert got == tree(tmp_path / "twin")
    # a batch keeps the fast path: no dumps, no per-step monitor clones unless asked for
    batch = MComCustom(config={"num_envs": 64})
    assert not batch.dumps and not batch.monitor_in_step
