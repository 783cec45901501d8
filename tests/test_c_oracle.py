"""The compiled CPU restatement (oracle/mbe_oracle_c.c, test infrastructure / CPU baseline only) against
the same fixtures that pin the Python oracle: episodes of the unmodified reference (fork_*.json), GYM-order
episodes run by the reference's own primitives (gymref_*.json) and the frozen GYM spec vectors."""
import json
import os
import shutil

import numpy as np
import pytest

from conftest import GOLDEN_DIR, golden_names, gymref_names, load_golden, load_gymref
from oracle import mbe_oracle as orc

pytestmark = pytest.mark.skipif(shutil.which("gcc") is None, reason="needs gcc")


def params_of(rec):
    return orc.Params(**rec["params"])


@pytest.mark.parametrize("name", golden_names())
def test_c_fork_step_replays_reference_episodes(name):
    from oracle.c_oracle import CEnvBatch

    rec = load_golden(name)
    E, U = 3, len(rec["init_pos"])
    env = CEnvBatch(params_of(rec), rec["bs_xy"], E, U, bs_over=rec.get("bs_over"))
    env.reset(rec["init_pos"])
    for k, g in enumerate(rec["steps"]):
        env.step_fork([[w[0], w[1]] for w in g["wp"]])
        for e in (0, E - 1):
            assert env.pos[e].tolist() == g["pos"], (name, k)
            assert env.assoc[e].tolist() == g["conn"], (name, k)
            assert env.drew[e].tolist() == [w[2] for w in g["wp"]], (name, k)
            np.testing.assert_allclose(env.rate[e], g["rate"], rtol=1e-12, atol=0)
            np.testing.assert_allclose(env.util[e], g["utility"], rtol=1e-12, atol=1e-15)
            assert bool(env.done[e]) == g["done"]
            m = env.metrics[e]
            assert m[0] == g["n_connections"] and m[1] == g["n_connected"]
            assert m[2] == pytest.approx(g["mean_utility"], rel=1e-12, abs=1e-15)
            assert m[3] == pytest.approx(g["mean_datarate"], rel=1e-12)


@pytest.mark.parametrize("handler", ["central", "ma"])
@pytest.mark.parametrize("name", gymref_names())
def test_c_gym_step_replays_reference_primitive_episodes(name, handler):
    from oracle.c_oracle import CEnvBatch

    rec = load_gymref(name)
    E, U, B = 2, len(rec["init_pos"]), len(rec["bs_xy"])
    env = CEnvBatch(params_of(rec), rec["bs_xy"], E, U, handler=handler, bs_over=rec.get("bs_over"))
    env.reset(rec["init_pos"])
    for k, (acts, g) in enumerate(zip(rec["actions"], rec["steps"])):
        env.step_gym(acts, [[w[0], w[1]] for w in g["wp"]])
        for e in range(E):
            assert [np.flatnonzero(c).tolist() for c in env.conn[e]] == g["conn_after"], (name, k)
            assert env.pos[e].tolist() == g["pos"], (name, k)
            np.testing.assert_allclose(env.rate[e], g["rate"], rtol=1e-12, atol=0)
            np.testing.assert_allclose(env.util[e], g["utility"], rtol=1e-12, atol=1e-15)
            np.testing.assert_allclose(env.bs_util[e], g["bs_utility"], rtol=1e-12, atol=1e-15)
            assert bool(env.done[e]) == g["done"]
            assert env.metrics[e][0] == sum(len(c) for c in g["conn"])
            if handler == "central":
                assert env.reward[e] == pytest.approx(float(np.mean(g["utility"])), rel=1e-12, abs=1e-15)


@pytest.mark.parametrize("name", ["central_rf", "ma_rf"])
def test_c_gym_step_matches_the_frozen_spec_vectors(name):
    """Observations and rewards (this build's GYM spec, frozen from the Python oracle)."""
    from oracle.c_oracle import CEnvBatch

    with open(os.path.join(GOLDEN_DIR, f"gymspec_{name}.json")) as f:
        rec = json.load(f)
    case = rec["case"]
    p = orc.Params(velocity=case["velocity"], ep_time=8, scheduler=case["scheduler"])
    U, B = case["U"], case["B"]
    env = CEnvBatch(p, rec["bs"], 1, U, handler=case["handler"])
    env.reset(rec["init"])
    used = [0] * U
    for acts, g in zip(rec["acts"], rec["steps"]):
        env.step_gym(acts, [rec["wps"][u][used[u]] for u in range(U)])
        used = [used[u] + int(env.drew[0, u]) for u in range(U)]
        assert [np.flatnonzero(c).tolist() for c in env.conn[0]] == ([[] for _ in range(U)] if g["done"] else g["conn"])
        assert env.pos[0].tolist() == g["pos"] and bool(env.done[0]) == g["done"]
        np.testing.assert_allclose(env.rate[0], g["rate"], rtol=1e-12, atol=0)
        np.testing.assert_allclose(env.obs[0], np.asarray(g["obs"], dtype=np.float32), rtol=1e-6, atol=1e-7)
        np.testing.assert_allclose(np.ravel(env.reward[0]), np.ravel(g["reward"]), rtol=1e-12, atol=1e-15)


def test_compiled_cpu_baseline_runs_bounded():
    """bench.py's compiled CPU baseline leg: a bounded sample in a child process, never raises."""
    from oracle import cpu_baseline

    out = cpu_baseline.run_compiled("mobile-small-central-v0", 0.2, envs=256)
    assert out["value"] > 0 and out["kind"].startswith("port (compiled C") and out["cores"] >= 1
    assert cpu_baseline.run_compiled("mobile-synthetic-central-v0", 0.2) is None  # ProportionalFair: not covered
    bad = cpu_baseline.run_compiled("mobile-small-central-v0", 0.2, envs=-5)
    assert "unavailable" in bad
