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
    if rec.get("ue_over") or rec["params"].get("channel"):
        pytest.skip("the compiled restatement models OkumuraHata and one UE class; the Python oracle covers this case")
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
    if rec.get("ue_over") or rec["params"].get("channel"):
        pytest.skip("the compiled restatement models OkumuraHata and one UE class; the Python oracle covers this case")
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


@pytest.mark.parametrize("name", ["central_rf", "ma_rf", "ma_pf"])
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
    wide = cpu_baseline.run_compiled("mobile-synthetic-central-v0", 0.2)  # 64 x 512 ProportionalFair
    assert wide["value"] > 0 and "128 envs" in wide["sample"]  # env count bounded for the wide shape
    assert cpu_baseline.run_compiled("no-such-workload-v0", 0.2) is None  # not in the baseline's workload table
    bad = cpu_baseline.run_compiled("mobile-small-central-v0", 0.2, envs=-5)
    assert "unavailable" in bad


from hypothesis import given, settings, strategies as st  # noqa: E402


@settings(max_examples=20, deadline=None)
@given(seed=st.integers(0, 10**6), U=st.integers(1, 12), B=st.integers(1, 8),
       v=st.sampled_from([0.5, 1.5, 2.5, 7.3, 10.0, 40.0]), handler=st.sampled_from(["central", "ma"]),
       sched=st.sampled_from(["resource_fair", "proportional_fair", "rate_fair"]))
def test_c_and_python_oracles_agree_on_random_gym_episodes(seed, U, B, v, handler, sched):
    """Two independent restatements (compiled C / scalar Python) on random scenarios: GYM step incl.
    observations and rewards."""
    from oracle.c_oracle import CEnvBatch

    rng = np.random.default_rng(seed)
    p = orc.Params(velocity=v, ep_time=7, tx=float(rng.choice([30, 40, 46])), snr_tr=float(rng.choice([2e-8, 2e-7])),
                   scheduler=sched)
    bs = rng.integers(0, 200, size=(B, 2)).tolist()
    init = rng.integers(0, 200, size=(U, 2)).tolist()
    wps = rng.integers(0, 200, size=(U, 16, 2)).tolist()
    py = orc.ScalarEnv(p, bs, U, wp_source=lambda u, k: wps[u][k])
    py.reset(init)
    c = CEnvBatch(p, bs, 2, U, handler=handler)
    c.reset(init)
    used = [0] * U
    for _ in range(7):
        acts = rng.integers(0, B + 1, size=U).tolist()
        obs, rew, done, info = py.step_gym(acts, handler)
        c.step_gym(acts, [wps[u][used[u]] for u in range(U)])
        used = [used[u] + int(c.drew[0, u]) for u in range(U)]
        for e in range(2):
            want_conn = [[] for _ in range(U)] if done else info["conn"]
            assert [np.flatnonzero(x).tolist() for x in c.conn[e]] == want_conn
            assert c.pos[e].tolist() == [list(q) for q in info["pos"]] and bool(c.done[e]) == done
            np.testing.assert_allclose(c.rate[e], info["rate"], rtol=1e-12, atol=0)
            np.testing.assert_allclose(c.util[e], info["utility"], rtol=1e-12, atol=1e-15)
            np.testing.assert_allclose(c.bs_util[e], info["bs_utility"], rtol=1e-12, atol=1e-15)
            np.testing.assert_allclose(c.obs[e], obs, rtol=1e-6, atol=1e-7)
            np.testing.assert_allclose(np.ravel(c.reward[e]), np.ravel(rew), rtol=1e-12, atol=1e-15)


@settings(max_examples=15, deadline=None)
@given(seed=st.integers(0, 10**6), U=st.integers(1, 9), v=st.sampled_from([1.5, 4.0, 10.0, 25.0]))
def test_c_and_python_oracles_agree_on_random_fork_layouts(seed, U, v):
    """FORK step on per-env layouts with 1..10 live BS slots (the fork's MComCustom shape)."""
    from oracle.c_oracle import CEnvBatch

    rng = np.random.default_rng(seed)
    E, B = 4, 10
    p = orc.Params(velocity=v, ep_time=6)
    nbs = rng.integers(1, B + 1, size=E)
    layout = rng.integers(0, 200, size=(E, B, 2))
    init = rng.integers(0, 200, size=(E, U, 2))
    c = CEnvBatch(p, layout, E, U, nbs=nbs)
    c.reset(init)
    pys = []
    for e in range(E):
        env = orc.ScalarEnv(p, layout[e, : nbs[e]].tolist(), U)
        env.reset(init[e].tolist())
        pys.append(env)
    for _ in range(6):
        new_wp = rng.integers(0, 200, size=(E, U, 2))
        c.step_fork(new_wp)
        for e, env in enumerate(pys):
            env.wp_source = lambda u, k, e=e: new_wp[e, u]
            out = env.step_fork()
            assert c.assoc[e].tolist() == out["assoc"] and c.pos[e].tolist() == [list(q) for q in out["pos"]]
            np.testing.assert_allclose(c.rate[e], out["rate"], rtol=1e-12, atol=0)
            np.testing.assert_allclose(c.util[e], out["utility"], rtol=1e-12, atol=1e-15)
            assert bool(c.done[e]) == out["done"] and c.metrics[e][1] == out["n_connected"]


@pytest.mark.parametrize("sched", ["resource_fair", "proportional_fair"])
def test_pinned_tables_make_the_compiled_rates_the_numpy_rates(sched):
    """``CEnvBatch(pinned_tables=True)`` feeds the C steps the snr / Shannon-rate tables of the numpy scalar
    chain (oracle/mbe_oracle.py snr_of / datarate_of -- the reference's operation order, pinned by the golden
    vectors); every FP64 rate of a 4,096-env GYM episode then equals the vectorised numpy oracle's bit for
    bit, which is what lets the full-size GPU tests demand bit equality against the compiled restatement."""
    from oracle.c_oracle import CEnvBatch

    p = orc.Params(scheduler=sched)
    bs = [(50, 50), (150, 50), (50, 150), (150, 150)]
    E, U = 4096, 15
    c = CEnvBatch(p, bs, E, U, handler="ma", pinned_tables=True)
    rng = np.random.default_rng(5)
    pos = rng.integers(0, 200, size=(E, U, 2))
    c.reset(pos)
    conn, wp, t = np.zeros((E, U, 4), dtype=bool), np.full((E, U, 2), -1), np.zeros(E, dtype=np.int64)
    for k in range(20):
        acts = rng.integers(0, 5, size=(E, U)).astype(np.int32)
        new_wp = rng.integers(0, 200, size=(E, U, 2))
        c.step_gym(acts, new_wp)
        out = orc.batch_step_gym(p, pos, wp, new_wp, np.array(bs), conn, acts, t, handler="ma")
        pos, wp, conn, t = out["pos"], out["wp"], out["conn"], out["t"]
        assert np.array_equal(out["conn"], c.conn.astype(bool)) and np.array_equal(out["pos"], c.pos), k
        assert np.array_equal(out["rate"], c.rate), (k, int((out["rate"] != c.rate).sum()))
    assert float(c.rate.max()) > 0
