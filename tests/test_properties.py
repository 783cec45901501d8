"""Property tests (hypothesis) of the host logic and the oracle, plus the frozen GYM spec vectors."""
import json
import math
import os

import numpy as np
import pytest
from hypothesis import given, settings, strategies as st

from conftest import GOLDEN_DIR
from oracle import mbe_oracle as orc
from oracle.gen_gym_spec_vectors import CASES, run
from mobile_env_gan_b200.core.channels import LogDistance, OkumuraHata
from mobile_env_gan_b200.core.entities import BaseStation, UserEquipment
from mobile_env_gan_b200.core.movement import RandomWaypointMovement
from mobile_env_gan_b200.sharding import shard_envs


@pytest.mark.parametrize("name", sorted(CASES))
def test_gym_spec_vectors_are_stable(name):
    """The GYM-mode spec (parity unpinned) must not drift: frozen outputs of the scalar oracle."""
    with open(os.path.join(GOLDEN_DIR, f"gymspec_{name}.json")) as f:
        frozen = json.load(f)
    now = json.loads(json.dumps(run(CASES[name])))
    assert now["reset_obs"] == frozen["reset_obs"]
    for a, b in zip(now["steps"], frozen["steps"]):
        assert a["conn"] == b["conn"] and a["pos"] == b["pos"] and a["done"] == b["done"]
        assert a["rate"] == b["rate"]
        np.testing.assert_allclose(a["obs"], b["obs"], rtol=1e-12, atol=1e-15)
        np.testing.assert_allclose(a["reward"], b["reward"], rtol=1e-12, atol=1e-15)


@settings(max_examples=25, deadline=None)
@given(tx=st.floats(10, 50), h=st.floats(1.0, 3.0), freq=st.floats(800, 3500), hb=st.floats(20, 80))
def test_fold_is_the_exact_threshold_of_the_scalar_chain(tx, h, freq, hb):
    bs = BaseStation(0, (0, 0), 9e6, freq, tx, hb)
    ue = UserEquipment(0, 1.5, 2e-8, 1e-9, h)
    f = OkumuraHata().fold(bs, ue, 80000)
    p = orc.Params(tx=tx, ue_height=h, freq=freq, bs_height=hb)
    d2max = f["d2max"]
    if d2max >= 0:
        assert orc.snr_of(p, math.sqrt(d2max)) > p.snr_tr
        assert len(f["rate_lut"]) == d2max + 1 and f["rate_lut"][d2max] > 0
    if d2max < 80000:
        assert not (orc.snr_of(p, math.sqrt(d2max + 1)) > p.snr_tr)


@settings(max_examples=40, deadline=None)
@given(v=st.floats(0.05, 400.0))
def test_move_threshold_is_exact(v):
    n = RandomWaypointMovement(width=200, height=200, seed=1, reset_rng_episode=True).device_params(v)["move_d2max"]
    assert n < 0 or math.sqrt(n) <= v
    assert not (math.sqrt(n + 1) <= v)


@settings(max_examples=40, deadline=None)
@given(total=st.integers(1, 10**6), world=st.integers(1, 64))
def test_shards_tile_the_env_range(total, world):
    spans = [shard_envs(total, r, world) for r in range(world)]
    assert spans[0][0] == 0 and spans[-1][0] + spans[-1][1] == total
    assert all(a[0] + a[1] == b[0] for a, b in zip(spans, spans[1:]))


@settings(max_examples=15, deadline=None)
@given(seed=st.integers(0, 10**6), U=st.integers(1, 9), B=st.integers(1, 6), v=st.sampled_from([0.5, 1.5, 3.0, 10.0]),
       sched=st.sampled_from(["resource_fair", "proportional_fair", "rate_fair"]))
def test_scalar_and_batch_oracles_agree_fork(seed, U, B, v, sched):
    rng = np.random.default_rng(seed)
    p = orc.Params(velocity=v, ep_time=6, scheduler=sched, tx=float(rng.choice([30, 40, 46])))
    bs = rng.integers(0, 200, size=(B, 2))
    init = rng.integers(0, 200, size=(1, U, 2))
    env = orc.ScalarEnv(p, bs.tolist(), U)
    env.reset(init[0].tolist())
    pos, wp, t = init.copy(), np.full_like(init, -1), np.zeros(1, dtype=np.int64)
    for k in range(6):
        new_wp = rng.integers(0, 200, size=(1, U, 2))
        env.wp_source = lambda u, kk: new_wp[0, u]
        out_s = env.step_fork()
        out_b = orc.batch_step_fork(p, pos, wp, new_wp, bs, t)
        pos, wp, t = out_b["pos"], out_b["wp"], out_b["t"]
        assert [list(q) for q in out_s["pos"]] == out_b["pos"][0].tolist()
        assert out_s["assoc"] == out_b["assoc"][0].tolist()
        assert out_s["rate"] == out_b["rate"][0].tolist()
        np.testing.assert_allclose(out_s["utility"], out_b["utility"][0], rtol=1e-12, atol=1e-15)


def test_log_distance_channel_folds_like_its_formula():
    bs = BaseStation(0, (0, 0), 9e6, 2500, 40, 50)
    ue = UserEquipment(0, 1.5, 2e-8, 1e-9, 1.6)
    ch = LogDistance(a=125.0, c=40.0)
    f = ch.fold(bs, ue, 80000)
    d = math.sqrt(f["d2max"])
    assert 10 ** ((40 - (125 + 40 * math.log10(d))) / 10) / 1e-9 > 2e-8
    d = math.sqrt(f["d2max"] + 1)
    assert not (10 ** ((40 - (125 + 40 * math.log10(d))) / 10) / 1e-9 > 2e-8)
