"""TEST INFRASTRUCTURE ONLY -- freezes the GYM-mode specification (oracle/mbe_oracle.py ScalarEnv
.step_gym) into tests/golden/gymspec_*.json so that the build's own spec cannot drift unnoticed.
These are NOT reference vectors: the fork has no GYM step (parity unpinned).

    python oracle/gen_gym_spec_vectors.py
"""
from __future__ import annotations

import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from oracle import mbe_oracle as orc  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")
CASES = {
    "central_rf": dict(handler="central", scheduler="resource_fair", B=4, U=6, velocity=6.0, seed=1),
    "ma_rf": dict(handler="ma", scheduler="resource_fair", B=3, U=5, velocity=1.5, seed=2),
    "ma_pf": dict(handler="ma", scheduler="proportional_fair", B=5, U=7, velocity=10.0, seed=3),
}


def run(case):
    rng = np.random.default_rng(case["seed"])
    p = orc.Params(velocity=case["velocity"], ep_time=8, scheduler=case["scheduler"])
    B, U = case["B"], case["U"]
    bs = rng.integers(0, 200, size=(B, 2)).tolist()
    init = rng.integers(0, 200, size=(U, 2)).tolist()
    wps = rng.integers(0, 200, size=(U, 16, 2)).tolist()
    acts = rng.integers(0, B + 1, size=(8, U)).tolist()
    env = orc.ScalarEnv(p, bs, U, wp_source=lambda u, k: wps[u][k])
    env.reset(init)
    rec = {"case": case, "bs": bs, "init": init, "wps": wps, "acts": acts,
           "reset_obs": env.observe(case["handler"]).tolist(), "steps": []}
    for a in acts:
        obs, rew, done, info = env.step_gym(a, case["handler"])
        rec["steps"].append({"obs": obs.tolist(), "reward": rew if isinstance(rew, float) else list(map(float, rew)),
                             "done": bool(done), "conn": info["conn"], "rate": info["rate"],
                             "utility": info["utility"], "pos": [list(q) for q in info["pos"]]})
    return rec


if __name__ == "__main__":
    for name, case in CASES.items():
        with open(os.path.join(OUT, f"gymspec_{name}.json"), "w") as f:
            json.dump(run(case), f, separators=(",", ":"))
        print("wrote", name)
