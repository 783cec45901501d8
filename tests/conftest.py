import json
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    # the C-ABI library is built in-tree (git-ignored): make sure it exists / is current before any
    # test loads it (nvcc cross-compiles sm_100a without a GPU; on the GPU box the built file travels)
    try:
        from mobile_env_gan_b200.csrc.build import build

        build(force=False)
    except Exception as exc:  # no nvcc: the tests that need the library will say so
        print(f"[conftest] libmbe.so not rebuilt: {exc}")


def golden_names():
    return sorted(f[len("fork_"):-len(".json")] for f in os.listdir(GOLDEN_DIR) if f.startswith("fork_"))


def load_golden(name):
    with open(os.path.join(GOLDEN_DIR, f"fork_{name}.json")) as f:
        return json.load(f)


def gymref_names():
    return sorted(f[len("gymref_"):-len(".json")] for f in os.listdir(GOLDEN_DIR) if f.startswith("gymref_"))


def load_gymref(name):
    """GYM-order episodes executed by the reference's own primitives (oracle/ref_harness.py:
    record_gym_pieces_episode, fixtures written by oracle/gen_golden.py:gym_pieces_golden)."""
    with open(os.path.join(GOLDEN_DIR, f"gymref_{name}.json")) as f:
        return json.load(f)


def golden_waypoints(rec):
    """Per-UE list of waypoints in draw order, from the recorded (wx, wy, drew) triples."""
    nue = len(rec["init_pos"])
    seq = [[] for _ in range(nue)]
    for st in rec["steps"]:
        for u, (wx, wy, drew) in enumerate(st["wp"]):
            if drew:
                seq[u].append((wx, wy))
    return seq


@pytest.fixture(scope="session")
def golden_loader():
    return load_golden
