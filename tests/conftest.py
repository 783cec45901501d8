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


def golden_names():
    return sorted(f[len("fork_"):-len(".json")] for f in os.listdir(GOLDEN_DIR) if f.startswith("fork_"))


def load_golden(name):
    with open(os.path.join(GOLDEN_DIR, f"fork_{name}.json")) as f:
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
