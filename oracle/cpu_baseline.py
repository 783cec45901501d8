"""TEST/BENCH INFRASTRUCTURE ONLY -- times the reference's step on host cores (``cpu_baseline`` and
``--impl reference`` legs of bench.py): the UNMODIFIED reference where it is installed (``oracle/_ref``,
built by ``oracle/build_ref.py``; kind "reference"), the oracle's scalar port (``ScalarEnv``: per-entity
Python loops exactly like mobile_env/core/base.py:230-296, the reference's speed class; kind "port") and
the compiled C restatement (``oracle/mbe_oracle_c.c``, OpenMP).  The numbers are reported baselines, not
targets."""
from __future__ import annotations

import multiprocessing as mp
import os
import time

import numpy as np

from oracle import mbe_oracle as orc

_MEDIUM = [(50, 50), (150, 50), (50, 150), (150, 150)]
_LARGE = [(20 + 45 * (i % 4) + (22 if (i // 4) % 2 else 0), 25 + 50 * (i // 4)) for i in range(13)]
_SYNTH = [((i * 7919 + 13) % 800, (i * 104729 + 71) % 800) for i in range(64)]
_SYNTH_P = {"width": 800.0, "height": 800.0, "scheduler": "proportional_fair"}
WORKLOADS = {
    # name: (bs_xy, num_ues, mode, handler, velocity, Params overrides) -- the layouts of
    # mobile_env_gan_b200/scenarios/*.py
    "mobile-small-central-v0": ([(110, 130), (65, 80), (120, 30)], 5, "gym", "central", 1.5, {}),
    "mobile-small-ma-v0": ([(110, 130), (65, 80), (120, 30)], 5, "gym", "ma", 1.5, {}),
    "mobile-medium-central-v0": (_MEDIUM, 15, "gym", "central", 1.5, {}),
    "mobile-medium-ma-v0": (_MEDIUM, 15, "gym", "ma", 1.5, {}),
    "mobile-large-central-v0": (_LARGE, 30, "gym", "central", 1.5, {}),
    "mobile-large-ma-v0": (_LARGE, 30, "gym", "ma", 1.5, {}),
    # BASELINE.json configs[4]: 64 BS x 512 UE, ProportionalFair, 800 x 800 map
    "mobile-synthetic-central-v0": (_SYNTH, 512, "gym", "central", 1.5, _SYNTH_P),
    "mobile-synthetic-ma-v0": (_SYNTH, 512, "gym", "ma", 1.5, _SYNTH_P),
    # the fork's own scenario (custom.py): 7 UEs at velocity 10, 5..10 random BSs per episode, FORK step
    "mobile-custom-v0": (None, 7, "fork", "central", 10, {}),
}


def _worker(args):
    workload, seconds, seed = args
    bs, U, mode, handler, vel, over = WORKLOADS[workload]
    p = orc.Params(velocity=vel, **over)
    rng = np.random.default_rng(seed)
    random_layout = bs is None
    env = orc.ScalarEnv(p, bs or [(0, 0)], U,
                        wp_source=lambda u, k: (int(rng.uniform(0, p.width)), int(rng.uniform(0, p.height))))

    def fresh():
        if random_layout:  # generate_base_stations (custom.py:68-77)
            env.bs_xy = [(int(rng.uniform(0, p.width)), int(rng.uniform(0, p.height))) for _ in range(int(rng.integers(5, 11)))]
        env.reset([(int(rng.uniform(0, p.width)), int(rng.uniform(0, p.height))) for _ in range(U)])

    fresh()
    B = len(env.bs_xy)
    steps = 0
    t0 = time.perf_counter()
    while True:
        if mode == "gym":
            _, _, done, _ = env.step_gym(rng.integers(0, B + 1, size=U), handler)
        else:
            done = env.step_fork()["done"]
        steps += 1
        if done:
            fresh()
        if (steps % 8 == 0 or U > 64) and time.perf_counter() - t0 >= seconds:
            break
    return steps, time.perf_counter() - t0


def reference_installed() -> bool:
    """True where the unmodified reference can be executed: /root/reference (build container) or its
    pip install under oracle/_ref (oracle/build_ref.py; travels to the GPU box)."""
    from oracle import ref_harness

    return ref_harness.reference_available()


def _ref_worker(args):
    """The UNMODIFIED reference's MComCore.step (mobile_env/core/base.py:230-296) -- BASELINE.md section 3:
    a fixed-layout subclass harness for the scenario shapes (like MComCustom's custom.py:40-62), the
    fork's MComCustom itself for mobile-custom-v0; FORK semantics (the fork's step takes no action and
    builds no observation), default plugins, per-step JSON dumps disabled unless ``dumps``."""
    workload, seconds, seed, dumps = args
    import random
    import tempfile

    from oracle import ref_harness as rh

    bs, U, _mode, _handler, vel, over = WORKLOADS[workload]
    ep_time = 20
    random.seed(seed)
    scratch = None
    if bs is None:
        _, _, custom = rh.import_reference()
        cls = custom.MComCustom
        if dumps:  # the dump paths are relative ("../collectData2/..."): run inside a scratch directory
            scratch = tempfile.TemporaryDirectory(prefix="mbe_ref_dumps_")
            os.makedirs(os.path.join(scratch.name, "cwd"))
            os.chdir(os.path.join(scratch.name, "cwd"))
        else:
            cls = type("MComCustomNoDumps", (cls,), {"save_layout_and_data_rates": lambda self, e, s: None})
        env = cls({"seed": seed, "EP_MAX_TIME": ep_time})
    else:
        cfg = {"seed": seed, "EP_MAX_TIME": ep_time, "width": over.get("width", 200.0), "height": over.get("height", 200.0)}
        env = rh.make_fixed_layout_env(bs, U, config=cfg, ue_params={"velocity": vel})
    env.reset()
    steps = epoch = s = 0
    t0 = time.perf_counter()
    while True:
        env.step(epoch, s)
        steps += 1
        s += 1
        if env.time_is_up:
            epoch, s = epoch + 1, 0
            env.reset()
        if (steps % 8 == 0 or U > 64) and time.perf_counter() - t0 >= seconds:
            break
    wall = time.perf_counter() - t0
    if scratch is not None:
        os.chdir("/")
        scratch.cleanup()
    return steps, wall


def _summary(workload, procs, seconds, res, kind="port"):
    total = sum(s for s, _ in res)
    wall = max(t for _, t in res)
    what = ("scalar port of MComCore.step" if kind == "port" else
            "the unmodified reference's MComCore.step from oracle/_ref, FORK semantics (the fork has no actions / "
            "observations), default plugins, JSON dumps off")
    return {
        "value": total / wall,
        "unit": "env-steps/s",
        "cores": procs,
        "kind": kind,
        "single_core": float(np.mean([s / t for s, t in res])),
        "sample": f"{procs} procs x {seconds:.2f} s of {workload} ({what}, one env per process, {total} env-steps)",
    }


def run(workload: str, seconds: float = 5.0, procs: int | None = None):
    """Steps independent envs of ``workload`` on ``procs`` host processes for ``seconds`` each.
    Returns dict(value=env-steps/s aggregate, cores, single_core, sample)."""
    procs = procs or os.cpu_count() or 1
    ctx = mp.get_context("spawn")
    with ctx.Pool(procs) as pool:
        res = pool.map(_worker, [(workload, seconds, 1000 + i) for i in range(procs)])
    return _summary(workload, procs, seconds, res)


def run_reference(workload: str, seconds: float = 5.0, procs: int | None = None, dumps: bool = False):
    """Like ``run`` but every process steps the UNMODIFIED reference (kind "reference"); None where the
    reference is not installed.  ``dumps=True`` (mobile-custom-v0 only) keeps the four JSON dumps per
    step the fork ships with (base.py:261, 298-349), written into a scratch directory."""
    if not reference_installed():
        return None
    procs = procs or os.cpu_count() or 1
    ctx = mp.get_context("spawn")
    with ctx.Pool(procs) as pool:
        res = pool.map(_ref_worker, [(workload, seconds, 1000 + i, dumps) for i in range(procs)])
    out = _summary(workload, procs, seconds, res, kind="reference")
    if dumps:
        out["sample"] = out["sample"].replace("JSON dumps off", "as shipped: 4 JSON dumps per step into a scratch directory")
    return out


class Runner:
    """Persistent worker pool for the ``--impl reference`` arm: one call of ``step`` = every host
    core steps its own env for ``seconds`` (a bounded sample of the workload)."""

    def __init__(self, workload: str, procs: int | None = None, kind: str = "port"):
        self.workload = workload
        self.procs = procs or os.cpu_count() or 1
        self.pool = mp.get_context("spawn").Pool(self.procs)
        self.calls = 0
        self.kind = kind  # "reference": the unmodified reference from oracle/_ref; "port": the scalar restatement

    def step(self, seconds: float):
        self.calls += 1
        if self.kind == "reference":
            args = [(self.workload, seconds, 1000 * self.calls + i, False) for i in range(self.procs)]
            return _summary(self.workload, self.procs, seconds, self.pool.map(_ref_worker, args), kind="reference")
        args = [(self.workload, seconds, 1000 * self.calls + i) for i in range(self.procs)]
        return _summary(self.workload, self.procs, seconds, self.pool.map(_worker, args))

    def close(self):
        self.pool.close()
        self.pool.join()


def _compiled_worker(args):
    """The same workload through the COMPILED restatement (oracle/mbe_oracle_c.c): E envs per call,
    one env per OpenMP iteration on all host threads."""
    workload, seconds, envs = args
    from oracle.c_oracle import CEnvBatch

    bs, U, mode, handler, vel, over = WORKLOADS[workload]
    p = orc.Params(velocity=vel, **over)
    rng = np.random.default_rng(7)
    envs = min(envs, max(64, (1 << 22) // (U * len(bs or [0] * 10))))  # bound the wide shapes' memory and time
    E = envs
    if bs is None:  # the fork's scenario: 5..10 random BSs per env (custom.py:68-77)
        layout = rng.integers(0, 200, size=(E, 10, 2)).astype(np.int32)
        env = CEnvBatch(p, layout, E, U, nbs=rng.integers(5, 11, size=E))
    else:
        env = CEnvBatch(p, bs, E, U, handler=handler)
    B = env.B
    pool_wp = [rng.integers(0, int(p.width), size=(E, U, 2)).astype(np.int32) for _ in range(4)]
    pool_act = [rng.integers(0, B + 1, size=(E, U)).astype(np.int32) for _ in range(4)]

    def fresh():
        env.reset(rng.integers(0, int(p.width), size=(E, U, 2)))

    fresh()
    calls = 0
    t0 = time.perf_counter()
    while True:
        if mode == "gym":
            env.step_gym(pool_act[calls % 4], pool_wp[calls % 4])
        else:
            env.step_fork(pool_wp[calls % 4])
        calls += 1
        if env.done[0]:
            fresh()
        if time.perf_counter() - t0 >= seconds:
            break
    return calls * E, time.perf_counter() - t0, E


def run_compiled(workload: str, seconds: float = 2.0, envs: int = 8192):
    """env-steps/s of the compiled C restatement on all host threads (OpenMP), or None when the
    workload is outside what it covers.  Runs in a child process with a time-out (keeps libgomp out of
    the caller; a missing compiler or a crash can never break or hang the bench line)."""
    import json
    import subprocess
    import sys

    if workload not in WORKLOADS:
        return None
    threads = os.cpu_count() or 1
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    try:
        res = subprocess.run([sys.executable, "-m", "oracle.cpu_baseline", "compiled", workload, str(seconds), str(envs)],
                             cwd=root, capture_output=True, text=True, timeout=seconds + 120)
        if res.returncode != 0:
            raise RuntimeError(res.stderr.strip().splitlines()[-1] if res.stderr.strip() else f"exit {res.returncode}")
        steps, wall, envs = json.loads(res.stdout.strip().splitlines()[-1])
    except Exception as exc:  # noqa: BLE001
        return {"unavailable": f"{type(exc).__name__}: {exc}"[:200]}
    return {
        "value": steps / wall, "unit": "env-steps/s", "cores": threads, "kind": "port (compiled C, OpenMP)",
        "sample": f"{envs} envs x {steps // envs} steps of {workload} in {wall:.2f} s through oracle/mbe_oracle_c.c "
                  f"(FP64, the reference's arithmetic per UE x BS pair, one env per OpenMP iteration)",
    }


if __name__ == "__main__":
    import json
    import sys

    if len(sys.argv) == 5 and sys.argv[1] == "compiled":
        from oracle import c_oracle

        c_oracle.build()
        print(json.dumps(_compiled_worker((sys.argv[2], float(sys.argv[3]), int(sys.argv[4])))))
    else:
        print(json.dumps(run(sys.argv[1] if len(sys.argv) > 1 else "mobile-medium-central-v0", 2.0)))
