#!/usr/bin/env python
"""Benchmark of the hot path: env-steps/s of batched MComCore.step (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" is one fused launch of the step kernel over one batch of ``--envs`` (default 65,536)
mobile-medium-central environments per GPU (BASELINE.json configs[1]); envs shard over ranks
with no collective (weak scaling: every GPU owns its own 65,536-env batches).  To keep the
timed inputs out of L2 the bench rotates over R independent env batches whose combined
footprint exceeds the 126 MB L2 (``config.l2``).  Steps are replayed from a CUDA graph.

Printed JSON line (rank 0): value = whole-job env-steps/s with inputs resident in HBM;
e2e = the same metric through the host-buffer C-ABI call (mbe_step_host: pinned host actions
in, obs/reward/done out, copies inside the timed region); roofline = algorithmic bytes of the
step kernel / its average duration against the measured HBM peak; cpu_baseline = the UNMODIFIED
reference's MComCore.step (oracle/_ref, kind "reference"; the oracle's scalar-Python port where the
reference is not installed) timed on this box's host cores, with ``cpu_baseline.port`` = the scalar
port and ``cpu_baseline.compiled`` = the same arithmetic as a compiled C + OpenMP restatement.
Secondary keys (never used for value / roofline): e2e.obs_stays_on_device, two_env_groups_in_flight,
fused_episode (FORK workloads, mbe_rollout).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "env-steps/sec (batched, whole box) for mobile-medium at 1/2/4/8 B200"
UNIT = "env-steps/s"
SIZES = {"mobile-small": (3, 5), "mobile-medium": (4, 15), "mobile-large": (13, 30), "mobile-synthetic": (64, 512),
         "mobile-custom": (10, 7)}  # custom = the fork's MComCustom (FORK mode, 5..10 random BSs per env)


def bytes_per_env_step_fork(U: int, B: int) -> dict:
    """FORK step (MComCustom): pos r+w 8, waypoint r 4, assoc w 4, rate f64 w 8, utility w 4 per UE;
    per env: BS table 4B r, nbs 4, clock r+w 8, episode 4, done 1, metrics 16."""
    ours = U * (8 + 4 + 4 + 8 + 4) + 4 * B + 4 + 8 + 4 + 1 + 16
    return {"layout": ours, "survey_8d": U * (32 + 4) + 8 * B + 13}


def bytes_per_env_step(U: int, B: int, handler: str) -> dict:
    """Algorithmic HBM bytes of one env-step for THIS build's layout and the
    canonical int32/fp32 accounting of SURVEY.md section 8(d)."""
    F = (2 * B + 1) if handler == "central" else (4 * B + 1)
    mw = (B + 31) // 32
    # pos r+w (int16x2) 8, waypoint read 4, actions 4, conn r+w 8*mw, rate f64 w 8, utility f32 w 4,
    # obs f32 w 4F per UE; per env: reward, done, clock r+w, episode r, metrics
    per_ue = 8 + 4 + 4 + 8 * mw + 8 + 4 + 4 * F
    per_env = (4 if handler == "central" else 4 * U) + 1 + 8 + 4 + 16
    ours = U * per_ue + per_env
    canon = U * (32 + 8 * mw + 4 * F) + 13 if handler == "central" else U * (36 + 8 * mw + 4 * F) + 9
    return {"layout": ours, "survey_8d": canon}


class ClockSampler(threading.Thread):
    """Polls NVML for SM clock and throttle reasons while the timed region runs."""

    BAD = {"hw_slowdown": 0x8, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20}
    NOTE = {"sw_power_cap": 0x4}

    def __init__(self, index: int, enabled: bool = True):
        super().__init__(daemon=True)
        self.samples, self.reasons, self.stop_flag, self.ok = [], set(), False, False
        self.err = "sampled on rank 0 only"
        if not enabled:
            return
        try:
            import pynvml

            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception as e:  # pragma: no cover
            self.err = str(e)

    def sample(self):
        nv = self.nv
        self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
        try:
            r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
        except Exception:
            r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
        for name, bit in {**self.BAD, **self.NOTE}.items():
            if r & bit:
                self.reasons.add(name)

    def run(self):
        while self.ok and not self.stop_flag:
            self.sample()
            time.sleep(0.02)  # 50 Hz: NVML takes a driver lock that graph launches also need

    def result(self):
        if not self.ok or not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvml unavailable"]}
        return {"sm_mhz": statistics.median(self.samples), "sm_max_mhz": self.max_mhz,
                "samples": len(self.samples), "reasons": sorted(self.reasons)}


def measured_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic(workload):
    """dram bytes per launch of the step kernel from the committed ncu capture, if any."""
    path = os.path.join(ROOT, "profiles", "step_kernel_traffic.json")
    if os.path.exists(path):
        return json.load(open(path)).get(workload)
    return None


def workload_shape(workload):
    fork = workload == "mobile-custom-v0"
    if fork:
        size, handler = "mobile-custom", "fork"
    else:
        size, handler = workload.rsplit("-", 2)[0], workload.split("-")[2]
    B, U = SIZES[size]
    return fork, handler, B, U


def rotation(workload, E):
    """(bytes-per-env-step dict, R, footprint of one batch): the bench rotates over R independent env
    batches so that a step's inputs and outputs are not L2-resident."""
    fork, handler, B, U = workload_shape(workload)
    bpe = bytes_per_env_step_fork(U, B) if fork else bytes_per_env_step(U, B, handler)
    footprint = E * bpe["layout"]
    R = max(2, -(-int(2.2 * 126e6) // footprint))
    return bpe, R, footprint


def make_config(args):
    """`config` of the JSON line: identical for both arms (the reference arm times the same workload
    on the host cores)."""
    fork, handler, B, U = workload_shape(args.workload)
    _, R, footprint = rotation(args.workload, args.envs)
    return {
        "workload": args.workload, "envs_per_gpu": args.envs, "bs": B, "ues": U,
        "actions": "none (FORK step)" if fork else "uniform int32 in [0,B], resident in HBM",
        "autoreset": True, "ep_time": 20,
        "l2": f"rotating {R} independent env batches, {R * footprint / 1e6:.0f} MB > 126 MB L2",
        "launch": "stream launches" if args.no_graph else "CUDA graph replay",
    }


def run_reference(args, rank, world):
    """--impl reference: the reference's own CPU implementation of the path on all host cores; each step
    is a bounded sample.  Where the reference is installed (oracle/_ref, built by build() from
    /root/reference with pip --no-deps; it travels to the GPU box) every process steps the UNMODIFIED
    ``MComCore.step`` (kind "reference": the fork's FORK-order step -- it has no actions or observations --
    on the workload's layout, JSON dumps off); otherwise the oracle's scalar port
    of the same step (kind "port")."""
    if rank != 0:
        return
    from oracle import cpu_baseline

    workload = args.workload
    procs = os.cpu_count() or 1
    total_steps = args.warmup + args.steps
    # every step is a bounded sample; the whole run is held to about two minutes whatever K is
    seconds = max(0.02, min(5.0, 120.0 / max(total_steps, 1)))
    kind = "reference" if cpu_baseline.reference_installed() else "port"
    runner = cpu_baseline.Runner(workload, procs, kind=kind)
    vals = []
    for i in range(total_steps):
        r = runner.step(seconds)
        if i >= args.warmup:
            vals.append(r)
    runner.close()
    value = statistics.mean(v["value"] for v in vals)
    # for transparency only (`value` stays the reference's own speed class, scalar Python): the same
    # arithmetic as a compiled C + OpenMP restatement on the same cores
    compiled = cpu_baseline.run_compiled(workload, 2.0)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": seconds * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": make_config(args),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": procs, "kind": kind,
                         "sample": vals[-1]["sample"] + "; each step = every host core stepping its own env for a "
                                   "bounded time", "single_core": vals[-1]["single_core"], "compiled": compiled},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2000)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="mobile-medium-central-v0")
    ap.add_argument("--envs", type=int, default=65536, help="envs per GPU per batch")
    ap.add_argument("--total-envs", type=int, default=0,
                    help="strong scaling: this many envs per batch over ALL GPUs (envs per GPU = total / N)")
    ap.add_argument("--cpu-seconds", type=float, default=4.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--preheat-seconds", type=float, default=1.0)
    ap.add_argument("--repeats", type=int, default=0,
                    help="times the K-step block is timed (median reported); 0 = about 2000 / K, at most 200")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.total_envs:
        if args.total_envs % (32 * world):
            ap.error("--total-envs must be a multiple of 32 x the number of GPUs")
        args.envs = args.total_envs // world
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    # CPU baseline first (rank 0, N=1 only), before CUDA is initialised in this process
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        from oracle import cpu_baseline

        procs = os.cpu_count() or 1
        port = cpu_baseline.run(args.workload, args.cpu_seconds, procs)
        # the unmodified reference where it is installed (oracle/_ref), else the oracle's scalar port
        cpu = cpu_baseline.run_reference(args.workload, args.cpu_seconds, procs) or port
        if cpu is not port:
            cpu["port"] = port  # the scalar restatement (GYM order with observations for the GYM workloads)
            if args.workload == "mobile-custom-v0":  # the fork as shipped: four JSON dumps per step
                cpu["as_shipped"] = cpu_baseline.run_reference(args.workload, min(2.0, args.cpu_seconds), procs, dumps=True)
        # beside them: the same arithmetic as a compiled C + OpenMP restatement
        cpu["compiled"] = cpu_baseline.run_compiled(args.workload, min(2.0, args.cpu_seconds))

    # pin this rank to the CPUs / NUMA node next to its GPU before any pinned host buffer exists,
    # so that the end-to-end leg's DMA targets local memory (matters at N > 1)
    affinity = "unchanged"
    if os.environ.get("MBE_BENCH_AFFINITY", "1") != "0":
        try:
            import pynvml

            pynvml.nvmlInit()
            pynvml.nvmlDeviceSetCpuAffinity(pynvml.nvmlDeviceGetHandleByIndex(local_rank))
            affinity = f"nvml ideal cpus ({len(os.sched_getaffinity(0))})"
        except Exception as exc:  # keep going: it is an optimisation only
            affinity = f"unchanged ({type(exc).__name__})"

    import torch
    import torch.distributed as dist

    import mobile_env_gan_b200 as mbe

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    fork, handler, B, U = workload_shape(args.workload)
    E = args.envs
    # rotate over R independent batches so that each step's inputs/outputs are not L2-resident
    bpe, R, footprint = rotation(args.workload, E)
    envs = []
    for r in range(R):
        env = mbe.make(args.workload, num_envs=E, device=str(dev), autoreset=True,
                       env_offset=(rank * R + r) * E)
        env.reset()
        envs.append(env)
    g = torch.Generator(device=dev).manual_seed(1234 + rank)
    for env in envs:  # synthetic policy output: uniform actions in [0, B]
        if not fork:
            env.actions.copy_(torch.randint(0, B + 1, (E, U), generator=g, device=dev, dtype=torch.int32))
    stream = torch.cuda.Stream(device=dev)

    def raw_steps(n, start=0):
        for i in range(n):
            env = envs[(start + i) % R]
            if fork:
                env.step(0, i)
            else:
                env.step(env.actions)

    chunk = R * 32
    graph = graph_rem = None
    rem = args.steps % chunk
    with torch.cuda.stream(stream):
        raw_steps(chunk)  # first touches outside the graph
        torch.cuda.synchronize()
        if not args.no_graph:
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph, stream=stream):
                raw_steps(chunk)
            if rem:  # the timed region is then graph replays only, whatever K is
                graph_rem = torch.cuda.CUDAGraph()
                with torch.cuda.graph(graph_rem, stream=stream):
                    raw_steps(rem)

        def run_steps(n):
            done = 0
            while graph is not None and n - done >= chunk:
                graph.replay()
                done += chunk
            if graph_rem is not None and n - done == rem:
                graph_rem.replay()
                done += rem
            raw_steps(n - done)

        # preheat (untimed) so that clocks are up, then W warm-up steps
        t_end = time.perf_counter() + args.preheat_seconds
        while time.perf_counter() < t_end:
            run_steps(chunk)
            torch.cuda.synchronize()
        # warm-up: at least W steps, issued as exactly the K-step sequence the timed region will issue,
        # so that every CUDA graph replayed under the clock (the full-chunk graph and the remainder
        # graph) has been launched -- and so uploaded to the device -- before
        for _ in range(max(2, -(-max(args.warmup, 3) // args.steps))):
            run_steps(args.steps)
        torch.cuda.synchronize()
        # A short K would make the timed region a few hundred microseconds: the K-step block is then
        # timed `repeats` times (each bracketed by barrier + synchronize, max over ranks per block) and
        # the median block is reported.  K x repeats ~ 2,000 steps whatever K is.
        repeats = max(1, min(200, 2000 // max(args.steps, 1))) if args.repeats <= 0 else args.repeats
        sampler = ClockSampler(local_rank, enabled=(rank == 0))
        if sampler.ok:
            sampler.sample()
        sampler.start()
        gate_cycles = int(os.environ.get("MBE_BENCH_GATE_CYCLES", "400000"))  # ~0.2 ms spin kernel
        pairs = []
        for _ in range(repeats):
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
            ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            # gate: a spin kernel runs while the host enqueues the event and the graph launches, so the
            # launch latency of the first graph is not inside the timed region
            if gate_cycles > 0:
                torch.cuda._sleep(gate_cycles)
            ev0.record(stream)
            run_steps(args.steps)
            ev1.record(stream)
            torch.cuda.synchronize()
            pairs.append((ev0, ev1))
        sampler.stop_flag = True
        sampler.join()
        if sampler.ok:
            sampler.sample()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        block_ms = torch.tensor([a_.elapsed_time(b_) for a_, b_ in pairs], dtype=torch.float64, device=dev)
        per_rank_median = None
        if world > 1:
            # transparency: every rank's own median block (the reported time is the max over ranks per block)
            mine = block_ms.median().reshape(1)
            gathered = [torch.zeros_like(mine) for _ in range(world)]
            dist.all_gather(gathered, mine)
            per_rank_median = [float(t_) for t_ in gathered]
            dist.all_reduce(block_ms, op=dist.ReduceOp.MAX)  # every block: max over ranks
        block_ms = sorted(float(v) for v in block_ms)
        ms_block_max = statistics.median(block_ms)  # median over blocks of (max over ranks of that block)
        # Reported time: every rank's median block, then the MAX over ranks (the slowest rank's typical
        # block).  The step path has no collective, so ranks never wait for each other inside a block;
        # taking the max inside every 0.3 ms block instead picks up whichever rank's launch jittered in
        # that block (8 ranks: +4 %), not a slower rank -- that figure stays in `timing` beside it.
        ms = max(per_rank_median) if per_rank_median else ms_block_max
        gpu_launches = args.steps  # one step_kernel launch per step (graph nodes included)

        # ---- secondary: two env groups in flight (two streams, each its own dependent chain) ----
        # A single chain of 18 us kernels pays ~3 us of ramp/drain per launch; when the caller
        # steps two groups of envs alternately (double-buffered actors) the groups' launches
        # overlap and hide it.  Reported separately; `value` above stays the single-stream number.
        two = None
        if not args.no_graph and R >= 2 and not fork:
            streams2 = [torch.cuda.Stream(device=dev) for _ in range(2)]
            graphs2, sub = [], 64
            for si, st2 in enumerate(streams2):
                mine = [envs[i] for i in range(R) if i % 2 == si]
                with torch.cuda.stream(st2):
                    for i in range(len(mine)):
                        mine[i].step(mine[i].actions)
                    torch.cuda.synchronize()
                    g2 = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(g2, stream=st2):
                        for i in range(sub):
                            mine[i % len(mine)].step(mine[i % len(mine)].actions)
                graphs2.append(g2)
            reps = max(1, args.steps // (2 * sub))
            for g2, st2 in zip(graphs2, streams2):
                with torch.cuda.stream(st2):
                    g2.replay()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            for st2 in streams2:
                st2.wait_event(e0)
            for _ in range(reps):
                for g2, st2 in zip(graphs2, streams2):
                    with torch.cuda.stream(st2):
                        g2.replay()
            for st2 in streams2:
                stream.wait_stream(st2)
            e1.record(stream)
            torch.cuda.synchronize()
            two = (e0.elapsed_time(e1), reps * 2 * sub)

        # secondary figure for the fork's own use (collect / layout search: 20 policy-free steps per
        # epoch, custom.py + chooseBaseStation.ipynb): whole episodes through mbe_rollout, one launch
        # per episode with the layout-score statistics accumulated in the kernel
        fused = None
        if fork:
            from mobile_env_gan_b200.scoring import LayoutScorer

            scorers = [LayoutScorer(e) for e in envs]
            T = envs[0].plan.ep_time
            reps = max(R, args.steps // T)
            for sc in scorers:
                sc.run_episode(T)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            for r in range(reps):
                scorers[r % R].run_episode(T)
            e1.record(stream)
            torch.cuda.synchronize()
            fused = (e0.elapsed_time(e1), reps * T, reps)

        # ---- end to end through the host-buffer ABI call ----
        e2e_steps = max(5, min(args.steps, 40))
        if fork:
            # the reference's step returns None; its caller reads positions, association, rates and
            # QoE from the env afterwards (base.py:264-269): step + those tensors to pinned host memory
            names = ("pos", "assoc", "rate", "utility_scaled", "metrics", "done")
            host = {n: torch.empty_like(getattr(envs[0], n), device="cpu").pin_memory() for n in names}
            h2d, d2h = 0, sum(t.numel() * t.element_size() for t in host.values())

            def e2e_step(i):
                env = envs[i % R]
                env.step(0, i)
                for n in names:
                    host[n].copy_(getattr(env, n), non_blocking=True)
                torch.cuda.current_stream().synchronize()

            api = "MComCustom.step + D2H of pos/assoc/rate/utility/metrics/done"
        else:
            F = envs[0].plan.feature_size
            acts_h = [torch.randint(0, B + 1, (E, U), dtype=torch.int32).pin_memory() for _ in range(2)]
            obs_h = torch.empty(E, U * F, dtype=torch.float32).pin_memory()
            rew_h = torch.empty(E, dtype=torch.float32).pin_memory() if handler == "central" else \
                torch.empty(E, U, dtype=torch.float32).pin_memory()
            done_h = torch.empty(E, dtype=torch.uint8).pin_memory()
            h2d = E * U * 4
            d2h = E * U * F * 4 + rew_h.numel() * 4 + E

            def e2e_step(i):
                envs[i % R].step_host(acts_h[i % 2], obs_h, rew_h, done_h)

            api = "mbe_step_host (pinned host buffers)"
        for i in range(3):
            e2e_step(i)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for i in range(e2e_steps):
            e2e_step(i)
        torch.cuda.synchronize()
        e2e_s = time.perf_counter() - t0
        checksum = float(host["utility_scaled"].sum()) if fork else float(rew_h.sum())
        # secondary figure: the deployment the step is built for -- the observation stays in the policy's
        # input tensor on the device, only reward and done come back to the host
        e2e_lite = None
        if not fork:
            for i in range(3):
                envs[i % R].step_host(acts_h[i % 2], None, rew_h, done_h)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for i in range(e2e_steps):
                envs[i % R].step_host(acts_h[i % 2], None, rew_h, done_h)
            torch.cuda.synchronize()
            e2e_lite = (time.perf_counter() - t0, rew_h.numel() * 4 + E)

    times = torch.tensor([0.0, e2e_s * 1e3, (e2e_lite[0] if e2e_lite else 0.0) * 1e3, two[0] if two else 0.0,
                          fused[0] if fused else 0.0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
    _, e2e_ms, lite_ms, two_ms, fused_ms = (float(t) for t in times)
    if rank == 0:
        peak, peak_src = measured_peak()
        per_launch_s = ms * 1e-3 / args.steps
        # roofline.achieved = env-steps/s x the canonical algorithmic bytes per env-step
        # (int32 / fp32 structure of arrays); this build's int16 position layout moves fewer bytes for the
        # same work -- that figure is reported beside it (frac_layout) and is what ncu's DRAM counters see
        achieved = bpe["survey_8d"] * E / per_launch_s / 1e9
        value = world * E * args.steps / (ms * 1e-3)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "strong" if args.total_envs else "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": make_config(args),
            "repeats": repeats,
            "timing": {"what": f"{repeats} timed blocks of K={args.steps} steps, each bracketed by barrier + synchronize; "
                               "CUDA events on the launching stream behind a spin-kernel gate; reported = each "
                               "rank's median block, MAX over ranks",
                       "block_ms_min": block_ms[0], "block_ms_median": ms, "block_ms_max": block_ms[-1],
                       "per_rank_block_ms_median": per_rank_median,
                       "median_of_per_block_max_over_ranks_ms": ms_block_max},
            "roofline": {
                "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": ncu_traffic(args.workload),
                "traffic_note": "ncu dram__bytes_read.sum + dram__bytes_write.sum per launch in STEADY STATE: mean of "
                                "launches 13-20 of this rotating-batch loop, application replay, caches not flushed "
                                "(profiles/step_kernel_traffic.json, profiles/r02_f_traffic_*.csv)",
                "peak_source": peak_src,
                "kernel": f"mbe::{envs[0].step_kernel_name}<{'fork' if fork else 'gym'},{handler},U={U},B={B}>",
                "bytes_per_env_step": bpe["survey_8d"], "bytes_per_env_step_survey_8d": bpe["survey_8d"],
                "frac_survey_8d": achieved / peak,
                "bytes_per_env_step_layout": bpe["layout"],
                "achieved_layout": bpe["layout"] * E / per_launch_s / 1e9,
                "frac_layout": bpe["layout"] * E / per_launch_s / 1e9 / peak,
                "accounting": "achieved / frac: SURVEY.md 8(d) canonical algorithmic bytes per env-step (int32/fp32 SoA); "
                              "*_layout: the bytes this build's int16-position layout really moves per env-step "
                              "(what `traffic` from ncu is compared with)",
            },
            "cpu_baseline": cpu,
            "e2e": {"value": world * E * e2e_steps / (e2e_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h, "steps": e2e_steps, "api": api, "cpu_affinity": affinity,
                    "checksum": checksum,
                    "obs_stays_on_device": None if not e2e_lite else {
                        "value": world * E * e2e_steps / (lite_ms * 1e-3), "unit": UNIT,
                        "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": e2e_lite[1],
                        "note": "same call with obs_host=NULL: actions in, reward+done out"}},
            "two_env_groups_in_flight": None if not two else {
                "value": world * E * two[1] / (two_ms * 1e-3), "unit": UNIT, "steps": two[1],
                "ms_per_step": two_ms / two[1], "streams": 2,
                "hbm_frac": bpe["survey_8d"] * E / (two_ms * 1e-3 / two[1]) / 1e9 / peak,
                "hbm_frac_layout": bpe["layout"] * E / (two_ms * 1e-3 / two[1]) / 1e9 / peak,
                "note": "throughput when two groups of env batches are stepped concurrently (each group a "
                        "dependent chain on its own stream); not used for `value` or `roofline`"},
            "fused_episode": None if not fused else {
                "value": world * E * fused[1] / (fused_ms * 1e-3), "unit": UNIT, "steps_per_launch": fused[1] // fused[2],
                "launches": fused[2], "ms_per_launch": fused_ms / fused[2],
                "note": "mbe_rollout: whole episodes in one launch each with the layout-score statistics "
                        "accumulated in the kernel; not used for `value` or `roofline`"},
            "gpu_launches": gpu_launches,
            "clocks": sampler.result(),
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
