#!/usr/bin/env python
"""Experiment: does a short-lived tail shorten the drain of the 65,536-env medium launch?

One step = TWO concurrent kernels on two graph branches: the UEs-per-thread kernel (long-lived CTAs,
fewest instructions) on envs [0, E1) and the one-thread-per-UE warp-segment kernel (short-lived CTAs)
on envs [E1, E).  Two handles are bound to the SAME tensors (the second created with MBE_UPT=0) and
stepped through mbe_step_window.  Prints us per step for several split points; E1 = E is the baseline.
    python profiles/hybrid_tail_experiment.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mobile_env_gan_b200 as mbe  # noqa: E402

E, R = 65536, 4
NAMES = ("pos", "wp", "t", "episode", "bs_xy", "conn", "actions", "rate", "utility_scaled", "obs", "reward", "done",
         "metrics", "_terminated")


def make_pair(r):
    a = mbe.make("mobile-medium-central-v0", num_envs=E, autoreset=True, env_offset=r * E)
    os.environ["MBE_UPT"] = "0"
    b = mbe.make("mobile-medium-central-v0", num_envs=E, autoreset=True, env_offset=r * E)
    del os.environ["MBE_UPT"]
    a.reset()
    for n in NAMES:
        setattr(b, n, getattr(a, n))
    b._bind()
    b._needs_reset = False
    assert a.step_kernel_name == "step_upt_kernel" and b.step_kernel_name == "step_spec_kernel"
    return a, b


def main():
    pairs = [make_pair(r) for r in range(R)]
    g = torch.Generator(device="cuda").manual_seed(0)
    for a, _ in pairs:
        a.actions.copy_(torch.randint(0, 5, (E, 15), generator=g, device="cuda", dtype=torch.int32))
    main_s, side = torch.cuda.Stream(), torch.cuda.Stream()

    def step(a, b, e1):
        if e1 >= E:
            a.step_window(0, E, main_s)
            return
        ev = torch.cuda.Event()
        ev.record(main_s)
        side.wait_event(ev)
        a.step_window(0, e1, main_s)
        b.step_window(e1, E - e1, side)
        ev2 = torch.cuda.Event()
        ev2.record(side)
        main_s.wait_event(ev2)

    for e1 in (E, 61440, 57600, 53760, 49920, 46080, 39936, E):
        with torch.cuda.stream(main_s):
            for i in range(2 * R):
                step(*pairs[i % R], e1)
            torch.cuda.synchronize()
            gr = torch.cuda.CUDAGraph()
            with torch.cuda.graph(gr, stream=main_s):
                for i in range(128):
                    step(*pairs[i % R], e1)
            for _ in range(4):
                gr.replay()
            torch.cuda.synchronize()
            e0, e1v = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(main_s)
            for _ in range(16):
                gr.replay()
            e1v.record(main_s)
            torch.cuda.synchronize()
            us = e0.elapsed_time(e1v) * 1e3 / (16 * 128)
        print(f"upt envs [0,{e1}) + spec envs [{e1},{E}): {us:6.2f} us per step", flush=True)


if __name__ == "__main__":
    main()
