#!/usr/bin/env python
"""Closed-loop stepping with a policy network on the same GPU: the step kernel writes the observation
straight into the policy's input tensor and reads the policy's int32 actions from device memory, so a
whole rollout runs without a host round trip (and replays from one CUDA graph).

    python examples/policy_in_the_loop.py [--envs 65536] [--steps 200] [--workload mobile-medium-central-v0]

The policy is a small per-UE MLP with shared weights (random initialisation: this measures the loop, it
does not learn).  Prints env-steps/s for the loop and for the env alone.
"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mobile_env_gan_b200 as mbe  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--envs", type=int, default=65536)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--hidden", type=int, default=64)
    ap.add_argument("--workload", default="mobile-medium-central-v0")
    args = ap.parse_args()

    env = mbe.make(args.workload, num_envs=args.envs, autoreset=True)
    E, U, B, F = env.num_envs, env.NUM_USERS, env.NUM_STATIONS, env.plan.feature_size
    torch.manual_seed(0)
    policy = torch.nn.Sequential(torch.nn.Linear(F, args.hidden), torch.nn.ReLU(),
                                 torch.nn.Linear(args.hidden, B + 1)).to(env.device, torch.bfloat16)
    obs, _ = env.reset()
    rows = env.obs.view(E * U, F)           # the tensor the step kernel writes: no copy
    returns = torch.zeros(E, device=env.device)

    @torch.no_grad()
    def loop_step():
        logits = policy(rows.to(torch.bfloat16))
        env.actions.copy_(logits.argmax(dim=1).view(E, U))   # int64 -> the env's int32 action tensor
        _, reward, _, _, _ = env.step(env.actions)
        returns.add_(reward if reward.dim() == 1 else reward.mean(dim=1))

    stream = torch.cuda.Stream()
    with torch.cuda.stream(stream):
        for _ in range(3):
            loop_step()
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph, stream=stream):
            for _ in range(20):
                loop_step()
        graph.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = max(1, args.steps // 20)
        e0.record(stream)
        for _ in range(reps):
            graph.replay()
        e1.record(stream)
        torch.cuda.synchronize()
        ms_loop = e0.elapsed_time(e1) / (reps * 20)
        e0.record(stream)
        for _ in range(reps * 20):
            env.step(env.actions)
        e1.record(stream)
        torch.cuda.synchronize()
        ms_env = e0.elapsed_time(e1) / (reps * 20)
    print(f"{args.workload}, {E} envs: policy + step {ms_loop * 1e3:.1f} us per step = {E / ms_loop * 1e3:.3e} env-steps/s "
          f"(env alone {ms_env * 1e3:.1f} us = {E / ms_env * 1e3:.3e}); mean return per step {float(returns.mean()) / (3 + 20 * (reps + 1)):.4f}")


if __name__ == "__main__":
    main()
