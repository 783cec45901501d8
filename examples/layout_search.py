#!/usr/bin/env python
"""The fork's layout search (chooseBaseStation.ipynb: run an epoch per random BS layout, score it with
`qoeValue`, keep the best) as a batched search on one B200: every env is one candidate layout, a
whole 20-step epoch is one kernel launch (mbe_rollout) and the score statistics never leave the GPU.

    python examples/layout_search.py --layouts 1000000 --rounds 5 --top 10
    torchrun --nproc-per-node 8 examples/layout_search.py --layouts 8388608   # layouts sharded over the GPUs

Multi-GPU: every rank scores its own contiguous slice of the global layout range (Philox layouts are
keyed by the global env id, so the candidates do not depend on the number of GPUs); the only exchange
is an all-gather of each rank's top-k (score, layout) rows per round.
"""
import argparse
import os
import sys
import time

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mobile_env_gan_b200.scenarios.custom import MComCustom  # noqa: E402
from mobile_env_gan_b200.scoring import LayoutScorer  # noqa: E402
from mobile_env_gan_b200.sharding import sharded_config  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--layouts", type=int, default=1 << 20, help="candidate layouts per round (envs)")
    ap.add_argument("--rounds", type=int, default=3, help="each round draws fresh layouts (next episode)")
    ap.add_argument("--top", type=int, default=5)
    args = ap.parse_args()
    world, rank, local = (int(os.environ.get(k, d)) for k, d in (("WORLD_SIZE", 1), ("RANK", 0), ("LOCAL_RANK", 0)))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    total = (args.layouts + 32 * world - 1) // (32 * world) * (32 * world)  # whole warps of envs on every rank
    # autoreset: the step that ends an epoch draws the next layout, like MComCustom.reset (custom.py:40-77)
    env = MComCustom(config={"autoreset": True, "device": f"cuda:{local}", **sharded_config(total, rank, world)})
    E, B = env.num_envs, env.plan.num_bs
    scorer = LayoutScorer(env)
    env.reset()
    best = []
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for r in range(args.rounds):
        layouts, nbs = env.bs_xy.clone(), env.nbs.clone()  # the layouts this epoch runs on
        scorer.reset()
        scorer.run_episode()  # one launch: 20 steps x E envs + score statistics
        idx, score = scorer.best(min(args.top, E))
        rows = torch.cat([score[:, None], nbs[idx, None].float(), layouts[idx].reshape(len(idx), -1).float()], dim=1)
        if world > 1:  # the only exchange: every rank's top-k rows
            out = rows.new_empty((world * rows.shape[0], rows.shape[1]))
            dist.all_gather_into_tensor(out, rows.contiguous())
            rows = out
        cand = [(float(row[0]), row[2:2 + 2 * int(row[1])].reshape(-1, 2).int().tolist()) for row in rows.cpu()]
        best = sorted(best + cand, key=lambda c: -c[0])[: args.top]
        if rank == 0:
            top = max(cand, key=lambda c: c[0])
            print(f"round {r}: best score {top[0]:+.4f} with {len(top[1])} base stations")
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    if rank == 0:
        n = args.rounds * total
        print(f"{n} layouts x {env.plan.ep_time} steps on {world} GPU(s) in {dt:.3f} s = "
              f"{n * env.plan.ep_time / dt:.3e} env-steps/s (wall clock, incl. top-k and gather)")
        for s, xy in best:
            print(f"{s:+.4f}  {xy}")
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
