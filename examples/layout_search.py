#!/usr/bin/env python
"""The fork's layout search (chooseBaseStation.ipynb: run an epoch per random BS layout, score it with
`qoeValue`, keep the best) as a batched search on one B200: every env is one candidate layout, a
whole 20-step epoch is one kernel launch (mbe_rollout) and the score statistics never leave the GPU.

    python examples/layout_search.py --layouts 1000000 --rounds 5 --top 10
"""
import argparse
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mobile_env_gan_b200.scenarios.custom import MComCustom  # noqa: E402
from mobile_env_gan_b200.scoring import LayoutScorer  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--layouts", type=int, default=1 << 20, help="candidate layouts per round (envs)")
    ap.add_argument("--rounds", type=int, default=3, help="each round draws fresh layouts (next episode)")
    ap.add_argument("--top", type=int, default=5)
    args = ap.parse_args()
    E = (args.layouts + 31) // 32 * 32
    # autoreset: the step that ends an epoch draws the next layout, like MComCustom.reset (custom.py:40-77)
    env = MComCustom(config={"num_envs": E, "autoreset": True})
    scorer = LayoutScorer(env)
    env.reset()
    best = None
    t0 = time.perf_counter()
    for r in range(args.rounds):
        layouts, nbs = env.bs_xy.clone(), env.nbs.clone()  # the layouts this epoch runs on
        scorer.reset()
        scorer.run_episode()  # one launch: 20 steps x E envs + score statistics
        idx, score = scorer.best(args.top)
        cand = [(float(s), layouts[i, : int(nbs[i])].cpu().tolist()) for i, s in zip(idx.tolist(), score.tolist())]
        best = sorted((best or []) + cand, key=lambda c: -c[0])[: args.top]
        print(f"round {r}: best score {cand[0][0]:+.4f} with {len(cand[0][1])} base stations")
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    n = args.rounds * E
    print(f"{n} layouts x {env.plan.ep_time} steps in {dt:.3f} s = {n * env.plan.ep_time / dt:.3e} env-steps/s (wall clock, incl. top-k)")
    for s, xy in best:
        print(f"{s:+.4f}  {xy}")


if __name__ == "__main__":
    main()
