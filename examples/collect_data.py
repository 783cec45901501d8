#!/usr/bin/env python
"""The fork's data collection (collectData2.ipynb: `for e in epochs: reset(); for s in range(20):
step(e, s)` with the JSON/CSV dumps of base.py:298-404 / custom.py:79-85) on the GPU: env index =
epoch number, one launch per batch of epochs, files byte-compatible with the reference's.

    python examples/collect_data.py --epochs 2000 --out /tmp/collect --per-step
"""
import argparse
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mobile_env_gan_b200.export import ReferenceDumpWriter  # noqa: E402
from mobile_env_gan_b200.scenarios.custom import MComCustom  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--epochs", type=int, default=1024)
    ap.add_argument("--out", default="/tmp/mbe_collect")
    ap.add_argument("--per-step", action="store_true", help="also write the four JSON files of every step")
    args = ap.parse_args()
    E = (args.epochs + 31) // 32 * 32
    env = MComCustom(config={"num_envs": E})
    writer = ReferenceDumpWriter(env, args.out, envs=range(args.epochs), per_step=args.per_step)
    t0 = time.perf_counter()
    env.reset()
    writer.begin_episode()
    series = env.rollout(env.plan.ep_time, record=("pos", "wp", "assoc", "rate"))  # one launch
    t1 = time.perf_counter()
    writer.write_rollout(series)
    writer.end_episode()
    writer.close()
    t2 = time.perf_counter()
    files = sum(len(f) for _, _, f in os.walk(args.out))
    print(f"simulated {args.epochs} epochs in {t1 - t0:.3f} s; wrote {files} files in {t2 - t1:.1f} s -> {args.out}")


if __name__ == "__main__":
    main()
