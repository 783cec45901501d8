#!/usr/bin/env python
"""Target for the ncu capture of the fused-episode kernel (mbe_rollout on the fork's own scenario):
   python profiles/rollout_profile.py            # plain run first (must exit 0)
   ncu --set full --clock-control none --import-source on -k regex:step_tpe_fork -c 2 \
       -o gpurun_out/prof_rollout python profiles/rollout_profile.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mobile_env_gan_b200.scenarios.custom import MComCustom  # noqa: E402
from mobile_env_gan_b200.scoring import LayoutScorer  # noqa: E402

E = int(os.environ.get("ROLLOUT_ENVS", 262144))
env = MComCustom(config={"num_envs": E, "autoreset": True})
scorer = LayoutScorer(env)
env.reset()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
scorer.run_episode()
torch.cuda.synchronize()
e0.record()
for _ in range(4):
    scorer.run_episode()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 4
print(f"fused episode: {E} envs x 20 steps in {ms * 1e3:.1f} us -> {E * 20 / ms * 1e3:.3e} env-steps/s; "
      f"best score {float(scorer.result()['Score'].max()):.4f}")
