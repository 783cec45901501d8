#!/usr/bin/env python
"""Per-stage launches for ncu: the channel kernel (Channel.calculateSNR for every UE x BS pair) and
the four phases of the step as separate launches of the generic kernel (mbe_stage), on
mobile-medium-ma-v0 with 65,536 envs.  Run plain first, then under
  ncu --set full --clock-control none -k regex:'channel_kernel|step_kernel' -c 6 -o gpurun_out/prof_stages python profiles/stage_profile.py
Each launch is preceded by an L2 flush so that its inputs come from HBM."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mobile_env_gan_b200 as mbe  # noqa: E402

E = 65536
env = mbe.make("mobile-medium-ma-v0", num_envs=E, autoreset=True)
env.reset()
g = torch.Generator(device="cuda").manual_seed(0)
flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")
for k in range(3):  # a few fused steps so that connections exist
    env.step(torch.randint(0, 5, (E, 15), generator=g, device="cuda", dtype=torch.int32))
env.actions.copy_(torch.randint(0, 5, (E, 15), generator=g, device="cuda", dtype=torch.int32))
torch.cuda.synchronize()


def cold(fn):
    flush.zero_()
    torch.cuda.synchronize()
    fn()
    torch.cuda.synchronize()


cold(lambda: env.channel_snr(want_elig=True))   # channel_kernel: SNR matrix + connectable mask
for phase in (2, 1, 4, 8):                      # PRE, MOVE, CLOCK, POST (GYM order)
    cold(lambda: env.stage(phase))
print("stage profile done")
