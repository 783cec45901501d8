"""Two env groups in flight on one handle (the double-buffered actor arrangement).

A single dependent chain of ~16 us step kernels pays launch ramp and drain on every step; when the
two halves of a batch are stepped alternately on two CUDA streams (``mbe_step_window``), the step of
one half overlaps the policy / drain of the other: 0.82 instead of 0.69 of the HBM roofline on
mobile-medium-central (``bench.py`` key ``two_env_groups_in_flight``).  Envs are independent objects
in the reference (``base.py:69-79``), so splitting a batch changes no result.
"""
from __future__ import annotations

from typing import Callable, List, Tuple

import torch

from . import _lib


class TwoGroupStepper:
    """``policy(obs_slice, group) -> actions_slice`` is called on the group's stream with the group's
    observation view ``[n, ...]`` and must return int32 actions ``[n, U]`` (a tensor on the device)."""

    def __init__(self, env, split: int = None):
        if env.plan.mode != _lib.MODE_GYM:
            raise ValueError("TwoGroupStepper needs a GYM-mode env (FORK episodes have no actions: use rollout())")
        E = env.num_envs
        cut = (E // 2 if split is None else int(split)) // 32 * 32
        if not 0 < cut < E:
            raise ValueError(f"cannot split {E} envs at a multiple of 32")
        self.env = env
        self.groups: List[Tuple[int, int]] = [(0, cut), (cut, E - cut)]
        self.streams = [torch.cuda.Stream(device=env.device) for _ in self.groups]

    def views(self, group: int):
        """(obs, reward, done) views of one group's slice of the env's tensors."""
        first, n = self.groups[group]
        env = self.env
        return env._obs_view()[first:first + n], env.reward[first:first + n], env.done[first:first + n]

    def step(self, group: int, actions):
        """Enqueues one step of ``group`` on its stream; returns its (obs, reward, done) views."""
        first, n = self.groups[group]
        env = self.env
        with torch.cuda.stream(self.streams[group]):
            env.actions[first:first + n].copy_(actions.reshape(n, -1), non_blocking=True)
            env.step_window(first, n, stream=self.streams[group])
        return self.views(group)

    def run(self, policy: Callable, steps: int):
        """``steps`` steps of both groups, alternating; returns after everything has finished."""
        cur = torch.cuda.current_stream(self.env.device)
        for st in self.streams:
            st.wait_stream(cur)
        for _ in range(steps):
            for g, st in enumerate(self.streams):
                with torch.cuda.stream(st):
                    obs, _, _ = self.views(g)
                    acts = policy(obs, g)
                self.step(g, acts)
        for st in self.streams:
            cur.wait_stream(st)
        cur.synchronize()
