"""Layout scoring: the fork searches BS layouts by running an episode per
random layout and ranking them with ``qoeValue`` (mobile_env/chooseBaseStation.ipynb cell 5):

    score = w1 * mean(qoe) - w2 * var(qoe) - w3 * P(qoe < low_qoe_threshold)

over all two-decimal per-UE QoE values of the epoch (base.py:269).  With one env per layout the
whole search is a batch: ``LayoutScorer`` accumulates the statistics on the device after every
step (``mbe_accumulate_qoe``) and ranks the envs without any file round trip."""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib


class LayoutScorer:
    def __init__(self, env, low_qoe_threshold: float = 0.0, weights=(1.0, 0.1, 10.0)):
        self.env, self.threshold, self.weights = env, float(low_qoe_threshold), weights
        self.acc = torch.zeros(env.num_envs, 4, dtype=torch.float32, device=env.device)

    def reset(self):
        self.acc.zero_()

    def update(self):
        """Call after every ``env.step``."""
        env = self.env
        _lib.check(env._lib.mbe_accumulate_qoe(env._handle, C.c_void_p(self.acc.data_ptr()),
                                               C.c_float(self.threshold), env._stream()))

    def run_episode(self, steps: int = None, record=()):
        """One whole episode through ``mbe_rollout`` with the statistics accumulated inside the
        kernel (one launch for the fork's own scenario) instead of ``step`` + ``update`` per step."""
        env = self.env
        return env.rollout(env.plan.ep_time if steps is None else steps, qoe_acc=self.acc,
                           threshold=self.threshold, record=record)

    def result(self):
        """Per-env dict of tensors with the notebook's keys."""
        s1, s2, neg, n = self.acc.unbind(dim=1)
        mean = s1 / n
        var = s2 / n - mean * mean  # np.var (population variance)
        low = neg / n
        w1, w2, w3 = self.weights
        return {"Average QoE": mean, "QoE Variance": var, "Low QoE Proportion": low,
                "Score": w1 * mean - w2 * var - w3 * low}

    def best(self, k: int = 1):
        """Indices (and scores) of the k best layouts, like the notebook's sort (cell 9)."""
        score = self.result()["Score"]
        top = torch.topk(score, k)
        return top.indices, top.values
