"""Utility (QoE) plugins.

On the GPU a utility is a kernel id plus folded constants (``device_params``); the scalar methods
keep the names and meaning of the reference's ``mobile_env/core/utilities.py`` so that host code
(exports, tests, user scripts) can evaluate single values, and carry upstream's spellings too."""
from __future__ import annotations

import math
from typing import Sequence

import numpy as np


class Utility:
    kernel_id = None

    def __init__(self, **kwargs):
        pass

    def reset(self) -> None:
        pass

    def device_params(self) -> dict:
        raise NotImplementedError(f"{type(self).__name__} has no CUDA kernel")


class BoundedLogUtility(Utility):
    """QoE = log of the data rate, clipped to ``[lower, upper]`` and mapped affinely onto [-1, 1]
    (reference utilities.py:30-58; defaults lower=-20, upper=20, coeffs=(10, 0, 10), i.e.
    ``10*log10(rate)`` in dB, base.py:136)."""

    kernel_id = 0

    def __init__(self, lower: float, upper: float, coeffs: Sequence[float], **kwargs):
        super().__init__(**kwargs)
        if not upper > lower:
            raise ValueError("BoundedLogUtility needs upper > lower")
        self.lower, self.upper = lower, upper
        self.coeffs = tuple(coeffs)

    # -- folded form consumed by the kernels: u = clip(c * log2(w2 + r), lower, upper) -----------
    def device_params(self) -> dict:
        w1, w2, w3 = self.coeffs
        return {"c": w1 * math.log(2.0) / math.log(w3), "w2": w2, "lower": self.lower, "upper": self.upper,
                "scale": 2.0 / (self.upper - self.lower)}

    # -- scalar surface of the reference ---------------------------------------------------------
    def calculateUtility(self, datarate) -> float:
        if datarate <= 0.0:  # no service: the lower bound (utilities.py:46-47)
            return self.lower
        w1, w2, w3 = self.coeffs
        return np.clip(w1 * np.log(w2 + datarate) / np.log(w3), self.lower, self.upper)

    def scaleUtility(self, utility) -> float:
        span = self.upper - self.lower
        return 2 * (utility - self.lower) / span - 1

    def unscaleUtility(self, utility) -> float:
        span = self.upper - self.lower
        return (utility + 1) / 2 * span + self.lower

    utility, scale, unscale = calculateUtility, scaleUtility, unscaleUtility  # upstream spellings
