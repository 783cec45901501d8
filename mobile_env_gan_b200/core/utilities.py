"""Utility plugins (reference mobile_env/core/utilities.py)."""
from __future__ import annotations

from typing import Tuple

import numpy as np


class Utility:
    kernel_id = None

    def __init__(self, **kwargs):
        pass

    def reset(self) -> None:
        pass


class BoundedLogUtility(Utility):
    """clip(w1*log(w2+rate)/log(w3), lower, upper), then scaled to [-1, 1] (utilities.py:30-58)."""

    kernel_id = 0

    def __init__(self, lower: float, upper: float, coeffs: Tuple[float, float, float], **kwargs):
        super().__init__(**kwargs)
        self.lower, self.upper, self.coeffs = lower, upper, coeffs

    def calculateUtility(self, datarate) -> float:
        w1, w2, w3 = self.coeffs
        if datarate <= 0.0:
            return self.lower
        return np.clip(w1 * np.log(w2 + datarate) / np.log(w3), self.lower, self.upper)

    def scaleUtility(self, utility) -> float:
        return 2 * (utility - self.lower) / (self.upper - self.lower) - 1

    def unscaleUtility(self, utility) -> float:
        return (utility + 1) / 2 * (self.upper - self.lower) + self.lower

    utility, scale, unscale = calculateUtility, scaleUtility, unscaleUtility  # upstream spellings
