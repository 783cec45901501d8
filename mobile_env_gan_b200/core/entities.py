"""Entity records.  On the GPU path they only carry *parameters*: positions live in the
[E,U,2] / [B,2] tensors of the batched env (reference mobile_env/core/entities.py:6-57 keeps
x/y on the objects).  Constructor signatures match the reference."""
from __future__ import annotations

import math
from typing import Tuple


class IntPoint:
    """What ``bs.point`` / ``ue.point`` give in the reference (entities.py:24-26,52-54):
    the coordinates truncated to int, with a planar ``distance``."""

    __slots__ = ("x", "y")

    def __init__(self, x, y):
        self.x, self.y = int(x), int(y)

    def distance(self, other: "IntPoint") -> float:
        return math.hypot(self.x - other.x, self.y - other.y)


class BaseStation:
    def __init__(self, bs_id: int, pos: Tuple[float, float], bw: float, freq: float, tx: float, height: float):
        self.bs_id = bs_id
        self.x, self.y = pos
        self.bw = bw  # Hz
        self.frequency = freq  # MHz
        self.tx_power = tx  # dBm
        self.height = height  # m

    @property
    def point(self) -> IntPoint:
        return IntPoint(self.x, self.y)

    def radio_key(self):
        return (float(self.bw), float(self.frequency), float(self.tx_power), float(self.height))

    def __str__(self):
        return f"BS: {self.bs_id}"


class UserEquipment:
    def __init__(self, ue_id: int, velocity: float, snr_tr: float, noise: float, height: float):
        self.ue_id = ue_id
        self.velocity = velocity
        self.snr_threshold = snr_tr
        self.noise = noise
        self.height = height
        self.x = None
        self.y = None
        self.startTime = None
        self.exitTime = None

    @property
    def point(self) -> IntPoint:
        return IntPoint(self.x, self.y)

    def radio_key(self):
        return (self.velocity, float(self.snr_threshold), float(self.noise), float(self.height))

    def __str__(self):
        return f"UE: {self.ue_id}"
