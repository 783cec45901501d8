"""Arrival plugins (reference mobile_env/core/arrival.py)."""
from __future__ import annotations

import numpy as np


class Arrival:
    kernel_id = None

    def __init__(self, ep_time: int, seed: int, reset_rng_episode: bool, **kwargs):
        self.ep_time = ep_time
        self.seed = seed
        self.reset_rng_episode = reset_rng_episode
        self.rng = None

    def reset(self) -> None:
        if self.reset_rng_episode or self.rng is None:  # arrival.py:14-16
            self.rng = np.random.default_rng(self.seed)

    def setArrivalTime(self, ue) -> int:
        raise NotImplementedError

    def setDepartureTime(self, ue) -> int:
        raise NotImplementedError


class NoDeparture(Arrival):
    """Every UE is present from t=0 to ep_time (arrival.py:28-36): the active mask is all-ones
    and ``done`` fires at ep_time."""

    kernel_id = 0

    def setArrivalTime(self, ue) -> int:
        return 0

    def setDepartureTime(self, ue) -> int:
        return self.ep_time

    arrival, departure = setArrivalTime, setDepartureTime  # upstream spellings
