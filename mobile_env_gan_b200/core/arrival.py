"""Arrival plugins (reference mobile_env/core/arrival.py)."""
from __future__ import annotations


class Arrival:
    kernel_id = None

    def __init__(self, ep_time: int, seed: int, reset_rng_episode: bool, **kwargs):
        self.ep_time = ep_time
        self.seed = seed
        self.reset_rng_episode = reset_rng_episode

    def reset(self) -> None:
        pass


class NoDeparture(Arrival):
    """Every UE is present from t=0 to ep_time (arrival.py:28-36): the active mask is all-ones
    and ``done`` fires at ep_time."""

    kernel_id = 0

    def setArrivalTime(self, ue) -> int:
        return 0

    def setDepartureTime(self, ue) -> int:
        return self.ep_time

    arrival, departure = setArrivalTime, setDepartureTime  # upstream spellings
