"""Scheduler plugins (reference mobile_env/core/schedules.py)."""
from __future__ import annotations

from typing import List


class Scheduler:
    kernel_id = None

    def __init__(self, **kwargs):
        pass

    def reset(self) -> None:
        pass

    def share(self, bs, rates: List[float]) -> List[float]:
        raise NotImplementedError


class ResourceFair(Scheduler):
    """Equal split of the BS's resources: rate / n (schedules.py:20-22)."""

    kernel_id = 0

    def share(self, bs, rates):
        n = len(rates)
        return [r / n for r in rates]


class ProportionalFair(Scheduler):
    """Not in the fork (its schedules.py has ResourceFair and a broken RateFair only); this build's
    definition: a UE receives the fraction r_u / sum(r) of its BS, i.e. ``r_u * r_u / total``, with
    ``total`` accumulated in 2^-20 fixed point so that it does not depend on the summation order
    (see oracle/mbe_oracle.py:pf_total).  Parity unpinned."""

    kernel_id = 1

    def share(self, bs, rates):
        total = float(sum(int(round(float(r) * 2.0**20)) for r in rates)) * 2.0**-20
        return [r * r / total for r in rates]


class RateFair(Scheduler):
    """Every UE of a BS receives the same rate ``1 / sum(1 / r_i)``.  The fork computes this scalar
    but returns it instead of a list (schedules.py:26-29), which allocateDataRate2User (base.py:435)
    cannot consume; this is the repaired form, with the sum of inverse rates accumulated in 2^-50
    fixed point (order independent, see oracle/mbe_oracle.py:rate_fair_share).  Parity unpinned."""

    kernel_id = 2

    def share(self, bs, rates):
        total = float(sum(int(round(2.0**50 / float(r))) for r in rates)) * 2.0**-50
        return [1.0 / total for _ in rates]
