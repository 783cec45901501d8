"""Channel plugins.

The reference evaluates ``power_loss`` per (BS, UE) pair in Python every step
(mobile_env/core/channels.py:132-146 via 24-27).  Here a channel is *folded once* on the host
into per-BS-class constants that the CUDA kernels consume (include/mbe.h ``mbe_bs_class``):

* any loss that is affine in log10(distance), ``loss = a + c*log10(d + EPSILON)``, gives
  ``log2(snr) = l0 - k*log2(d^2)`` -- evaluated on the FP32 SFU pipes for the observations;
* because ``bs.point`` / ``ue.point`` are integer points (entities.py:24-26,52-54) d^2 is an
  integer, so ``snr > snr_threshold`` (base.py:212-214) is exactly ``d2 <= d2max`` and the
  Shannon rate of a link (channels.py:78-83) is a table indexed by d2.  Both are computed
  with the reference's own FP64 scalar operation order, which is what makes connection sets
  and rounded rates bit-exact on the GPU.

The scalar methods (``power_loss``, ``calculateSNR``, ``datarate``) keep the reference's names
and meaning; they are used for folding and for host-side inspection, never inside ``step``."""
from __future__ import annotations

import math
from abc import abstractmethod
from typing import Tuple

import numpy as np

from .entities import BaseStation, UserEquipment

EPSILON = 1e-16  # reference core/channels.py:8


class Channel:
    def __init__(self, **kwargs):
        pass

    def reset(self) -> None:
        pass

    # ---- reference-compatible scalar surface ------------------------------------------
    @abstractmethod
    def power_loss(self, bs: BaseStation, ue: UserEquipment) -> float:
        ...

    def calculateSNR(self, bs: BaseStation, ue: UserEquipment):
        return self.snr_at_distance(bs, ue, bs.point.distance(ue.point))

    snr = calculateSNR  # upstream spelling

    @classmethod
    def datarate(cls, bs: BaseStation, ue: UserEquipment, snr: float):
        if snr > ue.snr_threshold:
            return bs.bw * np.log2(1 + snr)
        return 0.0

    # ---- folding for the device --------------------------------------------------------
    @abstractmethod
    def log_distance_coefficients(self, bs: BaseStation, ue: UserEquipment) -> Tuple[float, float]:
        """(a, c) such that power_loss = a + c * log10(distance + EPSILON)."""

    def loss_at_distance(self, bs, ue, distance: float):
        a, c = self.log_distance_coefficients(bs, ue)
        return a + c * np.log10(distance + EPSILON)

    def snr_at_distance(self, bs, ue, distance: float):
        loss = self.loss_at_distance(bs, ue, distance)
        power = 10 ** ((bs.tx_power - loss) / 10)
        return power / ue.noise

    def fold(self, bs: BaseStation, ue: UserEquipment, max_d2: int) -> dict:
        """Constants of one (BS class, UE class) pair for the kernels."""
        a, c = (float(v) for v in self.log_distance_coefficients(bs, ue))
        log2_10 = math.log2(10.0)
        l0 = (bs.tx_power - a) / 10.0 * log2_10 - math.log2(ue.noise)
        k = c / 20.0
        l_zero = (bs.tx_power - (a + c * math.log10(EPSILON))) / 10.0 * log2_10 - math.log2(ue.noise)
        # connectable range: scan d2 upwards with the scalar FP64 chain of the reference
        lut = []
        d2 = 0
        while d2 <= max_d2:
            snr = self.snr_at_distance(bs, ue, math.sqrt(d2))
            if not (snr > ue.snr_threshold):
                break
            lut.append(float(self.datarate(bs, ue, snr)))
            d2 += 1
        d2max = d2 - 1
        # the scan assumes monotone loss; verify a margin beyond the cut-off
        for extra in range(d2max + 1, min(max_d2, d2max + 64) + 1):
            if self.snr_at_distance(bs, ue, math.sqrt(extra)) > ue.snr_threshold:
                raise ValueError("channel loss is not monotone in distance; cannot fold a range threshold")
        return {"l0": l0, "k": k, "l_zero": l_zero, "d2max": d2max, "rate_lut": np.asarray(lut, dtype=np.float64)}


class OkumuraHata(Channel):
    """Okumura-Hata urban path loss, reference core/channels.py:131-146."""

    def log_distance_coefficients(self, bs, ue):
        lf = np.log10(bs.frequency)
        ch = 0.8 + (1.1 * lf - 0.7) * ue.height - 1.56 * lf
        a = 69.55 - ch + 26.16 * lf - 13.82 * np.log10(bs.height)
        c = 44.9 - 6.55 * np.log10(bs.height)
        return a, c

    def power_loss(self, bs, ue):
        return self.loss_at_distance(bs, ue, bs.point.distance(ue.point))


class LogDistance(Channel):
    """Generic ``loss = a + c*log10(d)`` channel (covers the README's custom PathLoss example)."""

    def __init__(self, a: float = 40.0, c: float = 30.0, **kwargs):
        super().__init__(**kwargs)
        self.a, self.c = a, c

    def log_distance_coefficients(self, bs, ue):
        return self.a, self.c

    def power_loss(self, bs, ue):
        return self.loss_at_distance(bs, ue, bs.point.distance(ue.point))
