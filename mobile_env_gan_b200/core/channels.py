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
from typing import Dict, Tuple

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

    # ---- coverage outline (rendering helper of the reference, channels.py:30-75, 86-127) ----
    @classmethod
    def boundary_collison(cls, theta: float, x0: float, y0: float, width: float, height: float) -> Tuple:
        """Where the ray leaving (x0, y0) at angle ``theta`` meets the map border (reference spelling
        and quadrant rules, channels.py:86-127: the candidate crossings are clamped to the map)."""
        half_pi = 1 / 2 * np.pi
        t, tq = np.tan(theta), np.tan(theta - half_pi)
        on_right = (width, t * (width - x0) + y0)
        on_top = ((-1) * tq * (height - y0) + x0, height)
        on_left = (0.0, t * (0.0 - x0) + y0)
        on_bottom = (tq * (y0 - 0.0) + x0, 0.0)
        axis = {0.0: (width, y0), half_pi: (x0, height), np.pi: (0.0, y0), 3 * half_pi: (x0, 0.0)}
        if theta in axis:
            return axis[theta]
        if 0.0 < theta < half_pi:
            return np.min((on_right[0], on_top[0], width)), np.min((on_right[1], on_top[1], height))
        if half_pi < theta < np.pi:
            return np.max((on_left[0], on_top[0], 0.0)), np.min((on_left[1], on_top[1], height))
        if np.pi < theta < 3 * half_pi:
            return np.max((on_left[0], on_bottom[0], 0.0)), np.max((on_left[1], on_bottom[1], 0.0))
        return np.min((on_right[0], on_bottom[0], width)), np.max((on_right[1], on_bottom[1], 0.0))

    def isoline(self, bs: BaseStation, ue_config: Dict, map_bounds: Tuple, dthresh: float, num: int = 32):
        """Outline of the area where a UE built from ``ue_config`` gets more than ``dthresh`` from
        ``bs``: along ``num`` rays the farthest of 100 sample points whose rate exceeds ``dthresh``
        (reference channels.py:30-75; rate at the integer-truncated sample point like ``ue.point``).
        Vectorised over the samples of a ray; raises ValueError like the reference when a ray has
        no such point."""
        width, height = map_bounds
        probe = UserEquipment(None, **ue_config)
        bx, by = int(bs.x), int(bs.y)
        outline_x, outline_y = [], []
        for theta in np.linspace(EPSILON, 2 * np.pi, num=num):
            x1, y1 = self.boundary_collison(theta, bs.x, bs.y, width, height)
            slope = (y1 - bs.y) / (x1 - bs.x)
            xs = np.linspace(bs.x, x1, num=100)
            ys = slope * (xs - bs.x) + bs.y
            dist = np.hypot(np.trunc(xs) - bx, np.trunc(ys) - by)
            rates = np.asarray([self.datarate(bs, probe, self.snr_at_distance(bs, probe, float(d))) for d in dist])
            (hit,) = np.where(rates > dthresh)
            far = np.max(hit)
            outline_x.append(xs[far])
            outline_y.append(ys[far])
        return tuple(outline_x), tuple(outline_y)

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
