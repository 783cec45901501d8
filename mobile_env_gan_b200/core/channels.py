"""Channel plugins.

The reference evaluates ``power_loss`` per (BS, UE) pair in Python every step
(mobile_env/core/channels.py:132-146 via 24-27).  Here a channel is *folded once* on the host
into per-link-class tables that the CUDA kernels consume (include/mbe.h ``mbe_link_class``).
Folding needs nothing but the reference's one abstract method, ``power_loss(bs, ue)``
(channels.py:18-21): ``bs.point`` / ``ue.point`` are integer points (entities.py:24-26,52-54), so a
radially symmetric loss is a function of the *integer* squared distance d2, and the folder
evaluates the subclass's own ``power_loss`` on probe entities placed at the integer offsets
(dx, dy) that realise each d2:

* ``snr > snr_threshold`` (base.py:212-214) becomes exactly ``d2 <= d2max``;
* the Shannon rate of a link (channels.py:78-83) becomes the FP64 table ``rate_lut[d2]``;
  both come out of the reference's own FP64 scalar operation order, which is what makes
  connection sets and rounded rates bit-exact on the GPU;
* ``log2(snr)`` for the observation path is either the affine form ``l0 - k*log2(d2)`` on the
  FP32 SFU pipes -- used only when a fit over the probes reproduces the subclass to 1e-7 -- or an
  FP32 table ``log2snr_lut[d2]`` over the whole map.

A subclass whose ``power_loss`` depends on more than the BS-UE distance (absolute position,
direction, state) cannot be folded and raises ``NotImplementedError``.  A subclass MAY override
``log_distance_coefficients`` to hand the folder the closed form directly (the built-ins do).

The scalar methods (``power_loss``, ``calculateSNR``, ``datarate``) keep the reference's names
and meaning; they are used for folding and for host-side inspection, never inside ``step``."""
from __future__ import annotations

import copy
import math
from abc import abstractmethod
from typing import Dict, Optional, Tuple

import numpy as np

from .entities import BaseStation, UserEquipment

EPSILON = 1e-16  # reference core/channels.py:8

# largest |log2 snr| handed to the FP32 kernels (d = 0 with a pure log-distance loss gives +inf)
_L_CLAMP = 1.0e30
# tolerance of the affine fit log2(snr) = l0 - k*log2(d2) (relative to the spread of the values)
_AFFINE_TOL = 1e-7
# a full scan of every realisable d2 is done up to this many probes; beyond it the connectable range
# is verified over a margin past the cut-off only
_FULL_SCAN_PROBES = 400_000


def _offsets(max_d2: int):
    """For every integer d2 <= max_d2 that is a sum of two squares, one (dx, dy) with dx >= dy >= 0
    and dx*dx + dy*dy == d2; dx = -1 where d2 is not realisable on the integer grid."""
    r = int(math.isqrt(max_d2))
    dx = np.full(max_d2 + 1, -1, dtype=np.int64)
    dy = np.zeros(max_d2 + 1, dtype=np.int64)
    xs = np.arange(r + 1, dtype=np.int64)
    for y in range(r + 1):
        x = xs[y:]
        d2 = x * x + y * y
        ok = d2 <= max_d2
        x, d2 = x[ok], d2[ok]
        if not len(x):
            break
        new = dx[d2] < 0
        dx[d2[new]] = x[new]
        dy[d2[new]] = y
    return dx, dy


class Channel:
    def __init__(self, **kwargs):
        pass

    def reset(self) -> None:
        pass

    # ---- reference-compatible scalar surface ------------------------------------------
    @abstractmethod
    def power_loss(self, bs: BaseStation, ue: UserEquipment) -> float:
        """The one method a channel model has to provide (reference channels.py:18-21)."""

    def calculateSNR(self, bs: BaseStation, ue: UserEquipment):
        """reference channels.py:24-27"""
        loss = self.power_loss(bs, ue)
        try:
            power = 10 ** ((bs.tx_power - loss) / 10)
        except OverflowError:  # a Python-float loss of -inf / -1e4: numpy would give inf
            power = math.inf
        return power / ue.noise

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
        (reference channels.py:30-75: a dummy UE is moved along the ray and ``calculateSNR`` /
        ``datarate`` are evaluated at each sample).  Raises ValueError like the reference when a ray
        has no such point."""
        width, height = map_bounds
        probe = UserEquipment(None, **ue_config)
        outline_x, outline_y = [], []
        with np.errstate(divide="ignore", over="ignore"):
            for theta in np.linspace(EPSILON, 2 * np.pi, num=num):
                x1, y1 = self.boundary_collison(theta, bs.x, bs.y, width, height)
                slope = (y1 - bs.y) / (x1 - bs.x)
                xs = np.linspace(bs.x, x1, num=100)
                ys = slope * (xs - bs.x) + bs.y
                rates = []
                for px, py in zip(xs.tolist(), ys.tolist()):
                    probe.x, probe.y = px, py
                    rates.append(self.datarate(bs, probe, self.calculateSNR(bs, probe)))
                (hit,) = np.where(np.asarray(rates) > dthresh)
                far = np.max(hit)
                outline_x.append(xs[far])
                outline_y.append(ys[far])
        return tuple(outline_x), tuple(outline_y)

    # ---- folding for the device --------------------------------------------------------
    def log_distance_coefficients(self, bs: BaseStation, ue: UserEquipment) -> Optional[Tuple[float, float]]:
        """Optional closed form: ``(a, c)`` with ``power_loss = a + c*log10(distance + EPSILON)``.
        ``None`` (the default) makes the folder derive everything from ``power_loss`` alone."""
        return None

    def _probe_snr(self, bs, ue, origin, dx: int, dy: int):
        """``calculateSNR`` of the subclass with the UE ``(dx, dy)`` away from the BS at ``origin``."""
        bs.x, bs.y = origin
        ue.x, ue.y = origin[0] + dx, origin[1] + dy
        return float(self.calculateSNR(bs, ue))

    def fold(self, bs: BaseStation, ue: UserEquipment, max_d2: int) -> dict:
        """Tables of one link class = one (BS parameter set, UE parameter set) pair for the kernels,
        derived from ``power_loss`` only.  ``max_d2``: largest squared distance on the map."""
        pbs, pue = copy.copy(bs), copy.copy(ue)
        origin = (int(bs.x), int(bs.y))
        odx, ody = _offsets(max_d2)
        real = np.flatnonzero(odx >= 0)  # realisable d2, ascending; real[0] == 0

        with np.errstate(divide="ignore", over="ignore", invalid="ignore"):
            self._check_radial(pbs, pue, origin, odx, ody, real)
            coeff = self.log_distance_coefficients(bs, ue)
            if coeff is not None and self._closed_form_matches(coeff, pbs, pue, origin, odx, ody, real):
                return self._fold_closed_form(bs, ue, max_d2, coeff)
            coeff = None  # e.g. a subclass that overrides power_loss only: the probes decide
            snr = {}

            def snr_at(d2: int) -> float:
                if d2 not in snr:
                    snr[d2] = self._probe_snr(pbs, pue, origin, int(odx[d2]), int(ody[d2]))
                return snr[d2]

            # connectable range: realisable d2 upwards with the subclass's own FP64 chain
            thr = ue.snr_threshold
            d2max, pos = -1, 0
            while pos < len(real) and snr_at(int(real[pos])) > thr:
                d2max = int(real[pos])
                pos += 1
            # the kernels test `d2 <= d2max`: the loss must be monotone past the cut-off
            full = len(real) <= _FULL_SCAN_PROBES
            stop = len(real) if full else min(len(real), pos + 64)
            for i in range(pos, stop):
                if snr_at(int(real[i])) > thr:
                    raise NotImplementedError(
                        f"{type(self).__name__}: the connectable set is not a disc (snr > snr_threshold again at "
                        f"d2={int(real[i])} beyond the cut-off d2={d2max}); cannot fold a range threshold")

            # Shannon rate per d2 (entries at non-realisable d2 are never indexed)
            rate_lut = np.zeros(d2max + 1, dtype=np.float64)
            for d2 in real[: pos]:
                rate_lut[int(d2)] = float(self.datarate(pbs, pue, snr_at(int(d2))))

            # log2 snr for the observation path
            def l_of(s: float) -> float:
                if not (s > 0.0):
                    return -_L_CLAMP
                return min(math.log2(s), _L_CLAMP) if math.isfinite(s) else _L_CLAMP

            l_zero = l_of(snr_at(0))
            out = {"d2max": d2max, "rate_lut": rate_lut, "l_zero": l_zero, "log2snr_lut": None}
            fit = self._affine_fit(snr_at, real, l_of)
            if fit is not None:
                out["l0"], out["k"] = fit
                return out
            # general radial loss: FP32 table over every d2 of the map (gaps interpolated in log2 d2)
            vals = np.array([l_of(snr_at(int(d2))) for d2 in real], dtype=np.float64)
            grid = np.arange(max_d2 + 1, dtype=np.float64)
            lut = np.interp(np.log2(np.maximum(grid, 0.5)), np.log2(np.maximum(real.astype(np.float64), 0.5)), vals)
            out["l0"], out["k"] = float(vals[min(1, len(vals) - 1)]), 0.0
            out["log2snr_lut"] = np.clip(lut, -3.0e38, 3.0e38).astype(np.float32)
            return out

    # closed form a + c*log10(d + EPSILON): every integer d2 can be evaluated, realisable or not
    @staticmethod
    def _snr_closed_form(coeff, bs, ue, distance: float):
        a, c = coeff
        loss = a + c * np.log10(distance + EPSILON)
        power = 10 ** ((bs.tx_power - loss) / 10)
        return power / ue.noise

    def _closed_form_matches(self, coeff, bs, ue, origin, odx, ody, real) -> bool:
        """Is ``log_distance_coefficients`` really this object's ``power_loss``?  (A subclass may
        override ``power_loss`` and inherit the coefficients of its parent.)"""
        sample = real[np.unique(np.linspace(0, len(real) - 1, num=min(48, len(real))).astype(int))]
        for d2 in sample:
            want = self._probe_snr(bs, ue, origin, int(odx[d2]), int(ody[d2]))
            got = float(self._snr_closed_form(coeff, bs, ue, math.sqrt(int(d2))))
            if not (got == want or (math.isfinite(want) and abs(got - want) <= 1e-9 * abs(want))):
                return False
        return True

    def _fold_closed_form(self, bs, ue, max_d2: int, coeff) -> dict:
        a, c = (float(v) for v in coeff)
        log2_10 = math.log2(10.0)
        l0 = (bs.tx_power - a) / 10.0 * log2_10 - math.log2(ue.noise)
        k = c / 20.0
        l_zero = (bs.tx_power - (a + c * math.log10(EPSILON))) / 10.0 * log2_10 - math.log2(ue.noise)
        # connectable range: scan d2 upwards with the scalar FP64 chain of the reference
        lut = []
        d2 = 0
        while d2 <= max_d2:
            snr = self._snr_closed_form(coeff, bs, ue, math.sqrt(d2))
            if not (snr > ue.snr_threshold):
                break
            lut.append(float(self.datarate(bs, ue, snr)))
            d2 += 1
        d2max = d2 - 1
        # the scan assumes monotone loss; verify a margin beyond the cut-off
        for extra in range(d2max + 1, min(max_d2, d2max + 64) + 1):
            if self._snr_closed_form(coeff, bs, ue, math.sqrt(extra)) > ue.snr_threshold:
                raise NotImplementedError("channel loss is not monotone in distance; cannot fold a range threshold")
        return {"l0": l0, "k": k, "l_zero": max(-_L_CLAMP, min(l_zero, _L_CLAMP)), "d2max": d2max,
                "rate_lut": np.asarray(lut, dtype=np.float64), "log2snr_lut": None}

    def _check_radial(self, bs, ue, origin, odx, ody, real) -> None:
        """``power_loss`` must depend on the integer offset through its length only, and not on where
        the pair sits: compare rotated / mirrored offsets and a shifted origin on a sample of d2."""
        sample = real[np.unique(np.linspace(0, len(real) - 1, num=min(24, len(real))).astype(int))]
        shifted = (origin[0] + 37, origin[1] + 11)
        for d2 in sample:
            dx, dy = int(odx[d2]), int(ody[d2])
            ref = self._probe_snr(bs, ue, origin, dx, dy)
            others = [self._probe_snr(bs, ue, origin, -dy, dx), self._probe_snr(bs, ue, origin, dy, -dx),
                      self._probe_snr(bs, ue, origin, -dx, -dy), self._probe_snr(bs, ue, shifted, dx, dy)]
            for o in others:
                same = (o == ref) or (math.isfinite(ref) and abs(o - ref) <= 1e-12 * abs(ref))
                if not same:
                    raise NotImplementedError(
                        f"{type(self).__name__}.power_loss depends on more than the BS-UE distance "
                        f"(offset ({dx},{dy}): {ref!r} vs {o!r}); only radially symmetric channels fold into the kernels")

    @staticmethod
    def _affine_fit(snr_at, real, l_of):
        """(l0, k) with log2 snr = l0 - k*log2(d2) when that holds to _AFFINE_TOL on a spread of probes
        (d2 >= 1), else None."""
        pts = real[real >= 1]
        if len(pts) < 2:
            return None
        pick = pts[np.unique(np.geomspace(1, len(pts), num=min(96, len(pts))).astype(int) - 1)]
        lg = np.log2(pick.astype(np.float64))
        l = np.array([l_of(snr_at(int(d2))) for d2 in pick])
        if not np.all(np.abs(l) < _L_CLAMP):
            return None
        k = -(l[-1] - l[0]) / (lg[-1] - lg[0]) if lg[-1] > lg[0] else 0.0
        l0 = l[0] + k * lg[0]
        err = np.max(np.abs((l0 - k * lg) - l))
        scale = max(1.0, float(np.max(np.abs(l))))
        return (float(l0), float(k)) if err <= _AFFINE_TOL * scale else None


class OkumuraHata(Channel):
    """Okumura-Hata urban path loss, reference core/channels.py:131-146."""

    def power_loss(self, bs, ue):
        distance = bs.point.distance(ue.point)
        ch = 0.8 + (1.1 * np.log10(bs.frequency) - 0.7) * ue.height - 1.56 * np.log10(bs.frequency)
        tmp_1 = 69.55 - ch + 26.16 * np.log10(bs.frequency) - 13.82 * np.log10(bs.height)
        tmp_2 = 44.9 - 6.55 * np.log10(bs.height)
        return tmp_1 + tmp_2 * np.log10(distance + EPSILON)

    def log_distance_coefficients(self, bs, ue):
        lf = np.log10(bs.frequency)
        ch = 0.8 + (1.1 * lf - 0.7) * ue.height - 1.56 * lf
        a = 69.55 - ch + 26.16 * lf - 13.82 * np.log10(bs.height)
        c = 44.9 - 6.55 * np.log10(bs.height)
        return a, c


class LogDistance(Channel):
    """Generic ``loss = a + c*log10(d + EPSILON)`` channel."""

    def __init__(self, a: float = 40.0, c: float = 30.0, **kwargs):
        super().__init__(**kwargs)
        self.a, self.c = a, c

    def log_distance_coefficients(self, bs, ue):
        return self.a, self.c

    def power_loss(self, bs, ue):
        return self.a + self.c * np.log10(bs.point.distance(ue.point) + EPSILON)
