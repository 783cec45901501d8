"""Gymnasium-shaped scenarios ``mobile-{small,medium,large}-{central,ma}-v0``.

The reference fork ships none of them (mobile_env/scenarios/__init__.py is empty); sizes follow
BASELINE.json (small 3 BS x 5 UE, large 13 BS x 30 UE) and SURVEY.md section 8 (medium 4 x 15).
Station coordinates are this build's own fixed integer layouts on the 200 x 200 map."""
from __future__ import annotations

from ..core.base import MComCore
from ..core.entities import BaseStation, UserEquipment


class _GymScenario(MComCore):
    STATIONS = ()
    NUM_UES = 0

    @classmethod
    def default_config(cls):
        config = super().default_config()
        config.update({"mode": "gym", "handler": "central"})
        return config

    def __init__(self, config=None, render_mode=None):
        config = config or {}
        base = self.default_config()
        bs_cfg = dict(base["bs"]); bs_cfg.update(config.get("bs", {}))
        ue_cfg = dict(base["ue"]); ue_cfg.update(config.get("ue", {}))
        stations = [BaseStation(i, pos, **bs_cfg) for i, pos in enumerate(self.STATIONS)]
        users = [UserEquipment(i, **ue_cfg) for i in range(self.NUM_UES)]
        super().__init__(stations, users, config, render_mode)


class MComSmall(_GymScenario):
    STATIONS = ((110, 130), (65, 80), (120, 30))
    NUM_UES = 5


class MComMedium(_GymScenario):
    STATIONS = ((50, 50), (150, 50), (50, 150), (150, 150))
    NUM_UES = 15


class MComLarge(_GymScenario):
    STATIONS = tuple((20 + 45 * (i % 4) + (22 if (i // 4) % 2 else 0), 25 + 50 * (i // 4)) for i in range(13))
    NUM_UES = 30


class MComSynthetic(_GymScenario):
    """Synthetic scale-up of BASELINE.json configs[4]: 64 BSs x 512 UEs on an 800 x 800 map with the
    ProportionalFair scheduler; BS coordinates are a fixed pseudo-random integer layout."""

    NUM_UES = 512
    SIZE = 800
    STATIONS = tuple(((i * 7919 + 13) % 800, (i * 104729 + 71) % 800) for i in range(64))

    @classmethod
    def default_config(cls):
        from ..core.schedules import ProportionalFair

        config = super().default_config()
        config.update({"width": cls.SIZE, "height": cls.SIZE, "scheduler": ProportionalFair})
        config["movement_params"].update({"width": cls.SIZE, "height": cls.SIZE})
        return config
