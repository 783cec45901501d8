"""MComCustom, batched (reference mobile_env/scenarios/custom.py:12-85): 7 UEs at velocity 10
and, per env and per episode, 5..10 base stations at uniform integer positions.  The reference
draws the layout from the unseeded global ``random`` (custom.py:68-77); here it is Philox keyed
by (seed, env, episode), so runs are reproducible.  UE trajectories follow the fork: with the default
``movement_params.reset_rng_episode=True`` every epoch of the reference replays ONE UE trajectory
(base.py:130-134), so with env index = epoch number all envs share it (``shared_trajectory``) and only
the BS layouts differ -- which is what makes the layout scores comparable."""
from __future__ import annotations

from ..core.base import MComCore
from ..core.entities import UserEquipment


class MComCustom(MComCore):
    NUM_UES = 7
    BS_RANGE = (5, 10)

    @classmethod
    def default_config(cls):
        config = super().default_config()
        config["ue"].update({"velocity": 10})
        config.update({"bs_random": cls.BS_RANGE, "max_bs": cls.BS_RANGE[1], "mode": "fork",
                       "shared_trajectory": "follow_movement",
                       # one env = the reference's own use (collectData2.ipynb): step() dumps like base.py:261
                       "dumps": None})
        return config

    def __init__(self, config=None, render_mode=None):
        config = config or {}
        ue_cfg = dict(self.default_config()["ue"])
        ue_cfg.update(config.get("ue", {}))
        users = [UserEquipment(ue_id=i, **ue_cfg) for i in range(self.NUM_UES)]
        super().__init__([], users, config, render_mode)
