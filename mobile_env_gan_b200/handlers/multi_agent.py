"""Multi-agent handler: every UE is an agent.  Action ``Dict{ue: Discrete(B+1)}``, observation
``Dict{ue: Box(-1, 1, (4B+1,))}`` = central features + broadcast BS utilities (B) + broadcast BS
connection counts (B); reward per UE = (own utility + sum of the utilities of the connectable
BSs) / (1 + their connection counts), with ``allStationUtilities`` (core/base.py:438-447) and
``available_connections`` (216-218) of the fork.  On the device the batch is dense: actions
[E,U], observations [E,U,4B+1], rewards [E,U]."""
from __future__ import annotations

import numpy as np

from .. import spaces


class MComMAHandler:
    features = ["connections", "snrs", "utility", "bcast", "stations_connected"]
    kernel_id = 1

    @classmethod
    def ue_obs_size(cls, env) -> int:
        return 4 * env.NUM_STATIONS + 1

    @classmethod
    def action_space(cls, env):
        return spaces.Dict({ue: spaces.Discrete(env.NUM_STATIONS + 1) for ue in sorted(env.userDict)})

    @classmethod
    def observation_space(cls, env):
        box = spaces.Box(-1.0, 1.0, (cls.ue_obs_size(env),), np.float32)
        return spaces.Dict({ue: box for ue in sorted(env.userDict)})
