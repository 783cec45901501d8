"""Central handler: one agent controls every UE.  The fork has no ``handlers`` package (SURVEY.md
section 0); the shapes follow upstream mobile-env: action ``MultiDiscrete([B+1]*U)`` (0 = NOOP,
core/base.py:29), observation ``Box(-1, 1, (U*(2B+1),))`` = per UE [connections one-hot (B),
snr / max snr (B), scaled utility (1)], reward = mean scaled utility (metrics.py:25-28).  The
features themselves are written by the POST phase of the step kernels."""
from __future__ import annotations

import numpy as np

from .. import spaces


class MComCentralHandler:
    features = ["connections", "snrs", "utility"]
    kernel_id = 0

    @classmethod
    def ue_obs_size(cls, env) -> int:
        return 2 * env.NUM_STATIONS + 1

    @classmethod
    def action_space(cls, env):
        return spaces.MultiDiscrete([env.NUM_STATIONS + 1] * env.NUM_USERS)

    @classmethod
    def observation_space(cls, env):
        return spaces.Box(-1.0, 1.0, (env.NUM_USERS * cls.ue_obs_size(env),), np.float32)
