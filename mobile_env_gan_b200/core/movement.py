"""Movement plugins.  On the GPU a movement model is a kernel id plus parameters; waypoints
come from counter-based Philox4x32-10 keyed by (seed; env, ue, step) instead of the
reference's shared PCG64 stream (mobile_env/core/movement.py:16-18,44-47)."""
from __future__ import annotations

import math


class Movement:
    kernel_id = None

    def __init__(self, width: float, height: float, seed: int, reset_rng_episode: bool, **kwargs):
        self.width, self.height = width, height
        self.seed = seed
        self.reset_rng_episode = reset_rng_episode

    def reset(self) -> None:  # state lives on the device
        pass

    def device_params(self, velocity: float) -> dict:
        raise NotImplementedError(
            f"{type(self).__name__} has no CUDA kernel; only RandomWaypointMovement runs on the GPU "
            "(arbitrary Python move() cannot execute inside the step kernel)"
        )


class RandomWaypointMovement(Movement):
    """movement.py:30-72: integer waypoints, snap when within one step, otherwise a unit step
    scaled by ``velocity`` and rounded half-to-even."""

    kernel_id = 0

    def device_params(self, velocity: float) -> dict:
        v = float(velocity)
        # largest integer d2 with sqrt(d2) <= velocity, in the FP64 arithmetic np.linalg.norm uses
        n = max(int(v * v) + 2, 0)
        while n >= 0 and not (math.sqrt(n) <= v):
            n -= 1
        return {"velocity": v, "move_d2max": n}
