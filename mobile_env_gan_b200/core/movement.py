"""Movement plugins.  On the GPU a movement model is a kernel id plus parameters; waypoints
come from counter-based Philox4x32-10 keyed by (seed; env, ue, step) instead of the
reference's shared PCG64 stream (mobile_env/core/movement.py:16-18,44-47)."""
from __future__ import annotations

import math

import numpy as np


class Movement:
    kernel_id = None

    def __init__(self, width: float, height: float, seed: int, reset_rng_episode: bool, **kwargs):
        self.width, self.height = width, height
        self.seed = seed
        self.reset_rng_episode = reset_rng_episode
        self.rng = None

    def reset(self) -> None:
        """The batched state lives on the device; the host generator below only serves the scalar
        per-entity methods (``move`` / ``initial_position``), re-seeded like movement.py:16-18."""
        if self.reset_rng_episode or self.rng is None:
            self.rng = np.random.default_rng(self.seed)

    def move(self, ue):
        raise NotImplementedError

    def initial_position(self, ue):
        raise NotImplementedError

    def device_params(self, velocity: float) -> dict:
        raise NotImplementedError(
            f"{type(self).__name__} has no CUDA kernel; only RandomWaypointMovement runs on the GPU "
            "(arbitrary Python move() cannot execute inside the step kernel)"
        )


class RandomWaypointMovement(Movement):
    """movement.py:30-72: integer waypoints, snap when within one step, otherwise a unit step
    scaled by ``velocity`` and rounded half-to-even."""

    kernel_id = 0

    def __init__(self, **kwargs):
        super().__init__(**kwargs)
        self.userMoveDirection = {}    # UE -> waypoint it is heading for (the reference's attribute names)
        self.userPositionInitial = {}  # UE -> the initial position it drew

    def reset(self) -> None:
        super().reset()
        self.userMoveDirection, self.userPositionInitial = {}, {}

    def _draw_point(self):
        # two uniform draws, x first, truncated to integers (movement.py:45-46, 67-68)
        x = int(self.rng.uniform(0, self.width))
        return x, int(self.rng.uniform(0, self.height))

    def move(self, ue):
        """The reference's scalar per-entity step (movement.py:42-62) on the host generator -- the plugin
        method ``MComCore.step`` calls per UE at base.py:233.  For inspection and single-entity use; the
        batched step runs the same arithmetic in the kernels with Philox draws."""
        if self.rng is None:
            self.reset()
        if ue not in self.userMoveDirection:
            self.userMoveDirection[ue] = self._draw_point()
        here = np.array([ue.x, ue.y])
        target = np.array(self.userMoveDirection[ue])
        if np.linalg.norm(here - target) <= ue.velocity:
            return self.userMoveDirection.pop(ue)  # arrived: snap onto the waypoint, draw a new one next time
        heading = target - here
        return tuple(np.round(here + ue.velocity * heading / np.linalg.norm(heading)).astype(int))

    def initial_position(self, ue):
        """movement.py:64-72: one uniform integer point per UE and episode, remembered."""
        if self.rng is None:
            self.reset()
        if ue not in self.userPositionInitial:
            self.userPositionInitial[ue] = self._draw_point()
        return self.userPositionInitial[ue]

    def device_params(self, velocity: float) -> dict:
        v = float(velocity)
        # largest integer d2 with sqrt(d2) <= velocity, in the FP64 arithmetic np.linalg.norm uses
        n = max(int(v * v) + 2, 0)
        while n >= 0 and not (math.sqrt(n) <= v):
            n -= 1
        return {"velocity": v, "move_d2max": n}
