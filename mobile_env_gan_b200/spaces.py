"""Gymnasium spaces when gymnasium is installed, otherwise a minimal stand-in with the same
constructor arguments and attributes (gymnasium is not a dependency of this package)."""
from __future__ import annotations

import numpy as np

try:  # pragma: no cover - depends on the environment
    from gymnasium.spaces import Box, Dict, Discrete, MultiDiscrete  # noqa: F401

    HAVE_GYMNASIUM = True
except Exception:
    HAVE_GYMNASIUM = False

    class _Space:
        def __init__(self, shape, dtype):
            self.shape, self.dtype = tuple(shape), np.dtype(dtype)

        def __repr__(self):
            return f"{type(self).__name__}{self.shape}"

    class Box(_Space):
        def __init__(self, low, high, shape, dtype=np.float32):
            super().__init__(shape, dtype)
            self.low = np.full(shape, low, dtype=dtype)
            self.high = np.full(shape, high, dtype=dtype)

        def contains(self, x):
            x = np.asarray(x)
            return x.shape == self.shape and bool(np.all(x >= self.low) and np.all(x <= self.high))

        def sample(self):
            return np.random.uniform(self.low, self.high).astype(self.dtype)

    class Discrete(_Space):
        def __init__(self, n):
            super().__init__((), np.int64)
            self.n = int(n)

        def contains(self, x):
            return 0 <= int(x) < self.n

        def sample(self):
            return int(np.random.randint(self.n))

    class MultiDiscrete(_Space):
        def __init__(self, nvec):
            self.nvec = np.asarray(nvec, dtype=np.int64)
            super().__init__(self.nvec.shape, np.int64)

        def contains(self, x):
            x = np.asarray(x)
            return x.shape == self.shape and bool(np.all(x >= 0) and np.all(x < self.nvec))

        def sample(self):
            return (np.random.random(self.shape) * self.nvec).astype(np.int64)

    class Dict(dict):
        def __init__(self, spaces):
            super().__init__(spaces)
            self.spaces = self

        def contains(self, x):
            return set(x) == set(self) and all(self[k].contains(v) for k, v in x.items())

        def sample(self):
            return {k: s.sample() for k, s in self.items()}
