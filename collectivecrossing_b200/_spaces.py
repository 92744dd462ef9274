"""Observation / action space objects.  When gymnasium is installed its classes are used (so RLlib
sees real ``gymnasium.spaces``); otherwise minimal stand-ins with the attributes the reference's
callers touch (``n``, ``shape``, ``dtype``, ``low``, ``high``, ``sample``, ``contains``)."""

from __future__ import annotations

import numpy as np

try:  # pragma: no cover - depends on the environment
    from gymnasium.spaces import Box, Discrete  # type: ignore

    HAVE_GYMNASIUM = True
except Exception:  # gymnasium absent (this image): lightweight equivalents
    HAVE_GYMNASIUM = False

    class Discrete:  # type: ignore[no-redef]
        def __init__(self, n: int, seed: int | None = None):
            self.n, self.shape, self.dtype = int(n), (), np.int64
            self._rng = np.random.default_rng(seed)

        def sample(self) -> int:
            return int(self._rng.integers(0, self.n))

        def contains(self, v) -> bool:
            try:
                return 0 <= int(v) < self.n and int(v) == v
            except (TypeError, ValueError):
                return False

        __contains__ = contains

        def __repr__(self) -> str:
            return f"Discrete({self.n})"

        def __eq__(self, other) -> bool:
            return isinstance(other, Discrete) and other.n == self.n

    class Box:  # type: ignore[no-redef]
        def __init__(self, low, high, shape=None, dtype=np.float32, seed: int | None = None):
            self.shape, self.dtype = tuple(shape), np.dtype(dtype)
            self.low = np.full(self.shape, low, dtype=self.dtype)
            self.high = np.full(self.shape, high, dtype=self.dtype)
            self._rng = np.random.default_rng(seed)

        def sample(self) -> np.ndarray:
            return self._rng.uniform(self.low, self.high).astype(self.dtype)

        def contains(self, v) -> bool:
            v = np.asarray(v)
            return v.shape == self.shape and bool(np.all(v >= self.low) and np.all(v <= self.high))

        __contains__ = contains

        def __repr__(self) -> str:
            return f"Box({self.low.flat[0]}, {self.high.flat[0]}, {self.shape}, {self.dtype})"

        def __eq__(self, other) -> bool:
            return isinstance(other, Box) and other.shape == self.shape and np.array_equal(other.low, self.low) and np.array_equal(other.high, self.high)
