"""Minimal stand-ins for the gymnasium pieces PickPlaceGymEnv uses (gymnasium is not installed in
this image): Box, Dict and an Env base with gymnasium's `reset(seed=...)` seeding
rule.  When gymnasium is importable the real classes are used instead."""
from __future__ import annotations

import numpy as np

try:  # pragma: no cover - not available in the build image
    import gymnasium as _gym
    from gymnasium import spaces as _spaces

    Box, Dict, Env = _spaces.Box, _spaces.Dict, _gym.Env
    HAVE_GYMNASIUM = True
except Exception:  # ModuleNotFoundError in this image
    HAVE_GYMNASIUM = False

    class Box:
        def __init__(self, low, high, shape=None, dtype=np.float32):
            self.dtype = np.dtype(dtype)
            if shape is None:
                shape = np.shape(low) if np.ndim(low) else np.shape(high)
            self.shape = tuple(shape)
            self.low = np.broadcast_to(np.asarray(low, dtype=self.dtype), self.shape).copy()
            self.high = np.broadcast_to(np.asarray(high, dtype=self.dtype), self.shape).copy()
            self._rng = np.random.default_rng()

        def seed(self, seed=None):
            self._rng = np.random.default_rng(seed)

        def sample(self):
            """gymnasium semantics: uniform on bounded dims, normal on unbounded ones."""
            if self.dtype.kind in "ui":
                return self._rng.integers(self.low, self.high.astype(np.int64) + 1, size=self.shape).astype(self.dtype)
            out = np.empty(self.shape, dtype=np.float64)
            lo_b, hi_b = np.isfinite(self.low), np.isfinite(self.high)
            both, neither = lo_b & hi_b, ~lo_b & ~hi_b
            out[both] = self._rng.uniform(self.low[both], self.high[both])
            out[neither] = self._rng.normal(size=int(neither.sum()))
            only_lo, only_hi = lo_b & ~hi_b, ~lo_b & hi_b
            out[only_lo] = self.low[only_lo] + self._rng.exponential(size=int(only_lo.sum()))
            out[only_hi] = self.high[only_hi] - self._rng.exponential(size=int(only_hi.sum()))
            return out.astype(self.dtype)

        def contains(self, x) -> bool:
            x = np.asarray(x)
            return x.shape == self.shape and bool(np.all(x >= self.low) and np.all(x <= self.high))

        def __repr__(self):
            return f"Box({self.low.min()}, {self.high.max()}, {self.shape}, {self.dtype})"

    class Dict:
        def __init__(self, spaces):
            self.spaces = dict(spaces)

        def __getitem__(self, k):
            return self.spaces[k]

        def keys(self):
            return self.spaces.keys()

        def sample(self):
            return {k: s.sample() for k, s in self.spaces.items()}

        def contains(self, x) -> bool:
            return set(x.keys()) == set(self.spaces.keys()) and all(s.contains(x[k]) for k, s in self.spaces.items())

    class Env:
        """gymnasium.Env seeding contract: reset(seed=s) re-creates np_random from s; without a seed
        the existing generator keeps advancing (created lazily from OS entropy)."""

        metadata: dict = {}
        _np_random = None

        @property
        def np_random(self) -> np.random.Generator:
            if self._np_random is None:
                self._np_random = np.random.default_rng()
            return self._np_random

        @np_random.setter
        def np_random(self, value):
            self._np_random = value

        def reset(self, *, seed=None, options=None):
            if seed is not None:
                self._np_random = np.random.default_rng(seed)

        def close(self):
            pass
