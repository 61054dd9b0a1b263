"""Minimal gym.spaces stand-ins (gym is an external dependency of the reference and is absent here).

Only what the learner boundary needs: `.shape`, `.dtype`, `.n`, `.sample()`.  Real
gym spaces are accepted everywhere too: the code duck-types on these attributes.
"""
import numpy as np


class Box:
    def __init__(self, low, high, shape=None, dtype=np.float32):
        self.dtype = np.dtype(dtype)
        if shape is None:
            low, high = np.asarray(low), np.asarray(high)
            shape = low.shape
        self.shape = tuple(shape)
        self.low = np.broadcast_to(np.asarray(low, dtype=self.dtype), self.shape).copy()
        self.high = np.broadcast_to(np.asarray(high, dtype=self.dtype), self.shape).copy()

    def seed(self, seed=None):
        self._rng = np.random.RandomState(seed)

    def sample(self):
        rng = getattr(self, "_rng", None) or np.random
        if self.dtype.kind in "iu":
            return rng.randint(self.low.astype(np.int64), self.high.astype(np.int64) + 1).astype(self.dtype)
        return rng.uniform(self.low, self.high).astype(self.dtype)

    def contains(self, x):
        x = np.asarray(x)
        return x.shape == self.shape and bool(np.all(x >= self.low) and np.all(x <= self.high))

    def __repr__(self):
        return f"Box{self.shape}"


class Discrete:
    def __init__(self, n):
        self.n = int(n)
        self.shape = ()
        self.dtype = np.dtype(np.int64)

    def seed(self, seed=None):
        self._rng = np.random.RandomState(seed)

    def sample(self):
        return int((getattr(self, "_rng", None) or np.random).randint(self.n))

    def contains(self, x):
        return 0 <= int(x) < self.n

    def __repr__(self):
        return f"Discrete({self.n})"


def is_discrete(space):
    return hasattr(space, "n") and not hasattr(space, "nvec")


def is_box(space):
    return hasattr(space, "low") and hasattr(space, "high")
