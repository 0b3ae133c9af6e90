"""Minimal Box / Discrete / MultiDiscrete with the attributes the reference reads from gymnasium spaces on this
path (vector.py:55-68 joint_space, :36-39 action check; clean_pufferl.py:40-42; models.py:26-36).
gymnasium is not a dependency of this package (and is not installed in the target image)."""
import numpy as np


class Box:
    def __init__(self, low, high, shape=None, dtype=np.float32):
        self.dtype = np.dtype(dtype)
        if shape is None:
            shape = np.broadcast(np.asarray(low), np.asarray(high)).shape
        self.shape = tuple(int(s) for s in shape)
        self.low = np.broadcast_to(np.asarray(low, dtype=self.dtype), self.shape)
        self.high = np.broadcast_to(np.asarray(high, dtype=self.dtype), self.shape)

    def contains(self, x):
        x = np.asarray(x)
        return bool(x.shape == self.shape and np.all(x >= self.low) and np.all(x <= self.high))

    def sample(self):
        rng = np.random.default_rng()
        if self.dtype.kind == 'f':
            return rng.uniform(self.low, self.high).astype(self.dtype)
        return rng.integers(self.low, self.high.astype(np.int64) + 1).astype(self.dtype)

    def __eq__(self, other):
        return (isinstance(other, Box) and self.shape == other.shape and self.dtype == other.dtype
                and np.array_equal(self.low, other.low) and np.array_equal(self.high, other.high))

    def __repr__(self):
        return f'Box({self.low.min()}, {self.high.max()}, {self.shape}, {self.dtype})'


class Discrete:
    def __init__(self, n):
        self.n = int(n)
        self.shape = ()
        self.dtype = np.dtype(np.int64)

    def contains(self, x):
        x = np.asarray(x)
        return bool(x.shape == () and x.dtype.kind in 'iu' and 0 <= int(x) < self.n)

    def sample(self):
        return int(np.random.default_rng().integers(0, self.n))

    def __eq__(self, other):
        return isinstance(other, Discrete) and self.n == other.n

    def __repr__(self):
        return f'Discrete({self.n})'


class MultiDiscrete:
    def __init__(self, nvec, dtype=np.int64):
        self.nvec = np.asarray(nvec, dtype=dtype)
        self.shape = self.nvec.shape
        self.dtype = np.dtype(dtype)

    def contains(self, x):
        x = np.asarray(x)
        return bool(x.shape == self.shape and x.dtype.kind in 'iu' and np.all(x >= 0) and np.all(x < self.nvec))

    def sample(self):
        return np.random.default_rng().integers(0, self.nvec).astype(self.dtype)

    def __len__(self):
        return len(self.nvec)

    def __eq__(self, other):
        return isinstance(other, MultiDiscrete) and np.array_equal(self.nvec, other.nvec)

    def __repr__(self):
        return f'MultiDiscrete(n={self.nvec[0] if self.nvec.size else 0} x {self.nvec.size})'


class Dict:
    """Ordered mapping of sub-spaces (gymnasium.spaces.Dict surface used by emulation.dtype_from_space)."""

    def __init__(self, spaces=None, **kw):
        # gymnasium.spaces.Dict sorts the keys of a plain dict and keeps the order of an OrderedDict; the structured
        # dtype layout (emulation.dtype_from_space) follows this order
        from collections import OrderedDict
        if isinstance(spaces, dict) and not isinstance(spaces, OrderedDict):
            try:
                spaces = dict(sorted(spaces.items()))
            except TypeError:
                pass
        self.spaces = dict(spaces or {}, **kw)

    def items(self):
        return self.spaces.items()

    def values(self):
        return self.spaces.values()

    def keys(self):
        return self.spaces.keys()

    def __getitem__(self, k):
        return self.spaces[k]

    def __iter__(self):
        return iter(self.spaces)

    def sample(self):
        return {k: s.sample() for k, s in self.spaces.items()}

    def contains(self, x):
        return isinstance(x, dict) and all(k in x and s.contains(x[k]) for k, s in self.spaces.items())

    def __eq__(self, other):
        return isinstance(other, Dict) and self.spaces == other.spaces


class Tuple:
    def __init__(self, spaces):
        self.spaces = tuple(spaces)

    def __iter__(self):
        return iter(self.spaces)

    def __len__(self):
        return len(self.spaces)

    def __getitem__(self, i):
        return self.spaces[i]

    def sample(self):
        return tuple(s.sample() for s in self.spaces)

    def contains(self, x):
        return isinstance(x, tuple) and len(x) == len(self.spaces) and all(
            s.contains(v) for s, v in zip(self.spaces, x))

    def __eq__(self, other):
        return isinstance(other, Tuple) and self.spaces == other.spaces
