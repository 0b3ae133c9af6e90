"""Test-only stand-in for `gymnasium.spaces` (gymnasium is not installed in this image).

Only what the *unmodified* PufferLib reference touches on the env-step / rollout path
(vector.py joint_space/check_envs, emulation.py check_space/dtype_from_space) is provided.
This is oracle infrastructure: it is never imported by the product package.
"""
import numpy as np


class Space:
    shape = None
    dtype = None

    def contains(self, x):
        raise NotImplementedError

    def __contains__(self, x):
        return self.contains(x)

    def seed(self, seed=None):
        self._rng = np.random.default_rng(seed)

    @property
    def np_random(self):
        if not hasattr(self, '_rng'):
            self._rng = np.random.default_rng()
        return self._rng


class Box(Space):
    def __init__(self, low, high, shape=None, dtype=np.float32):
        self.dtype = np.dtype(dtype)
        if shape is None:
            shape = np.broadcast(np.asarray(low), np.asarray(high)).shape
        self.shape = tuple(int(s) for s in shape)
        self.low = np.broadcast_to(np.asarray(low, dtype=self.dtype), self.shape).copy()
        self.high = np.broadcast_to(np.asarray(high, dtype=self.dtype), self.shape).copy()

    def contains(self, x):
        x = np.asarray(x)
        if x.shape != self.shape:
            return False
        if not np.can_cast(x.dtype, self.dtype) and x.dtype.kind != self.dtype.kind:
            return False
        return bool(np.all(x >= self.low) and np.all(x <= self.high))

    def sample(self):
        if self.dtype.kind == 'f':
            lo = np.where(np.isfinite(self.low), self.low, -1e3)
            hi = np.where(np.isfinite(self.high), self.high, 1e3)
            return self.np_random.uniform(lo, hi).astype(self.dtype)
        return self.np_random.integers(self.low, self.high.astype(np.int64) + 1).astype(self.dtype)

    def __eq__(self, other):
        return (isinstance(other, Box) and self.shape == other.shape and self.dtype == other.dtype
                and np.array_equal(self.low, other.low) and np.array_equal(self.high, other.high))

    def __repr__(self):
        return f'Box({self.low.min()}, {self.high.max()}, {self.shape}, {self.dtype})'


class Discrete(Space):
    def __init__(self, n, start=0):
        self.n = int(n)
        self.start = int(start)
        self.shape = ()
        self.dtype = np.dtype(np.int64)

    def contains(self, x):
        if isinstance(x, (int, np.integer)):
            v = int(x)
        elif isinstance(x, np.ndarray) and x.shape == () and x.dtype.kind in 'iu':
            v = int(x)
        else:
            return False
        return self.start <= v < self.start + self.n

    def sample(self):
        return int(self.np_random.integers(self.start, self.start + self.n))

    def __eq__(self, other):
        return isinstance(other, Discrete) and self.n == other.n and self.start == other.start

    def __repr__(self):
        return f'Discrete({self.n})'


class MultiDiscrete(Space):
    def __init__(self, nvec, dtype=np.int64):
        self.nvec = np.asarray(nvec, dtype=dtype)
        self.shape = self.nvec.shape
        self.dtype = np.dtype(dtype)

    def contains(self, x):
        x = np.asarray(x)
        return bool(x.shape == self.shape and x.dtype.kind in 'iu'
                    and np.all(x >= 0) and np.all(x < self.nvec))

    def sample(self):
        return self.np_random.integers(0, self.nvec).astype(self.dtype)

    def __len__(self):
        return len(self.nvec)

    def __eq__(self, other):
        return isinstance(other, MultiDiscrete) and np.array_equal(self.nvec, other.nvec)

    def __repr__(self):
        return f'MultiDiscrete({self.nvec.tolist() if self.nvec.size < 8 else self.nvec.shape})'


class MultiBinary(Space):
    def __init__(self, n):
        self.n = n
        self.shape = (n,) if isinstance(n, int) else tuple(n)
        self.dtype = np.dtype(np.int8)

    def contains(self, x):
        x = np.asarray(x)
        return bool(x.shape == self.shape and np.all((x == 0) | (x == 1)))

    def sample(self):
        return self.np_random.integers(0, 2, self.shape).astype(self.dtype)

    def __eq__(self, other):
        return isinstance(other, MultiBinary) and self.shape == other.shape


class Dict(Space):
    def __init__(self, spaces=None, **kw):
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

    def contains(self, x):
        return isinstance(x, dict) and all(k in x and s.contains(x[k]) for k, s in self.spaces.items())

    def sample(self):
        return {k: s.sample() for k, s in self.spaces.items()}

    def __eq__(self, other):
        return isinstance(other, Dict) and self.spaces == other.spaces


class Tuple(Space):
    def __init__(self, spaces):
        self.spaces = tuple(spaces)

    def __iter__(self):
        return iter(self.spaces)

    def __len__(self):
        return len(self.spaces)

    def __getitem__(self, i):
        return self.spaces[i]

    def contains(self, x):
        return isinstance(x, tuple) and len(x) == len(self.spaces) and all(
            s.contains(v) for s, v in zip(self.spaces, x))

    def sample(self):
        return tuple(s.sample() for s in self.spaces)

    def __eq__(self, other):
        return isinstance(other, Tuple) and self.spaces == other.spaces
