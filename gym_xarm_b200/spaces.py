"""gym-compatible spaces.  Uses gymnasium or gym when importable (so SB3 type checks pass); otherwise a minimal
stand-in with the same attributes (`low`, `high`, `shape`, `dtype`, `sample`, `contains`, `spaces`)."""
import numpy as np

try:  # pragma: no cover - depends on the environment
    import gymnasium as _gym
    from gymnasium import spaces as _spaces
    Box, Dict = _spaces.Box, _spaces.Dict
    BACKEND = "gymnasium"
except Exception:  # noqa: BLE001
    try:  # pragma: no cover
        import gym as _gym
        from gym import spaces as _spaces
        Box, Dict = _spaces.Box, _spaces.Dict
        BACKEND = "gym"
    except Exception:  # noqa: BLE001
        _gym = None
        BACKEND = None

        class Box:
            def __init__(self, low, high, shape=None, dtype=np.float32):
                self.dtype = np.dtype(dtype)
                if shape is None:
                    shape = np.shape(low)
                self.shape = tuple(shape)
                self.low = np.broadcast_to(np.asarray(low, dtype=self.dtype), self.shape).copy()
                self.high = np.broadcast_to(np.asarray(high, dtype=self.dtype), self.shape).copy()
                self._rng = np.random.default_rng()

            def seed(self, seed=None):
                self._rng = np.random.default_rng(seed)
                return [seed]

            def sample(self):
                lo = np.where(np.isfinite(self.low), self.low, -1.0)
                hi = np.where(np.isfinite(self.high), self.high, 1.0)
                return self._rng.uniform(lo, hi).astype(self.dtype)

            def contains(self, x):
                x = np.asarray(x)
                return x.shape == self.shape and np.can_cast(x.dtype, self.dtype) and bool(np.all(x >= self.low) and np.all(x <= self.high))

            def __contains__(self, x):
                return self.contains(x)

            def __repr__(self):
                return f"Box({self.low.min()}, {self.high.max()}, {self.shape}, {self.dtype})"

        class Dict:
            def __init__(self, spaces):
                self.spaces = dict(spaces)

            def __getitem__(self, k):
                return self.spaces[k]

            def sample(self):
                return {k: s.sample() for k, s in self.spaces.items()}

            def contains(self, x):
                return isinstance(x, dict) and set(x) == set(self.spaces) and all(self.spaces[k].contains(x[k]) for k in x)

            def __contains__(self, x):
                return self.contains(x)

            def __repr__(self):
                return "Dict(" + ", ".join(f"{k}: {v}" for k, v in self.spaces.items()) + ")"
