"""Minimal ``gym.spaces`` stand-in (``Box``, ``Dict``) for the dict API of ``compat.py``.

The reference programs against ``gym.spaces.Box/Dict`` [REF dreamer/wrappers.py:43-60]; gym is not a dependency
of this package.  When a real ``gym`` (or ``gymnasium``) is importable its classes are used instead, so agents
that do ``isinstance(space, gym.spaces.Box)`` keep working.
"""
from __future__ import annotations

import numpy as np

try:  # pragma: no cover - depends on the host environment
    from gym.spaces import Box, Dict  # type: ignore
except Exception:  # noqa: BLE001
    try:  # pragma: no cover
        from gymnasium.spaces import Box, Dict  # type: ignore
    except Exception:  # noqa: BLE001

        class Box:  # type: ignore[no-redef]
            def __init__(self, low, high, shape=None, dtype=np.float32):
                if shape is None:
                    low, high = np.asarray(low, dtype=np.float64), np.asarray(high, dtype=np.float64)
                    shape = np.broadcast(low, high).shape
                self.shape = tuple(shape)
                self.dtype = np.dtype(dtype)
                self.low = np.broadcast_to(np.asarray(low, dtype=np.float64), self.shape).astype(self.dtype)
                self.high = np.broadcast_to(np.asarray(high, dtype=np.float64), self.shape).astype(self.dtype)

            def sample(self):
                lo = np.where(np.isfinite(self.low), self.low, -1.0)
                hi = np.where(np.isfinite(self.high), self.high, 1.0)
                return np.random.uniform(lo, hi).astype(self.dtype)

            def contains(self, x):
                x = np.asarray(x)
                return x.shape == self.shape and bool(np.all(x >= self.low) and np.all(x <= self.high))

            def __repr__(self):
                return f"Box({self.low.min()}, {self.high.max()}, {self.shape}, {self.dtype})"

        class Dict:  # type: ignore[no-redef]
            def __init__(self, spaces=None, **kwargs):
                self.spaces = dict(spaces or {}, **kwargs)

            def __getitem__(self, key):
                return self.spaces[key]

            def __iter__(self):
                return iter(self.spaces)

            def keys(self):
                return self.spaces.keys()

            def items(self):
                return self.spaces.items()

            def sample(self):
                return {k: s.sample() for k, s in self.spaces.items()}

            def __repr__(self):
                return f"Dict({self.spaces!r})"
