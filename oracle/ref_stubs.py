"""Import the UNMODIFIED reference modules in the build container (test infrastructure).

``gym``, ``racecar_gym`` (and TensorFlow, for dreamer/tools.py) are not installed; the reference's
``dreamer/wrappers.py`` and ``baselines/racing/environment/single_agent.py`` only need a handful of names from
them at import time.  ``install()`` puts minimal stand-ins into ``sys.modules`` and returns the reference
module objects loaded straight from ``/root/reference`` (nothing is copied).  Used only by
``tests/golden/make_golden.py`` and by tests that are skipped when /root/reference is absent.
"""
from __future__ import annotations

import importlib.util
import sys
import types
from pathlib import Path

import numpy as np

REFERENCE_ROOT = Path("/root/reference")


class _Space:
    pass


class Box(_Space):
    def __init__(self, low, high, shape=None, dtype=np.float32):
        if shape is None:
            low = np.asarray(low, dtype=np.float64)
            high = np.asarray(high, dtype=np.float64)
            shape = low.shape
        else:
            low = np.full(shape, low, dtype=np.float64)
            high = np.full(shape, high, dtype=np.float64)
        self.low, self.high, self.shape, self.dtype = low.astype(dtype), high.astype(dtype), tuple(shape), dtype

    def sample(self):
        return np.random.uniform(self.low, self.high).astype(self.dtype)


class Dict(_Space):
    def __init__(self, spaces=None, **kw):
        if isinstance(spaces, (list, tuple)):
            spaces = dict(spaces)
        self.spaces = dict(spaces or {}, **kw)

    def __getitem__(self, k):
        return self.spaces[k]

    def sample(self):
        return {k: s.sample() for k, s in self.spaces.items()}


class Discrete(_Space):
    def __init__(self, n):
        self.n = n


class Wrapper:
    def __init__(self, env):
        self.env = env
        self.observation_space = getattr(env, "observation_space", None)
        self.action_space = getattr(env, "action_space", None)

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return getattr(self.env, name)

    def step(self, action):
        return self.env.step(action)

    def reset(self, **kwargs):
        return self.env.reset(**kwargs)

    def render(self, mode="human", **kwargs):
        return self.env.render(mode, **kwargs)


class ObservationWrapper(Wrapper):
    def reset(self, **kwargs):
        return self.observation(self.env.reset(**kwargs))

    def step(self, action):
        obs, r, d, i = self.env.step(action)
        return self.observation(obs), r, d, i


def _flatten_space(space):
    if isinstance(space, Dict):
        lows = [np.asarray(s.low, np.float64).ravel() for s in space.spaces.values()]
        highs = [np.asarray(s.high, np.float64).ravel() for s in space.spaces.values()]
        return Box(np.concatenate(lows), np.concatenate(highs), dtype=np.float64)
    return space


def _flatten(space, x):
    if isinstance(space, Dict):
        return np.concatenate([np.asarray(x[k], np.float64).ravel() for k in space.spaces])
    return np.asarray(x)


def _unflatten(space, x):
    if isinstance(space, Dict):
        out, i = {}, 0
        for k, s in space.spaces.items():
            n = int(np.prod(s.shape)) if s.shape else 1
            out[k] = np.asarray(x[i:i + n]).reshape(s.shape)
            i += n
        return out
    return x


def install():
    """Stub gym / racecar_gym in sys.modules (idempotent)."""
    if "gym" not in sys.modules or not hasattr(sys.modules["gym"], "_rd_stub"):
        gym = types.ModuleType("gym")
        gym._rd_stub = True
        spaces = types.ModuleType("gym.spaces")
        spaces.Box, spaces.Dict, spaces.Discrete = Box, Dict, Discrete
        spaces.flatten_space, spaces.flatten, spaces.unflatten = _flatten_space, _flatten, _unflatten
        gym.spaces = spaces
        gym.Wrapper, gym.ObservationWrapper = Wrapper, ObservationWrapper
        gym.Env = object
        sys.modules["gym"] = gym
        sys.modules["gym.spaces"] = spaces
    if "racecar_gym" not in sys.modules:
        rg = types.ModuleType("racecar_gym")
        rg.Task = object
        rg.register_task = lambda name, task: None
        envs = types.ModuleType("racecar_gym.envs")
        mar = types.ModuleType("racecar_gym.envs.multi_agent_race")
        mar.MultiAgentScenario = type("MultiAgentScenario", (), {})
        mar.MultiAgentRaceEnv = type("MultiAgentRaceEnv", (), {})
        rg.envs = envs
        envs.multi_agent_race = mar
        sys.modules["racecar_gym"] = rg
        sys.modules["racecar_gym.envs"] = envs
        sys.modules["racecar_gym.envs.multi_agent_race"] = mar


def _load(name: str, rel: str):
    path = REFERENCE_ROOT / rel
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def available() -> bool:
    return (REFERENCE_ROOT / "dreamer" / "wrappers.py").exists()


def reference_wrappers():
    """The reference's dreamer/wrappers.py, unmodified [REF dreamer/wrappers.py]."""
    install()
    return _load("_ref_dreamer_wrappers", "dreamer/wrappers.py")


def reference_baselines_env():
    """The reference's baselines single-agent wrappers, unmodified [REF baselines/racing/environment/single_agent.py]."""
    install()
    return _load("_ref_baselines_single_agent", "baselines/racing/environment/single_agent.py")


def reference_tasks():
    install()
    return _load("_ref_baselines_tasks", "baselines/racing/environment/tasks.py")
