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


def reference_baselines_common():
    """The reference's baselines/racing/environment/common.py, unmodified (FixedResetMode, InfoToObservation)."""
    install()
    return _load("_ref_baselines_common", "baselines/racing/environment/common.py")


class FilterObservation(ObservationWrapper):
    """Restatement of gym.wrappers.FilterObservation (gym==0.18.0, pinned at baselines/docker/requirements_acme.txt:40;
    gym itself is not in this image): keeps the `filter_keys` entries of a Dict observation and of its space.  Used by the
    model-free wrap chains [REF baselines/racing/experiments/acme/experiment.py:67,78]."""

    def __init__(self, env, filter_keys=None):
        super().__init__(env)
        keys = list(env.observation_space.spaces) if filter_keys is None else list(filter_keys)
        self._filter_keys = keys
        self.observation_space = Dict([(k, sp) for k, sp in env.observation_space.spaces.items() if k in keys])

    def observation(self, observation):
        return type(observation)([(k, v) for k, v in observation.items() if k in self._filter_keys])


class GymTimeLimit(Wrapper):
    """Restatement of gym.wrappers.TimeLimit (gym==0.18.0): counts env.step calls since reset; the step that reaches
    max_episode_steps returns done=True and info['TimeLimit.truncated'] = not done
    [REF baselines/racing/experiments/acme/experiment.py:71,83 TimeLimit(env, max_episode_steps=...)]."""

    def __init__(self, env, max_episode_steps):
        super().__init__(env)
        self._max_episode_steps = max_episode_steps
        self._elapsed_steps = None

    def step(self, action):
        assert self._elapsed_steps is not None, "Cannot call env.step() before calling reset()"
        observation, reward, done, info = self.env.step(action)
        self._elapsed_steps += 1
        if self._elapsed_steps >= self._max_episode_steps:
            info["TimeLimit.truncated"] = not done
            done = True
        return observation, reward, done, info

    def reset(self, **kwargs):
        self._elapsed_steps = 0
        return self.env.reset(**kwargs)


def reference_tools():
    """The reference's dreamer/tools.py, unmodified, with tensorflow / tensorflow_probability / tfplot / imageio stubbed
    (none of them is in this image): only the pure-Python driver loop `simulate` [REF dreamer/tools.py:154-206] is run."""
    install()

    class _Any:
        """Permissive stand-in: any attribute is another stand-in, calling it returns a stand-in -- or the argument itself
        when used as a decorator on a function."""
        def __init__(self, name="stub"):
            self.__name__ = name
        def __getattr__(self, k):
            if k.startswith("__"):
                raise AttributeError(k)
            return _Any(k)
        def __call__(self, *a, **kw):
            if len(a) == 1 and not kw and callable(a[0]) and not isinstance(a[0], _Any):
                return a[0]
            return _Any()
        def __enter__(self):
            return self
        def __exit__(self, *exc):
            return False

    def module(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        m.__getattr__ = lambda k: _Any(k)   # PEP 562: anything else
        sys.modules[name] = m
        return m

    class _Base:
        pass

    tfd = module("tensorflow_probability.distributions", MultivariateNormalDiag=type("MultivariateNormalDiag", (), {}),
                 Categorical=type("Categorical", (), {}))
    module("tensorflow_probability", distributions=tfd, bijectors=types.SimpleNamespace(Bijector=_Base))
    tf1 = module("tensorflow.compat.v1")
    module("tensorflow.compat", v1=tf1)
    prec = module("tensorflow.keras.mixed_precision.experimental")
    mp = module("tensorflow.keras.mixed_precision", experimental=prec)
    keras = module("tensorflow.keras", mixed_precision=mp)
    module("tensorflow", Module=_Base, compat=sys.modules["tensorflow.compat"], keras=keras)
    module("tfplot", autowrap=lambda *a, **kw: (lambda f: f))
    module("imageio")
    return _load("_ref_dreamer_tools", "dreamer/tools.py")


def reference_tasks():
    install()
    return _load("_ref_baselines_tasks", "baselines/racing/environment/tasks.py")
