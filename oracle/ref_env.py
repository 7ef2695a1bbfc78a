"""A racecar_gym-shaped single env over the CPU oracle (test infrastructure; see oracle/__init__.py).

``OracleRaceEnv`` plays the role of ``racecar_gym.envs.MultiAgentRaceEnv`` (one sim tick per ``step``,
dict actions ``{'A': {'motor', 'steering'}}``, info keys ``pose, velocity, progress, lap, time, wrong_way,
wall_collision`` [REF dreamer/wrappers.py:62-69,218-219,395]) so that the UNMODIFIED reference wrapper stack
[REF dreamer/dream.py:134-140] can run on top of it.  That reference-composed path is what the fused
oracle/GPU step is pinned against (tests/golden/make_golden.py).
"""
from __future__ import annotations

import types

import numpy as np

from racing_dreamer_b200 import _abi
from racing_dreamer_b200.maps import TrackMap

from . import ref_stubs
from .binding import Oracle, default_config


class GridMapShim:
    """``world._maps['occupancy']``: ``._map`` (bool, full image) and ``.to_pixel(pose)`` [REF dreamer/wrappers.py:376,396]."""

    def __init__(self, tm: TrackMap):
        self._tm = tm
        self._map = tm.full_drivable()

    @property
    def map(self):
        return self._map

    def to_pixel(self, pose):
        return self._tm.to_pixel(float(pose[0]), float(pose[1]))


AGENT_IDS = ("A", "B", "C", "D")


class OracleRaceEnv:
    """n_agents = 1: the single-car world of the dreamer scenarios; n_agents > 1: one world of cars 'A', 'B', ... that
    see and hit each other [REF baselines/scenarios/max_progress/austria.yml:3-34], `tasks` = task name per agent."""

    def __init__(self, tm: TrackMap, laps=10, time_limit=180.0, terminate_on_collision=True, collision_reward=-1.0,
                 reset_mode="grid", seed=0, n_agents=1, tasks=None, n_step_progress=10, ball_spacing=1.5):
        ref_stubs.install()
        import gym
        cfg = default_config()
        cfg.n_envs = n_agents
        cfg.action_repeat = 1          # one sim tick per step; the reference's ActionRepeat wrapper loops
        cfg.rescale_actions = 0        # the reference's ReduceActionSpace wrapper rescales
        cfg.auto_reset = 0
        cfg.laps = laps
        cfg.time_limit = time_limit
        cfg.terminate_on_collision = int(terminate_on_collision)
        cfg.collision_reward = collision_reward
        cfg.seed = seed
        cfg.agents_per_world = n_agents
        tasks = list(tasks) if tasks is not None else ["maximize_progress"] * n_agents
        for a in range(n_agents):
            cfg.agent_task[a] = _abi.TASKS[tasks[a]]
        if n_agents == 1:
            cfg.task = _abi.TASKS[tasks[0]]
        cfg.n_step_progress = n_step_progress
        cfg.ball_spacing = ball_spacing
        self.cfg = cfg
        self.tm = tm
        self.ids = list(AGENT_IDS[:n_agents])
        self._orc = Oracle(cfg, [tm])
        self._mode = _abi.RESET_MODES[reset_mode]
        self.scenario = types.SimpleNamespace(world=types.SimpleNamespace(
            _maps={"occupancy": GridMapShim(tm)}, _config=types.SimpleNamespace(name=tm.name)))
        box = gym.spaces.Box
        self.observation_space = gym.spaces.Dict({i: gym.spaces.Dict({
            "lidar": box(0.0, 15.0, shape=(1080,), dtype=np.float64),
            "pose": box(-100.0, 100.0, shape=(6,), dtype=np.float64),
            "velocity": box(-10.0, 10.0, shape=(6,), dtype=np.float64)}) for i in self.ids})
        self.action_space = gym.spaces.Dict({i: gym.spaces.Dict({
            "motor": box(-1.0, 1.0, shape=(1,), dtype=np.float64),
            "steering": box(-1.0, 1.0, shape=(1,), dtype=np.float64)}) for i in self.ids})

    # -- racecar_gym API --
    def _obs(self, out):
        return {i: {"lidar": out["lidar"][k].astype(np.float64), "pose": self._pose(k), "velocity": self._velocity(k)}
                for k, i in enumerate(self.ids)}

    def _pose(self, k=0):
        f = self._orc.f64
        yaw = f[_abi.S_YAW, k]
        two_pi = 6.283185307179586
        return np.array([f[_abi.S_X, k], f[_abi.S_Y, k], 0.0, 0.0, 0.0, yaw - np.rint(yaw / two_pi) * two_pi])

    def _velocity(self, k=0):
        f = self._orc.f64
        v, b = f[_abi.S_V, k], f[_abi.S_SLIP, k]
        return np.array([v * np.cos(b), v * np.sin(b), 0.0, 0.0, 0.0, f[_abi.S_YAWRATE, k]])

    def _info(self, out):
        info = {}
        for k, i in enumerate(self.ids):
            fl = int(out["flags"][k])
            info[i] = {"pose": self._pose(k), "velocity": self._velocity(k),
                       "progress": float(self._orc.f64[_abi.S_PROGRESS, k]), "lap": int(out["lap"][k]),
                       "time": float(self._orc.f64[_abi.S_TIME, k]), "wrong_way": bool(fl & _abi.F_WRONG_WAY),
                       "wall_collision": bool(fl & _abi.F_COLLISION),
                       "opponent_collisions": [self.ids[j] for j in range(len(self.ids)) if int(out["opponents"][k]) >> j & 1],
                       "rank": int(out["rank"][k])}
        return info

    def reset(self, mode="grid"):
        out = self._orc.reset(mode=_abi.RESET_MODES[mode])
        return self._obs(out)

    def step(self, action):
        cmd = np.array([[float(np.asarray(action[i]["motor"]).reshape(-1)[0]),
                         float(np.asarray(action[i]["steering"]).reshape(-1)[0])] for i in self.ids], dtype=np.float64)
        self._orc.i32[_abi.I_FLAGS] &= ~_abi.F_NEEDS_RESET  # racecar_gym keeps stepping after done
        out = self._orc.step(commands=cmd)
        return (self._obs(out), {i: float(out["reward64"][k]) for k, i in enumerate(self.ids)},
                {i: bool(out["done"][k]) for k, i in enumerate(self.ids)}, self._info(out))

    def render(self, **kwargs):
        return np.zeros((8, 8, 3), np.uint8)

    def close(self):
        pass


def make_reference_stack(tm: TrackMap, action_repeat=4, time_limit_steps=500, reset_mode="grid", **env_kw):
    """The reference's ``make_base_env`` + train/test wrappers [REF dreamer/dream.py:103-140], built from the
    unmodified reference classes over ``OracleRaceEnv``."""
    W = ref_stubs.reference_wrappers()
    env = OracleRaceEnv(tm, **env_kw)
    env = W.RaceCarWrapper(env, agent_id="A")
    env = W.ActionRepeat(env, action_repeat)
    env = W.ReduceActionSpace(env, low=[0.005, -1.0], high=[1.0, 1.0])
    env = W.OccupancyMapObs(env)
    env = W.FixedResetMode(env, reset_mode)
    env = W.TimeLimit(env, time_limit_steps)
    env = W.Collect(env, callbacks=[], precision=32)
    return env
