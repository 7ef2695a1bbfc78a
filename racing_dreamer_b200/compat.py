"""Gym-style dict API over the CUDA env step -- the drop-in for the reference's env path.

Two views of the same kernels:

* ``ReferenceEnv`` / ``make_reference_env``: the OUTERMOST interface of the reference's wrapper stack
  ``RaceCarBaseEnv -> RaceCarWrapper -> ActionRepeat -> ReduceActionSpace -> OccupancyMapObs -> FixedResetMode
  -> TimeLimit`` (+ the float32/int32 casts of ``Collect._convert``) [REF dreamer/dream.py:103-140;
  dreamer/wrappers.py:10-158,210-250,372-414], fused into ONE ``rd_step`` per agent step.
  ``dream.py::make_train_env/make_test_env`` can return this object instead of the wrapped PyBullet env.
* ``RaceCarGymCompat``: the ``racecar_gym.MultiAgentRaceEnv`` boundary itself (one 10 ms sim tick per ``step``,
  ``{'A': {'motor','steering'}}`` actions, ``reset(mode=...)``) [REF dreamer/wrappers.py:14-15,62-77,92], for
  callers that want to stack the reference's own unmodified wrapper classes on top.

Both are single-agent (agent id ``'A'``) views of a ``BatchedRaceEnv`` with ``n_envs = 1``; results come from the
same kernels as the batched path.  Training-scale callers should use ``BatchedRaceEnv`` directly.
"""
from __future__ import annotations

import types
from pathlib import Path
from typing import Dict as TDict, Optional, Union

import numpy as np
import torch

from . import _abi, spaces
from .env import BatchedRaceEnv, EnvConfig
from .maps import TrackMap, load_track

# task parameters of the reference's scenario files [REF dreamer/scenarios/max_progress/*.yml: laps 10;
# dreamer/scenarios/eval/*.yml: laps 1; both time_limit 180 s, terminate_on_collision, collision_reward -1]
SCENARIO_DEFAULTS = {
    "max_progress": dict(task="maximize_progress", laps=10, time_limit=180.0, terminate_on_collision=True,
                         collision_reward=-1.0),
    "eval": dict(task="maximize_progress", laps=1, time_limit=180.0, terminate_on_collision=True,
                 collision_reward=-1.0),
    "max_speed": dict(task="max_speed", laps=10, time_limit=180.0, terminate_on_collision=False,
                      collision_reward=-1.0),
}


def load_scenario(path: Union[str, Path], agent_id: str = "A") -> dict:
    """Parse a reference scenario YAML [REF dreamer/scenarios/max_progress/austria.yml:1-10] into
    ``{'track': world.name, 'task': ..., 'laps': ..., ...}`` (the keys ``EnvConfig`` understands)."""
    import yaml
    spec = yaml.safe_load(Path(path).read_text())
    out = {"track": spec["world"]["name"]}
    for agent in spec.get("agents", []):
        if agent.get("id") != agent_id:
            continue
        task = agent.get("task", {})
        out["task"] = task.get("task_name", "maximize_progress")
        for k, v in (task.get("params") or {}).items():
            if k in ("laps", "time_limit", "terminate_on_collision", "collision_reward", "progress_reward",
                     "frame_reward"):
                out[k] = v
        out["sensors"] = list(agent.get("vehicle", {}).get("sensors", []))
    return out


class _GridMap:
    """``scenario.world._maps[...]``: ``._map`` + ``.to_pixel(pose)`` [REF dreamer/wrappers.py:376,396-399]."""

    def __init__(self, tm: TrackMap, array: np.ndarray):
        self._tm = tm
        self._map = array

    @property
    def map(self):
        return self._map

    def to_pixel(self, pose):
        return self._tm.to_pixel(float(pose[0]), float(pose[1]))


def _scenario_of(tm: TrackMap):
    """Lazy stand-in for ``env.scenario`` [REF dreamer/wrappers.py:30-32; dreamer/evaluations/run_evaluation.py:48-49]."""
    class _Maps(dict):
        def __missing__(self, key):
            arr = {"occupancy": tm.full_drivable, "progress": tm.full_norm_distance_from_start,
                   "obstacle": tm.full_norm_distance_to_obstacle}[key]()
            self[key] = _GridMap(tm, arr)
            return self[key]
    world = types.SimpleNamespace(_maps=_Maps(), _config=types.SimpleNamespace(name=tm.name))
    return types.SimpleNamespace(world=world)


class _SingleAgentBase:
    agent_id = "A"

    def _make(self, track, n_envs=1, **cfg_kw):
        self._tm = track if isinstance(track, TrackMap) else load_track(track)
        self._env = BatchedRaceEnv(EnvConfig(tracks=(self._tm,), n_envs=n_envs, auto_reset=False, **cfg_kw),
                                   device=self._device)
        self.scenario = _scenario_of(self._tm)

    @property
    def agent_ids(self):
        return [self.agent_id]

    @property
    def n_agents(self):
        return 1

    def _host(self) -> TDict[str, np.ndarray]:
        """One device->host read of every (tiny) result buffer of env 0."""
        buf = self._env.buf
        torch.cuda.current_stream(self._env.device).synchronize()
        return {k: v[0].cpu().numpy() for k, v in buf.items() if v is not None}

    def _info(self, h, f64=None) -> dict:
        fl = int(h["flags"])
        pose = h["pose"].astype(np.float64)
        vel = h["velocity"].astype(np.float64)
        if f64 is not None:  # float64 pose straight from the state (racecar_gym reports float64)
            yaw = f64[_abi.S_YAW]
            pose = np.array([f64[_abi.S_X], f64[_abi.S_Y], 0.0, 0.0, 0.0, yaw - np.rint(yaw / (2 * np.pi)) * 2 * np.pi])
            v, b = f64[_abi.S_V], f64[_abi.S_SLIP]
            vel = np.array([v * np.cos(b), v * np.sin(b), 0.0, 0.0, 0.0, f64[_abi.S_YAWRATE]])
        return {"pose": pose, "velocity": vel, "progress": float(h["progress"]), "lap": int(h["lap"]),
                "time": float(h["time"]), "wrong_way": bool(fl & _abi.F_WRONG_WAY),
                "wall_collision": bool(fl & _abi.F_COLLISION), "opponent_collisions": [],
                "left_map": bool(fl & _abi.F_LEFT_MAP)}

    def render(self, mode: str = "birds_eye", agent: str = "A", **kwargs) -> np.ndarray:
        """Top-down RGB view of the drivable area around the car (the reference renders through PyBullet
        [REF dreamer/wrappers.py:178-195]; videos are not on the hot path, so this is a plain map crop)."""
        h = self._host()
        occ = self.scenario.world._maps["occupancy"]
        pr, pc = occ.to_pixel(h["pose"])
        half = 100
        m = np.pad(occ._map, half, mode="constant")
        crop = m[pr:pr + 2 * half, pc:pc + 2 * half]
        img = np.repeat((crop.astype(np.uint8) * 255)[..., None], 3, axis=2)
        img[half - 2:half + 3, half - 2:half + 3] = (255, 0, 0)
        return img

    def close(self):
        self._env.close()

    @property
    def launch_count(self) -> int:
        return self._env.launch_count


class ReferenceEnv(_SingleAgentBase):
    """Fused equivalent of the reference's fully wrapped env (see module docstring).

    ``reset() -> {'A': obs}``; ``step({'A': a}) -> (obs, rewards, dones, infos)`` with ``a`` a length-2 array
    ``[motor, steering]`` in [-1, 1] [REF dreamer/wrappers.py:55-63,129-134].  obs keys: ``lidar`` f32[1080],
    ``pose`` f32[6], ``velocity`` f32[6], ``speed`` f32 scalar, ``lidar_occupancy`` u8[64,64,1] (zeros on reset
    [REF dreamer/wrappers.py:410-414]) -- dtypes as ``Collect._convert`` leaves them at precision 32.
    """

    def __init__(self, track="austria", task="max_progress", action_repeat=4, time_limit_steps=None,
                 reset_mode="random", occupancy=True, device=None, scenario: Optional[str] = None, **overrides):
        self._device = device
        params = dict(SCENARIO_DEFAULTS.get(task, SCENARIO_DEFAULTS["max_progress"]))
        if scenario is not None:
            sc = load_scenario(scenario)
            track = sc.pop("track", track)
            sc.pop("sensors", None)
            params.update(sc)
        params.update(overrides)
        if time_limit_steps is None:  # dream.py: time_limit_train 2000 sim ticks / action_repeat [REF dreamer/dream.py:57,109]
            time_limit_steps = 2000 // int(action_repeat)
        self._make(track, action_repeat=int(action_repeat), reset_mode=reset_mode,
                   obs_type="lidar_occupancy" if occupancy else "lidar", time_limit_steps=int(time_limit_steps),
                   rescale_actions=True, **params)
        self._occupancy = occupancy
        self._needs_reset = True
        self._action = torch.zeros((1, 2), dtype=torch.float32, device=self._env.device)

    @property
    def observation_space(self):
        box = spaces.Box
        sp = {"lidar": box(0.0, 15.0, shape=(1080,), dtype=np.float32),
              "pose": box(-100.0, 100.0, shape=(6,), dtype=np.float32),
              "velocity": box(-10.0, 10.0, shape=(6,), dtype=np.float32),
              "speed": box(-np.inf, np.inf, shape=(1,), dtype=np.float32)}  # [REF dreamer/wrappers.py:50]
        if self._occupancy:
            sp["lidar_occupancy"] = box(0, 1, shape=(64, 64, 1), dtype=np.uint8)  # [REF dreamer/wrappers.py:380-385]
        return spaces.Dict({self.agent_id: spaces.Dict(sp)})

    @property
    def action_space(self):  # [REF dreamer/wrappers.py:55-60]: Box(append(motor.low, steering.low), ...)
        return spaces.Dict({self.agent_id: spaces.Box(np.array([-1.0, -1.0], np.float32), np.array([1.0, 1.0], np.float32))})

    def _obs(self, h, reset: bool):
        obs = {"lidar": h["lidar"], "pose": h["pose"], "velocity": h["velocity"],
               "speed": np.float32(0.0) if reset else np.float32(h["speed"])}
        if self._occupancy:
            obs["lidar_occupancy"] = h["occupancy"]
        return {self.agent_id: obs}

    def reset(self, mode: Optional[str] = None):
        self._env.reset(mode=mode)
        self._needs_reset = False
        return self._obs(self._host(), reset=True)

    def step(self, actions):
        assert not self._needs_reset, "Must reset environment."  # [REF dreamer/wrappers.py:148]
        a = np.asarray(actions[self.agent_id], dtype=np.float32).reshape(1, 2)
        self._action.copy_(torch.from_numpy(a))
        self._env.step(self._action)
        h = self._host()
        done = bool(h["done"])
        self._needs_reset = done
        aid = self.agent_id
        return self._obs(h, reset=False), {aid: float(h["reward"])}, {aid: done}, {aid: self._info(h)}


class RaceCarGymCompat(_SingleAgentBase):
    """``racecar_gym.envs.MultiAgentRaceEnv``-shaped view: one sim tick per ``step`` (see module docstring)."""

    def __init__(self, track="austria", task="max_progress", device=None, scenario: Optional[str] = None, **overrides):
        self._device = device
        params = dict(SCENARIO_DEFAULTS.get(task, SCENARIO_DEFAULTS["max_progress"]))
        if scenario is not None:
            sc = load_scenario(scenario)
            track = sc.pop("track", track)
            sc.pop("sensors", None)
            params.update(sc)
        params.update(overrides)
        self._make(track, action_repeat=1, rescale_actions=False, obs_type="lidar", time_limit_steps=0, **params)
        self._action = torch.zeros((1, 2), dtype=torch.float32, device=self._env.device)

    @property
    def observation_space(self):
        box = spaces.Box
        return spaces.Dict({self.agent_id: spaces.Dict({
            "lidar": box(0.0, 15.0, shape=(1080,), dtype=np.float64),
            "pose": box(-100.0, 100.0, shape=(6,), dtype=np.float64),
            "velocity": box(-10.0, 10.0, shape=(6,), dtype=np.float64)})})

    @property
    def action_space(self):
        box = spaces.Box
        return spaces.Dict({self.agent_id: spaces.Dict({
            "motor": box(-1.0, 1.0, shape=(1,), dtype=np.float64),
            "steering": box(-1.0, 1.0, shape=(1,), dtype=np.float64)})})

    def _state(self):
        f, _ = self._env.get_state()
        return f[:, 0].cpu().numpy()

    def _obs(self, h, info):
        return {self.agent_id: {"lidar": h["lidar"].astype(np.float64), "pose": info["pose"], "velocity": info["velocity"]}}

    def reset(self, mode: str = "grid"):
        self._env.reset(mode=mode)
        h = self._host()
        return self._obs(h, self._info(h, self._state()))

    def step(self, actions):
        a = actions[self.agent_id]
        cmd = np.array([[float(np.asarray(a["motor"]).reshape(-1)[0]), float(np.asarray(a["steering"]).reshape(-1)[0])]],
                       dtype=np.float32)
        # racecar_gym keeps stepping after a terminal tick; the wrappers above decide when to reset
        f, i = self._env.get_state()
        i[_abi.I_FLAGS] &= ~_abi.F_NEEDS_RESET
        self._env.set_state(f, i)
        self._action.copy_(torch.from_numpy(cmd))
        self._env.step(self._action)
        h = self._host()
        info = self._info(h, self._state())
        aid = self.agent_id
        return self._obs(h, info), {aid: float(h["reward"])}, {aid: bool(h["done"])}, {aid: info}


def make_reference_env(track: str, task: str = "max_progress", action_repeat: int = 4, mode: str = "train",
                       device=None, **kw) -> ReferenceEnv:
    """``make_train_env`` / ``make_test_env`` of dream.py without the PyBullet sim [REF dreamer/dream.py:103-131]:
    train = reset mode 'random', TimeLimit 2000/action_repeat; test = 'grid', 4000/action_repeat."""
    if mode == "train":
        return ReferenceEnv(track, task, action_repeat, time_limit_steps=2000 // action_repeat, reset_mode="random",
                            device=device, **kw)
    return ReferenceEnv(track, task, action_repeat, time_limit_steps=4000 // action_repeat, reset_mode="grid",
                        device=device, **kw)
