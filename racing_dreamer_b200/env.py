"""BatchedRaceEnv -- the batched tensor API over librd_env.so.

PyTorch is plumbing here (device memory, streams); every number is produced by the CUDA kernels behind
the C ABI in ``include/rd_env.h``.  There is no CPU path: constructing an env without the native library
or without a CUDA device raises.

Reference interfaces mirrored (argument meaning and defaults):
* ``RaceCarBaseEnv(track, task)`` + scenario YAML params [REF dreamer/wrappers.py:10-16;
  dreamer/scenarios/max_progress/austria.yml:8-10]
* ``ActionRepeat(amount)``, ``ReduceActionSpace(low, high)``, ``TimeLimit(duration)``, ``FixedResetMode(mode)``,
  ``OccupancyMapObs`` [REF dreamer/wrappers.py:86-158,372-414; dreamer/dream.py:134-140]
"""
from __future__ import annotations

import ctypes as C
import dataclasses
from typing import Dict, Optional, Sequence, Union

import numpy as np
import torch

from . import _abi
from .maps import TrackMap, load_track


@dataclasses.dataclass
class EnvConfig:
    """Host-side mirror of ``rd_config``; defaults = the reference's dreamer training setup."""
    tracks: Sequence[Union[str, TrackMap]] = ("austria",)
    n_envs: int = 1
    action_repeat: int = 4                      # [REF dreamer/dream.py:55]
    obs_type: str = "lidar"                     # 'lidar' | 'lidar_occupancy' [REF dreamer/dream.py:64-66]
    normalize_lidar: bool = False               # store r/15-0.5 [REF dreamer/tools.py:274]
    normalize_obs: Optional[str] = None         # 'baselines': lidar / pose / velocity as (x-low)/(high-low), the model-free
                                                # chain's NormalizeObservations [REF baselines/racing/environment/single_agent.py:66-99]
    obs_low: Sequence[float] = (0.0, -100.0, -10.0)    # Box bounds of lidar / pose / velocity (normalize_obs)
    obs_high: Sequence[float] = (15.0, 100.0, 10.0)
    lidar_dtype: str = "float32"                # 'float16' = what Collect hands on at precision 16 [REF dreamer/wrappers.py:240-250]
    task: str = "maximize_progress"
    laps: int = 10
    time_limit: float = 180.0
    terminate_on_collision: bool = True
    collision_reward: float = -1.0
    progress_reward: float = 100.0
    frame_reward: float = 0.0
    n_checkpoints: int = 20
    time_limit_steps: int = 0                   # TimeLimit(duration) in agent steps; 0 = off
    time_limit_ticks: int = 0                   # gym TimeLimit inside ActionRepeat (baselines chain), in sim ticks; 0 = off
                                                # [REF baselines/racing/experiments/acme/experiment.py:66-72]
    auto_reset: bool = True
    reset_mode: str = "grid"
    rescale_actions: bool = True                # ReduceActionSpace
    action_low: Sequence[float] = (0.005, -1.0)  # [REF dreamer/dream.py:138]
    action_high: Sequence[float] = (1.0, 1.0)
    clip_actions: bool = False
    repeat_semantics: str = "dreamer"           # 'dreamer' | 'baselines'
    progress_abs: bool = False
    n_beams: int = 1080
    lidar_noise: float = 0.0
    seed: int = 0
    env_id_offset: int = 0
    map_ids: Optional[Sequence[int]] = None     # per-env index into `tracks`; default: env i -> track i % len
    # multi-agent worlds [REF baselines/scenarios/max_progress/austria.yml:3-34; dreamer/dream.py:105-106]:
    # env e = agent (e % agents_per_world) of world (e // agents_per_world); n_envs counts CARS
    agents_per_world: int = 1
    agent_tasks: Optional[Sequence[str]] = None  # task of agent A, B, ...; default: `task` for every agent
    n_step_progress: int = 10                   # [REF baselines/scenarios/max_progress/austria.yml:18]
    ball_spacing: float = 1.5                   # metres of track between the cars of a world after a random reset


def _fill_config(cfg: _abi.RdConfig, ec: EnvConfig) -> None:
    cfg.n_envs = int(ec.n_envs)
    cfg.n_beams = int(ec.n_beams)
    cfg.action_repeat = int(ec.action_repeat)
    cfg.repeat_semantics = {"dreamer": _abi.REPEAT_DREAMER, "baselines": _abi.REPEAT_BASELINES}[ec.repeat_semantics]
    flags = _abi.OBS_LIDAR
    if ec.obs_type == "lidar_occupancy":
        flags |= _abi.OBS_OCCUPANCY
    elif ec.obs_type != "lidar":
        raise ValueError(f"obs_type {ec.obs_type!r}: expected 'lidar' or 'lidar_occupancy'")
    if ec.normalize_lidar:
        flags |= _abi.OBS_LIDAR_NORM
    if ec.normalize_obs == "baselines":
        if ec.normalize_lidar:
            raise ValueError("normalize_obs='baselines' excludes normalize_lidar (dreamer's r/15 - 0.5)")
        flags |= _abi.OBS_NORM_BASELINES
        for k in range(3):
            cfg.obs_low[k] = float(ec.obs_low[k])
            cfg.obs_high[k] = float(ec.obs_high[k])
    elif ec.normalize_obs is not None:
        raise ValueError(f"normalize_obs {ec.normalize_obs!r}: expected None or 'baselines'")
    if ec.lidar_dtype == "float16":
        flags |= _abi.OBS_LIDAR_F16
    elif ec.lidar_dtype != "float32":
        raise ValueError(f"lidar_dtype {ec.lidar_dtype!r}: expected 'float32' or 'float16'")
    cfg.obs_flags = flags
    cfg.task = _abi.TASKS[ec.task]
    cfg.laps = int(ec.laps)
    cfg.terminate_on_collision = int(ec.terminate_on_collision)
    cfg.n_checkpoints = int(ec.n_checkpoints)
    cfg.time_limit_steps = int(ec.time_limit_steps)
    cfg.time_limit_ticks = int(ec.time_limit_ticks)
    cfg.auto_reset = int(ec.auto_reset)
    cfg.reset_mode = _abi.RESET_MODES[ec.reset_mode]
    cfg.rescale_actions = int(ec.rescale_actions)
    cfg.clip_actions = int(ec.clip_actions)
    cfg.progress_abs = int(ec.progress_abs)
    cfg.env_id_offset = int(ec.env_id_offset)
    cfg.seed = int(ec.seed) & 0xFFFFFFFFFFFFFFFF
    cfg.time_limit = float(ec.time_limit)
    cfg.collision_reward = float(ec.collision_reward)
    cfg.progress_reward = float(ec.progress_reward)
    cfg.frame_reward = float(ec.frame_reward)
    for k in range(2):
        cfg.action_low[k] = float(ec.action_low[k])
        cfg.action_high[k] = float(ec.action_high[k])
    cfg.lidar_noise = float(ec.lidar_noise)
    A = int(ec.agents_per_world)
    if not 1 <= A <= _abi.MAX_AGENTS:
        raise ValueError(f"agents_per_world {A}: expected 1..{_abi.MAX_AGENTS}")
    tasks = list(ec.agent_tasks) if ec.agent_tasks is not None else [ec.task] * A
    if len(tasks) != A:
        raise ValueError("agent_tasks must name one task per agent of a world")
    cfg.agents_per_world = A
    for a in range(_abi.MAX_AGENTS):
        cfg.agent_task[a] = _abi.TASKS[tasks[a]] if a < A else _abi.TASK_MAX_PROGRESS
    if A == 1:
        cfg.task = _abi.TASKS[tasks[0]]
    cfg.n_step_progress = int(ec.n_step_progress)
    cfg.ball_spacing = float(ec.ball_spacing)


class BatchedRaceEnv:
    """N independent single-car racing envs stepped by one GPU.

    ``reset(mask=None, mode=None) -> obs`` and ``step(actions f32[N,2]) -> (obs, reward, done, info)``, all
    values CUDA tensors that alias persistent output buffers (valid until the next call).
    """

    def __init__(self, config: Optional[EnvConfig] = None, device: Union[str, torch.device, None] = None,
                 raw_config: Optional[_abi.RdConfig] = None, out_buffers: Optional[Dict[str, torch.Tensor]] = None,
                 **kwargs):
        """out_buffers: optional caller-owned result tensors (keys of OUT_SPEC) the kernels write into -- e.g. slices
        of one larger allocation (HostSteppedEnv); anything not given is allocated here."""
        self.lib = _abi.load_library()
        self._out_buffers = out_buffers
        if not torch.cuda.is_available():
            raise _abi.NativeLibraryError("BatchedRaceEnv needs a CUDA device: the env step has no CPU fallback")
        ec = config if config is not None else EnvConfig(**kwargs)
        self.config = ec
        self.device = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
        self.tracks = [t if isinstance(t, TrackMap) else load_track(t) for t in ec.tracks]
        if raw_config is not None:
            cfg = raw_config.copy()
        else:
            cfg = _abi.default_config()
            _fill_config(cfg, ec)
        self.cfg = cfg
        self.n = int(cfg.n_envs)
        self.n_beams = int(cfg.n_beams)
        self._handle = C.c_void_p()
        with torch.cuda.device(self.device):
            self._check(self.lib.rd_create(C.byref(cfg), C.byref(self._handle)), None)
            self._keep = []
            for mid, tm in enumerate(self.tracks):
                bits = np.ascontiguousarray(tm.packed_bits_yup())
                dist = np.ascontiguousarray(tm.dist_yup())
                start = np.ascontiguousarray(tm.start_poses, dtype=np.float64)
                rst = np.ascontiguousarray(tm.reset_poses, dtype=np.float64)
                self._check(self.lib.rd_upload_map(
                    self._handle, mid, bits.ctypes.data, tm.h, tm.w, tm.row_words(), dist.ctypes.data, tm.dmax,
                    tm.resolution, tm.origin[0], tm.origin[1], tm.c0, tm.cy0, tm.full_shape[0],
                    start.ctypes.data, start.shape[0], rst.ctypes.data, rst.shape[0]))
            if ec.map_ids is not None:
                ids = np.ascontiguousarray(ec.map_ids, dtype=np.int32)
            else:  # worlds (not cars) alternate over the tracks
                A = max(1, int(cfg.agents_per_world))
                ids = ((np.arange(self.n, dtype=np.int32) // A) % len(self.tracks)).astype(np.int32)
            if ids.shape != (self.n,):
                raise ValueError("map_ids must have one entry per env")
            self.map_ids = ids
            self._check(self.lib.rd_assign_maps(self._handle, ids.ctypes.data))
            self._alloc()

    def assign_maps(self, map_ids: Sequence[int]) -> None:
        """Re-assign the tracks (indices into ``tracks``) of the envs -- every track was uploaded at construction, so
        switching is one small host->device copy (racecar_gym's ChangingTrack* envs [REF dreamer/evaluations/make_env.py:
        6-11; baselines/racing/experiments/acme/experiment.py:90-93]).  The envs must be reset afterwards."""
        ids = np.ascontiguousarray(map_ids, dtype=np.int32)
        if ids.shape != (self.n,):
            raise ValueError("map_ids must have one entry per env")
        with torch.cuda.device(self.device):
            torch.cuda.current_stream(self.device).synchronize()
            self._check(self.lib.rd_assign_maps(self._handle, ids.ctypes.data))
        self.map_ids = ids

    # ------------------------------------------------------------------ buffers
    OUT_SPEC = (  # key, per-env shape, dtype  [REF dreamer/wrappers.py:62-69,210-226: obs + what Collect records]
        ("lidar", None, torch.float32), ("occupancy", (64, 64, 1), torch.uint8), ("pose", (6,), torch.float32),
        ("velocity", (6,), torch.float32), ("speed", (), torch.float32), ("reward", (), torch.float32),
        ("done", (), torch.uint8), ("progress", (), torch.float32), ("lap", (), torch.int32), ("time", (), torch.float32),
        ("flags", (), torch.uint8), ("rank", (), torch.int32), ("opponents", (), torch.uint8),
        ("wrong_way", (), torch.bool), ("wall_collision", (), torch.bool), ("done_bool", (), torch.bool))

    def _alloc(self):
        n, dev = self.n, self.device
        occ = bool(self.cfg.obs_flags & _abi.OBS_OCCUPANCY)
        given = self._out_buffers or {}
        self.buf: Dict[str, Optional[torch.Tensor]] = {}
        for key, shape, dtype in self.OUT_SPEC:
            if key == "occupancy" and not occ:
                self.buf[key] = None
                continue
            full = (n,) + ((self.n_beams,) if key == "lidar" else tuple(shape))
            if key == "lidar" and (self.cfg.obs_flags & _abi.OBS_LIDAR_F16):
                dtype = torch.float16
            if key == "done_bool":      # a bool VIEW of the kernel's 0/1 bytes: step() returns it without launching anything
                self.buf[key] = self.buf["done"].view(torch.bool)
                continue
            t = given.get(key)
            if t is None:
                t = torch.zeros(full, dtype=dtype, device=dev)
            elif tuple(t.shape) != full or t.dtype != dtype or t.device != dev or not t.is_contiguous():
                raise ValueError(f"out_buffers[{key!r}] must be a contiguous {dtype} tensor of shape {full} on {dev}")
            self.buf[key] = t
        o = _abi.RdOutputs()
        for k, f in (("lidar", "lidar_dev"), ("occupancy", "occupancy_dev"), ("pose", "pose_dev"),
                     ("velocity", "velocity_dev"), ("speed", "speed_dev"), ("reward", "reward_dev"),
                     ("done", "done_dev"), ("progress", "progress_dev"), ("lap", "lap_dev"), ("time", "time_dev"),
                     ("flags", "flags_dev"), ("rank", "rank_dev"), ("opponents", "opponents_dev"),
                     ("wrong_way", "wrong_way_dev"), ("wall_collision", "wall_collision_dev")):
            t = self.buf[k]
            setattr(o, f, t.data_ptr() if t is not None else None)
        self._out = o
        self._actions = torch.zeros((n, 2), dtype=torch.float32, device=dev)

    def _check(self, rc: int, handle="self"):
        if rc != 0:
            h = self._handle if handle == "self" else None
            msg = self.lib.rd_last_error(h)
            raise RuntimeError(f"librd_env error {rc}: {msg.decode() if msg else '?'}")

    def _stream(self) -> int:
        return torch.cuda.current_stream(self.device).cuda_stream

    # ------------------------------------------------------------------ env API
    def _obs(self) -> Dict[str, torch.Tensor]:
        obs = {"lidar": self.buf["lidar"], "pose": self.buf["pose"], "velocity": self.buf["velocity"],
               "speed": self.buf["speed"]}
        if self.buf["occupancy"] is not None:
            obs["lidar_occupancy"] = self.buf["occupancy"]
        return obs

    def _info(self) -> Dict[str, torch.Tensor]:
        # every value is a persistent buffer the kernels wrote (the two booleans included): no torch kernel runs here
        return {"progress": self.buf["progress"], "lap": self.buf["lap"], "time": self.buf["time"],
                "wrong_way": self.buf["wrong_way"], "wall_collision": self.buf["wall_collision"],
                "flags": self.buf["flags"], "pose": self.buf["pose"], "velocity": self.buf["velocity"],
                # multi-agent worlds: position in the world (1 = leader) and a bit mask over the world's agent
                # indices the car is in contact with (racecar_gym: info['rank'], info['opponent_collisions'])
                "rank": self.buf["rank"], "opponent_collisions": self.buf["opponents"]}

    def reset(self, mask: Optional[torch.Tensor] = None, mode: Optional[str] = None) -> Dict[str, torch.Tensor]:
        m = _abi.RESET_MODES[mode] if mode is not None else int(self.cfg.reset_mode)
        mp = None
        if mask is not None:
            mask = mask.to(device=self.device, dtype=torch.uint8).contiguous()
            if mask.shape != (self.n,):
                raise ValueError("mask must have shape (n_envs,)")
            mp = mask.data_ptr()
        with torch.cuda.device(self.device):
            self._check(self.lib.rd_reset(self._handle, mp, m, C.byref(self._out), self._stream()))
        return self._obs()

    def step(self, actions: torch.Tensor):
        if actions.shape != (self.n, 2):
            raise ValueError(f"actions must have shape ({self.n}, 2), got {tuple(actions.shape)}")
        a = actions
        if a.dtype != torch.float32 or a.device != self.device or not a.is_contiguous():
            self._actions.copy_(a, non_blocking=True)
            a = self._actions
        with torch.cuda.device(self.device):
            self._check(self.lib.rd_step(self._handle, a.data_ptr(), C.byref(self._out), self._stream()))
        return self._obs(), self.buf["reward"], self.buf["done_bool"], self._info()

    def step_raw(self, actions_ptr: int) -> None:
        """rd_step on a raw device pointer; results land in ``self.buf`` (bench / tight loops)."""
        self._check(self.lib.rd_step(self._handle, actions_ptr, C.byref(self._out), self._stream()))

    # ------------------------------------------------------------------ stage entry points (parity tests)
    def lidar_cast(self, poses: torch.Tensor, map_ids: Optional[np.ndarray] = None) -> torch.Tensor:
        p = poses.to(device=self.device, dtype=torch.float64).contiguous().reshape(-1, 3)
        dt = torch.float16 if (self.cfg.obs_flags & _abi.OBS_LIDAR_F16) else torch.float32
        out = torch.empty((p.shape[0], self.n_beams), dtype=dt, device=self.device)
        ids = None if map_ids is None else np.ascontiguousarray(map_ids, dtype=np.int32)
        with torch.cuda.device(self.device):
            self._check(self.lib.rd_lidar_cast(self._handle, p.data_ptr(), ids.ctypes.data if ids is not None else None,
                                               p.shape[0], out.data_ptr(), self._stream()))
        return out

    def occupancy_obs(self, poses: torch.Tensor, map_ids: Optional[np.ndarray] = None) -> torch.Tensor:
        p = poses.to(device=self.device, dtype=torch.float64).contiguous().reshape(-1, 3)
        out = torch.empty((p.shape[0], 64, 64), dtype=torch.uint8, device=self.device)
        ids = None if map_ids is None else np.ascontiguousarray(map_ids, dtype=np.int32)
        with torch.cuda.device(self.device):
            self._check(self.lib.rd_occupancy_obs(self._handle, p.data_ptr(),
                                                  ids.ctypes.data if ids is not None else None, p.shape[0],
                                                  out.data_ptr(), self._stream()))
        return out

    def dynamics(self, state: torch.Tensor, commands: torch.Tensor, n_ticks: int) -> torch.Tensor:
        s = state.to(device=self.device, dtype=torch.float64).contiguous().clone()
        if s.shape[0] != 7:
            raise ValueError("state must be [7, n]")
        cmd = commands.to(device=self.device, dtype=torch.float64).contiguous().reshape(-1, 2)
        with torch.cuda.device(self.device):
            self._check(self.lib.rd_dynamics(self._handle, s.data_ptr(), cmd.data_ptr(), s.shape[1], int(n_ticks),
                                             self._stream()))
        return s

    def reward_done(self, kin: torch.Tensor, steering: Optional[torch.Tensor], book_f64: torch.Tensor,
                    book_i32: torch.Tensor, map_ids: Optional[np.ndarray] = None):
        """a7/a8 stage entry (rd_reward_done): one tick of progress / lap / wrong-way bookkeeping + task reward / done for
        teacher-forced poses.  kin f64 [5, n] = (x, y, yaw, v, slip) after the tick, book_f64 [3, n] = (time, progress,
        last), book_i32 [3, n] = (lap, checkpoint, flags).  Returns (book_f64', book_i32', reward f64[n], done u8[n])."""
        k = kin.to(device=self.device, dtype=torch.float64).contiguous()
        n = k.shape[1]
        st = None if steering is None else steering.to(device=self.device, dtype=torch.float64).contiguous()
        bf = book_f64.to(device=self.device, dtype=torch.float64).contiguous().clone()
        bi = book_i32.to(device=self.device, dtype=torch.int32).contiguous().clone()
        rew = torch.empty(n, dtype=torch.float64, device=self.device)
        done = torch.empty(n, dtype=torch.uint8, device=self.device)
        ids = None if map_ids is None else np.ascontiguousarray(map_ids, dtype=np.int32)
        with torch.cuda.device(self.device):
            self._check(self.lib.rd_reward_done(self._handle, k.data_ptr(), st.data_ptr() if st is not None else None,
                                                ids.ctypes.data if ids is not None else None, n, bf.data_ptr(),
                                                bi.data_ptr(), rew.data_ptr(), done.data_ptr(), self._stream()))
        return bf, bi, rew, done

    # ------------------------------------------------------------------ state / stats
    def get_state(self):
        f = torch.empty((_abi.NF64, self.n), dtype=torch.float64, device=self.device)
        i = torch.empty((_abi.NI32, self.n), dtype=torch.int32, device=self.device)
        with torch.cuda.device(self.device):
            self._check(self.lib.rd_get_state(self._handle, f.data_ptr(), i.data_ptr(), self._stream()))
        return f, i

    def set_state(self, f64: Optional[torch.Tensor], i32: Optional[torch.Tensor]) -> None:
        """Either part may be None (left as it is).  Restoring the float64 part restarts the n_step_progress ring."""
        f = None if f64 is None else f64.to(device=self.device, dtype=torch.float64).contiguous()
        i = None if i32 is None else i32.to(device=self.device, dtype=torch.int32).contiguous()
        if (f is not None and f.shape != (_abi.NF64, self.n)) or (i is not None and i.shape != (_abi.NI32, self.n)):
            raise ValueError("state shapes must be [NF64, n] and [NI32, n]")
        with torch.cuda.device(self.device):
            self._check(self.lib.rd_set_state(self._handle, f.data_ptr() if f is not None else None,
                                              i.data_ptr() if i is not None else None, self._stream()))
            torch.cuda.current_stream(self.device).synchronize()  # f/i may be temporaries

    def read_stats(self, reset: bool = False) -> Dict[str, float]:
        st = _abi.RdStats()
        with torch.cuda.device(self.device):
            self._check(self.lib.rd_read_stats(self._handle, C.byref(st), int(reset), self._stream()))
        return st.as_dict()

    def enable_timing(self, enable: bool = True) -> None:
        self._check(self.lib.rd_enable_timing(self._handle, int(enable)))

    def read_timing(self, reset: bool = False) -> Dict[str, float]:
        t = _abi.RdTiming()
        with torch.cuda.device(self.device):
            self._check(self.lib.rd_read_timing(self._handle, C.byref(t), int(reset)))
        return t.as_dict()

    @property
    def launch_count(self) -> int:
        return int(self.lib.rd_launch_count(self._handle))

    def close(self):
        if getattr(self, "_handle", None) is not None and self._handle.value:
            self.lib.rd_destroy(self._handle)
            self._handle = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
