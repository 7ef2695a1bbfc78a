"""HostSteppedEnv -- the env step for callers whose actions and observations live in HOST memory.

This is the call a reference user makes: numpy actions in, numpy observation dict out
[REF dreamer/tools.py:178-195 simulate(): obs = env.step(actions) with numpy arrays].  It is a thin wrapper over the
C ABI's host-facing entry points (include/rd_env.h: rd_host_init / rd_reset_host / rd_step_host): the library owns the
device result buffers and their pinned host mirrors, copies the actions in, runs k_step once over the batch, then the
observation kernels chunk by chunk on internal streams so that chunk k's device->host copy overlaps chunk k+1's ray
casting, and returns when every result is in host memory.  The arrays handed back are zero-copy numpy views of the
pinned mirrors (valid until the next call).
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Dict, Optional

import numpy as np
import torch

from . import _abi
from .env import BatchedRaceEnv, EnvConfig

_OUT_KEYS = ("lidar", "occupancy", "pose", "velocity", "speed", "reward", "done", "progress", "lap", "time", "flags",
             "rank", "opponents", "wrong_way", "wall_collision")
_FIELDS = {"lidar": ("lidar_dev", np.float32), "occupancy": ("occupancy_dev", np.uint8), "pose": ("pose_dev", np.float32),
           "velocity": ("velocity_dev", np.float32), "speed": ("speed_dev", np.float32), "reward": ("reward_dev", np.float32),
           "done": ("done_dev", np.uint8), "progress": ("progress_dev", np.float32), "lap": ("lap_dev", np.int32),
           "time": ("time_dev", np.float32), "flags": ("flags_dev", np.uint8),
           "rank": ("rank_dev", np.int32), "opponents": ("opponents_dev", np.uint8),
           # the two info booleans the reference's consumers read, written as 0 / 1 bytes by the step kernel
           "wrong_way": ("wrong_way_dev", np.bool_), "wall_collision": ("wall_collision_dev", np.bool_)}


def bind_to_gpu_cpus(device_index: int) -> Optional[list]:
    """Restrict this process to the CPU cores NVML reports as local to GPU `device_index` (same NUMA node / PCIe root),
    so that the pinned result buffers allocated next are first-touched on that node and the device->host copies of one
    rank do not cross the socket interconnect.  Returns the core list, or None when NVML / sched_setaffinity is missing
    or the mask is empty (then nothing is changed)."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(int(device_index))
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
        cpus = [64 * w + b for w, word in enumerate(words) for b in range(64) if (int(word) >> b) & 1]
        allowed = os.sched_getaffinity(0)
        cpus = [c for c in cpus if c in allowed]
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return cpus
    except Exception:
        return None


class HostSteppedEnv:
    def __init__(self, config: EnvConfig, device=None, n_shards: int = 8, bind_cpu: bool = False):
        """n_shards: env-chunks the observation kernels and their device->host copies are pipelined over.
        bind_cpu: pin this PROCESS to the GPU's local cores first (multi-GPU boxes, one process per GPU)."""
        self.cpus = None
        if bind_cpu:
            import torch as _t
            d = _t.device(device if device is not None else f"cuda:{_t.cuda.current_device()}")
            self.cpus = bind_to_gpu_cpus(d.index if d.index is not None else _t.cuda.current_device())
        self.env = BatchedRaceEnv(config, device=device)
        self.n = self.env.n
        self.device = self.env.device
        self.lib = self.env.lib
        ho = _abi.RdOutputs()
        with torch.cuda.device(self.device):
            self.env._check(self.lib.rd_host_init(self.env._handle, int(n_shards), C.byref(ho)))
        n, nb = self.n, self.env.n_beams
        shapes = {"lidar": (n, nb), "occupancy": (n, 64, 64, 1), "pose": (n, 6), "velocity": (n, 6)}
        self.host_np: Dict[str, np.ndarray] = {}
        for key in _OUT_KEYS:
            field, dt = _FIELDS[key]
            if key == "lidar" and (self.env.cfg.obs_flags & _abi.OBS_LIDAR_F16):
                dt = np.float16
            ptr = getattr(ho, field)
            if not ptr:
                continue
            shape = shapes.get(key, (n,))
            nbytes = int(np.prod(shape)) * np.dtype(dt).itemsize
            buf = (C.c_char * nbytes).from_address(ptr)
            self.host_np[key] = np.frombuffer(buf, dtype=dt).reshape(shape)
        self.n_shards = max(1, min(int(n_shards), n))
        self.h2d_bytes_per_step = n * 2 * 4
        self.d2h_bytes_per_step = int(sum(v.nbytes for v in self.host_np.values()))

    def reset(self, mask: Optional[np.ndarray] = None, mode: Optional[str] = None) -> Dict[str, np.ndarray]:
        m = _abi.RESET_MODES[mode] if mode is not None else int(self.env.cfg.reset_mode)
        mp = None
        if mask is not None:
            mask = np.ascontiguousarray(mask, dtype=np.uint8)
            if mask.shape != (self.n,):
                raise ValueError("mask must have shape (n_envs,)")
            mp = mask.ctypes.data
        with torch.cuda.device(self.device):
            self.env._check(self.lib.rd_reset_host(self.env._handle, mp, m))
        return self.host_np

    def step(self, actions: np.ndarray) -> Dict[str, np.ndarray]:
        """actions: float32 [n_envs, 2] in host memory.  Returns numpy views of the pinned result buffers
        (valid until the next call): lidar, pose, velocity, speed, reward, done, progress, lap, time, flags
        (+ occupancy for obs_type='lidar_occupancy')."""
        a = np.ascontiguousarray(actions, dtype=np.float32)
        if a.shape != (self.n, 2):
            raise ValueError(f"actions must have shape ({self.n}, 2), got {a.shape}")
        self.env._check(self.lib.rd_step_host(self.env._handle, a.ctypes.data))
        return self.host_np

    def step_async(self, actions: np.ndarray) -> None:
        """Enqueue one step (actions are copied now) and return without waiting -- ``step_wait()`` delivers the results.
        With two HostSteppedEnv groups a caller overlaps group A's kernels and policy evaluation with group B's
        device->host copy (the asynchronous vector-env pattern); ``step()`` == ``step_async()`` + ``step_wait()``."""
        a = np.ascontiguousarray(actions, dtype=np.float32)
        if a.shape != (self.n, 2):
            raise ValueError(f"actions must have shape ({self.n}, 2), got {a.shape}")
        self.env._check(self.lib.rd_step_host_begin(self.env._handle, a.ctypes.data))

    def step_wait(self) -> Dict[str, np.ndarray]:
        self.env._check(self.lib.rd_step_host_end(self.env._handle))
        return self.host_np

    @property
    def launch_count(self) -> int:
        return self.env.launch_count

    @property
    def shards(self):
        return range(self.n_shards)

    def read_stats(self, reset: bool = False) -> Dict[str, float]:
        return self.env.read_stats(reset)

    def close(self):
        self.env.close()
