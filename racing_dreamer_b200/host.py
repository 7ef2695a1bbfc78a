"""HostSteppedEnv -- the env step for callers whose actions and observations live in HOST memory.

This is the call a reference user makes: numpy actions in, numpy observation dict out
[REF dreamer/tools.py:178-195 simulate(): obs = env.step(actions) with numpy arrays].  The batch is split into
`n_shards` independent BatchedRaceEnv handles, each on its own CUDA stream, so that the host->device copy of shard
k+1's actions and the device->host copy of shard k's observations overlap with the kernels of the other shards.
Pinned staging buffers are allocated once.  Shards keep global env ids (env_id_offset), so results are identical
to one big batch.
"""
from __future__ import annotations

import dataclasses
from typing import Dict, List, Optional

import numpy as np
import torch

from .env import BatchedRaceEnv, EnvConfig

_OUT_KEYS = ("lidar", "occupancy", "pose", "velocity", "speed", "reward", "done", "progress", "lap", "time", "flags")


class HostSteppedEnv:
    def __init__(self, config: EnvConfig, device=None, n_shards: int = 4, copy_back=_OUT_KEYS):
        n = int(config.n_envs)
        n_shards = max(1, min(int(n_shards), n))
        bounds = np.linspace(0, n, n_shards + 1).astype(np.int64)
        self.n = n
        self.device = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
        self.shards: List[BatchedRaceEnv] = []
        self.slices = []
        ntr = len(config.tracks)
        all_ids = np.asarray(config.map_ids, np.int32) if config.map_ids is not None else (np.arange(n) % ntr).astype(np.int32)
        for k in range(n_shards):
            a, b = int(bounds[k]), int(bounds[k + 1])
            ec = dataclasses.replace(config, n_envs=b - a, env_id_offset=int(config.env_id_offset) + a,
                                     map_ids=all_ids[a:b].tolist())
            self.shards.append(BatchedRaceEnv(ec, device=self.device))
            self.slices.append(slice(a, b))
        self.streams = [torch.cuda.Stream(device=self.device) for _ in self.shards]
        self.copy_back = tuple(k for k in copy_back if self.shards[0].buf.get(k) is not None)
        ref = self.shards[0].buf
        self.host: Dict[str, torch.Tensor] = {
            k: torch.empty((n,) + tuple(ref[k].shape[1:]), dtype=ref[k].dtype, pin_memory=True) for k in self.copy_back}
        self.host_np = {k: v.numpy() for k, v in self.host.items()}
        self.actions_pinned = torch.empty((n, 2), dtype=torch.float32, pin_memory=True)
        self.actions_dev = [torch.empty((s.stop - s.start, 2), dtype=torch.float32, device=self.device) for s in self.slices]
        self.h2d_bytes_per_step = n * 2 * 4
        self.d2h_bytes_per_step = int(sum(v.numel() * v.element_size() for v in self.host.values()))

    # ------------------------------------------------------------------
    def _copy_back(self, k: int):
        sl = self.slices[k]
        for key in self.copy_back:
            self.host[key][sl].copy_(self.shards[k].buf[key], non_blocking=True)

    def reset(self, mode: Optional[str] = None) -> Dict[str, np.ndarray]:
        for k, env in enumerate(self.shards):
            with torch.cuda.stream(self.streams[k]):
                env.reset(mode=mode)
                self._copy_back(k)
        for st in self.streams:
            st.synchronize()
        return self.host_np

    def step(self, actions: np.ndarray) -> Dict[str, np.ndarray]:
        """actions: float32 [n_envs, 2] in host memory.  Returns numpy views of the pinned result buffers
        (valid until the next call): lidar, pose, velocity, speed, reward, done, progress, lap, time, flags
        (+ occupancy for obs_type='lidar_occupancy')."""
        self.actions_pinned.numpy()[...] = actions
        for k, env in enumerate(self.shards):
            with torch.cuda.stream(self.streams[k]):
                self.actions_dev[k].copy_(self.actions_pinned[self.slices[k]], non_blocking=True)
                env.step_raw(self.actions_dev[k].data_ptr())
                self._copy_back(k)
        for st in self.streams:
            st.synchronize()
        return self.host_np

    @property
    def launch_count(self) -> int:
        return sum(e.launch_count for e in self.shards)

    def read_stats(self, reset: bool = False) -> Dict[str, float]:
        tot: Dict[str, float] = {}
        for k, e in enumerate(self.shards):
            with torch.cuda.stream(self.streams[k]):
                for key, v in e.read_stats(reset).items():
                    tot[key] = tot.get(key, 0.0) + v
        return tot

    def close(self):
        for e in self.shards:
            e.close()
