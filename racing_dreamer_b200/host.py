"""HostSteppedEnv -- the env step for callers whose actions and observations live in HOST memory.

This is the call a reference user makes: numpy actions in, numpy observation dict out
[REF dreamer/tools.py:178-195 simulate(): obs = env.step(actions) with numpy arrays].  The batch is split into
`n_shards` independent BatchedRaceEnv handles, each on its own CUDA stream, so that the device->host copy of shard k's
observations overlaps with the kernels of the other shards.  Pinned staging buffers are allocated once.  Shards keep global env ids (env_id_offset), so results are identical
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
    """See module docstring.  Device results live in ONE key-major allocation per key for the whole batch; each shard's
    kernels write their slice of it.  Per step the host issues one H2D copy (all actions), per shard two kernels and one
    D2H copy of its LiDAR rows (plus its occupancy rows), and one D2H copy of the slab that holds every small array."""

    _BIG = ("lidar", "occupancy")

    def __init__(self, config: EnvConfig, device=None, n_shards: int = 4, copy_back=_OUT_KEYS):
        n = int(config.n_envs)
        n_shards = max(1, min(int(n_shards), n))
        bounds = np.linspace(0, n, n_shards + 1).astype(np.int64)
        self.n = n
        self.device = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
        ntr = len(config.tracks)
        all_ids = np.asarray(config.map_ids, np.int32) if config.map_ids is not None else (np.arange(n) % ntr).astype(np.int32)
        occ = config.obs_type == "lidar_occupancy"
        nb = int(config.n_beams)
        # ---- device + pinned host storage, key-major over the whole batch ----
        self.dev: Dict[str, torch.Tensor] = {}
        self.host: Dict[str, torch.Tensor] = {}
        small = [(k, sh, dt) for k, sh, dt in BatchedRaceEnv.OUT_SPEC if k not in self._BIG]
        offs, total = {}, 0
        for k, sh, dt in small:
            nbytes = n * int(np.prod(sh, dtype=np.int64)) * torch.empty((), dtype=dt).element_size()
            offs[k] = (total, nbytes)
            total = (total + nbytes + 255) & ~255
        self._small_dev = torch.zeros(total, dtype=torch.uint8, device=self.device)
        self._small_host = torch.zeros(total, dtype=torch.uint8, pin_memory=True)
        for k, sh, dt in small:
            o, nbytes = offs[k]
            self.dev[k] = self._small_dev[o:o + nbytes].view(dt).view((n,) + tuple(sh))
            self.host[k] = self._small_host[o:o + nbytes].view(dt).view((n,) + tuple(sh))
        self.dev["lidar"] = torch.zeros((n, nb), dtype=torch.float32, device=self.device)
        self.host["lidar"] = torch.zeros((n, nb), dtype=torch.float32, pin_memory=True)
        if occ:
            self.dev["occupancy"] = torch.zeros((n, 64, 64, 1), dtype=torch.uint8, device=self.device)
            self.host["occupancy"] = torch.zeros((n, 64, 64, 1), dtype=torch.uint8, pin_memory=True)
        self.copy_back = tuple(k for k in copy_back if k in self.host)
        self._copy_small = any(k not in self._BIG for k in self.copy_back)
        self._big_keys = tuple(k for k in self._BIG if k in self.copy_back)
        # ---- shards ----
        self.shards: List[BatchedRaceEnv] = []
        self.slices = []
        for k in range(n_shards):
            a, b = int(bounds[k]), int(bounds[k + 1])
            ec = dataclasses.replace(config, n_envs=b - a, env_id_offset=int(config.env_id_offset) + a,
                                     map_ids=all_ids[a:b].tolist())
            bufs = {key: t[a:b] for key, t in self.dev.items()}
            self.shards.append(BatchedRaceEnv(ec, device=self.device, out_buffers=bufs))
            self.slices.append(slice(a, b))
        self.streams = [torch.cuda.Stream(device=self.device) for _ in self.shards]
        self._ev_act = torch.cuda.Event()
        self._ev_done = [torch.cuda.Event() for _ in self.shards]
        self.host_np = {k: self.host[k].numpy() for k in self.copy_back}
        self.actions_pinned = torch.empty((n, 2), dtype=torch.float32, pin_memory=True)
        self.actions_dev = torch.empty((n, 2), dtype=torch.float32, device=self.device)
        self.h2d_bytes_per_step = n * 2 * 4
        self.d2h_bytes_per_step = int(sum(self.host[k].numel() * self.host[k].element_size() for k in self._big_keys)
                                      + (self._small_host.numel() if self._copy_small else 0))

    # ------------------------------------------------------------------
    def _run(self, launch) -> Dict[str, np.ndarray]:
        """launch(k, env): enqueue shard k's kernels on the current stream."""
        for k, env in enumerate(self.shards):
            st = self.streams[k]
            with torch.cuda.stream(st):
                if k > 0:
                    st.wait_event(self._ev_act)
                launch(k, env)
                sl = self.slices[k]
                for key in self._big_keys:
                    self.host[key][sl].copy_(self.dev[key][sl], non_blocking=True)
                if k > 0:
                    self._ev_done[k].record(st)
        st0 = self.streams[0]
        with torch.cuda.stream(st0):
            for k in range(1, len(self.shards)):
                st0.wait_event(self._ev_done[k])
            if self._copy_small:
                self._small_host.copy_(self._small_dev, non_blocking=True)
        st0.synchronize()
        return self.host_np

    def reset(self, mode: Optional[str] = None) -> Dict[str, np.ndarray]:
        with torch.cuda.stream(self.streams[0]):
            self._ev_act.record(self.streams[0])
        return self._run(lambda k, env: env.reset(mode=mode))

    def step(self, actions: np.ndarray) -> Dict[str, np.ndarray]:
        """actions: float32 [n_envs, 2] in host memory.  Returns numpy views of the pinned result buffers
        (valid until the next call): lidar, pose, velocity, speed, reward, done, progress, lap, time, flags
        (+ occupancy for obs_type='lidar_occupancy')."""
        self.actions_pinned.numpy()[...] = actions
        st0 = self.streams[0]
        with torch.cuda.stream(st0):
            self.actions_dev.copy_(self.actions_pinned, non_blocking=True)
            self._ev_act.record(st0)
        base = self.actions_dev.data_ptr()
        return self._run(lambda k, env: env.step_raw(base + self.slices[k].start * 8))

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
