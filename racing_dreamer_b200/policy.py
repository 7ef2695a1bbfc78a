"""On-device policies for closed-loop rollouts.

``GapFollowerPolicy`` is the reference's follow-the-gap controller
[REF ros_agent/agents/follow_the_gap/src/agent.py:60-238] run for every env of a ``BatchedRaceEnv`` by the CUDA kernel
``k_gap_follower`` behind ``rd_policy_gap_follower`` (include/rd_env.h): scans are read where ``k_lidar`` wrote them and
the actions are written where ``k_step`` reads them, so ``rollout(n)`` advances the whole batch ``n`` steps without a
host round trip.  One controller per env; its state (PID memory, the node's two first-message gates) is cleared by the
env kernels whenever that env is reset.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Optional

import torch

from . import _abi
from .env import BatchedRaceEnv


class GapFollowerPolicy:
    def __init__(self, env: BatchedRaceEnv, **overrides):
        """overrides: fields of ``rd_gap_follower`` (e.g. ``speed_scale=0.6``, ``kp=1.2``)."""
        self.env = env
        g = _abi.RdGapFollower()
        env.lib.rd_gap_follower_defaults(C.byref(env.cfg), C.byref(g))
        for k, v in overrides.items():
            if not hasattr(g, k):
                raise TypeError(f"unknown gap-follower parameter {k!r}")
            setattr(g, k, v)
        self.params = g
        with torch.cuda.device(env.device):
            env._check(env.lib.rd_policy_gap_follower_init(env._handle, C.byref(g)))
        self.actions = torch.zeros((env.n, 2), dtype=torch.float32, device=env.device)
        self.debug = torch.zeros((env.n, 4), dtype=torch.float64, device=env.device)

    def reset(self) -> None:
        """Clears every controller (env resets do this per env on the device already)."""
        with torch.cuda.device(self.env.device):
            self.env._check(self.env.lib.rd_policy_gap_follower_init(self.env._handle, C.byref(self.params)))

    def act(self, lidar: Optional[torch.Tensor] = None, speed: Optional[torch.Tensor] = None,
            debug: bool = False) -> torch.Tensor:
        """lidar f32 [N, n_beams] metres (default: the env's current observation), speed f32 [N] (default: the env's own
        longitudinal speed) -> agent-facing actions f32 [N, 2] (a persistent buffer)."""
        env = self.env
        li = env.buf["lidar"] if lidar is None else lidar
        if li.shape != (env.n, env.n_beams) or li.dtype != torch.float32 or li.device != env.device or not li.is_contiguous():
            raise ValueError(f"lidar must be a contiguous float32 CUDA tensor of shape ({env.n}, {env.n_beams})")
        sp = None
        if speed is not None:
            sp = speed.to(device=env.device, dtype=torch.float32).contiguous()
            if sp.shape != (env.n,):
                raise ValueError("speed must have shape (n_envs,)")
        with torch.cuda.device(env.device):
            env._check(env.lib.rd_policy_gap_follower(env._handle, li.data_ptr(), sp.data_ptr() if sp is not None else None,
                                                      self.actions.data_ptr(), self.debug.data_ptr() if debug else None,
                                                      env._stream()))
        return self.actions

    def rollout(self, n_steps: int):
        """n_steps x (controller -> env.step) enqueued on the current stream; returns the env's (obs, reward, done, info)
        views of the LAST step.  Episode statistics accumulate on the device (env.read_stats())."""
        env = self.env
        with torch.cuda.device(env.device):
            env._check(env.lib.rd_rollout_gap_follower(env._handle, int(n_steps), C.byref(env._out),
                                                       self.actions.data_ptr(), env._stream()))
        return env._obs(), env.buf["reward"], env.buf["done"].bool(), env._info()

    def drive_command(self) -> Dict[str, torch.Tensor]:
        """The node's last published command per env (valid after ``act(debug=True)``)."""
        return {"steering_angle": self.debug[:, 0], "speed": self.debug[:, 1], "heading": self.debug[:, 2],
                "heading_distance": self.debug[:, 3]}
