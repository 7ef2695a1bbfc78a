#!/usr/bin/env python
"""Workload for compute-sanitizer (memcheck / racecheck / synccheck / initcheck), SURVEY.md §5.

Runs, at sizes a sanitizer finishes in minutes: smoke() (64 envs with lidar_occupancy), one config-5-shaped step
(two maps, Barcelona + Austria, ragged batch), one lidar_occupancy step on Columbia, a multi-agent world step, a
follow-the-gap and a Dreamer rollout, and the host-facing chunked step.  Exits non-zero on any Python-side failure;
the sanitizer's own verdict is its exit code (--error-exitcode).

usage: compute-sanitizer --tool memcheck --error-exitcode 9 python tools/sanitize_workload.py [part ...]
       parts: smoke config5 occupancy multi gap dreamer host   (default: all)
"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))


def main():
    import torch
    from racing_dreamer_b200 import BatchedRaceEnv, EnvConfig, GapFollowerPolicy, DreamerPolicy
    from racing_dreamer_b200.host import HostSteppedEnv
    parts = sys.argv[1:] or ["smoke", "config5", "occupancy", "multi", "gap", "dreamer", "host"]
    dev = torch.device("cuda:0")
    torch.cuda.set_device(0)
    rng = np.random.RandomState(0)

    def run(ec, steps=2):
        env = BatchedRaceEnv(ec, device=dev)
        env.reset()
        for _ in range(steps):
            a = torch.from_numpy(rng.uniform(-1, 1, (env.n, 2)).astype(np.float32)).to(dev)
            env.step(a)
        torch.cuda.synchronize()
        env.read_stats()
        env.close()

    if "smoke" in parts:
        import __graft_entry__ as g
        g.smoke()
    if "config5" in parts:   # two maps, 32-warp and 16-warp k_lidar instantiations, ragged env count
        run(EnvConfig(tracks=("barcelona", "austria"), n_envs=1537, action_repeat=8, obs_type="lidar", auto_reset=True,
                      reset_mode="random", seed=5, time_limit_steps=3))
    if "occupancy" in parts:
        run(EnvConfig(tracks=("columbia",), n_envs=160, action_repeat=8, obs_type="lidar_occupancy", auto_reset=True,
                      reset_mode="random", seed=3))
    if "multi" in parts:
        run(EnvConfig(tracks=("austria",), n_envs=512, action_repeat=4, obs_type="lidar", auto_reset=True,
                      reset_mode="random_ball", seed=6, agents_per_world=4,
                      agent_tasks=("maximize_progress",) + ("n_step_progress",) * 3, time_limit_steps=3))
    if "gap" in parts:
        env = BatchedRaceEnv(EnvConfig(tracks=("austria",), n_envs=256, action_repeat=8, auto_reset=True,
                                       reset_mode="random", seed=7), device=dev)
        pol = GapFollowerPolicy(env)
        env.reset()
        pol.rollout(3)
        torch.cuda.synchronize()
        env.close()
    if "dreamer" in parts:
        env = BatchedRaceEnv(EnvConfig(tracks=("austria",), n_envs=256, action_repeat=8, auto_reset=True,
                                       reset_mode="random", seed=8), device=dev)
        pol = DreamerPolicy(env, "austria_dreamer", noise="philox")
        env.reset()
        pol.rollout(3)
        torch.cuda.synchronize()
        env.close()
    if "host" in parts:
        henv = HostSteppedEnv(EnvConfig(tracks=("austria",), n_envs=300, action_repeat=8, obs_type="lidar_occupancy",
                                        auto_reset=True, reset_mode="random", seed=9), device=dev, n_shards=4)
        henv.reset()
        for _ in range(2):
            henv.step(rng.uniform(-1, 1, (300, 2)).astype(np.float32))
        henv.close()
    print("sanitize workload ok:", " ".join(parts))


if __name__ == "__main__":
    main()
