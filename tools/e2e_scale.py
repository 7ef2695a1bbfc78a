#!/usr/bin/env python
"""Host-facing step (rd_step_host) under torchrun: per-rank ms/step with and without CPU/NUMA binding.
usage: python -m torch.distributed.run --nproc-per-node N tools/e2e_scale.py   (prints one line per rank and variant)"""
import os
import sys
import time
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from racing_dreamer_b200 import EnvConfig  # noqa: E402
from racing_dreamer_b200.host import HostSteppedEnv  # noqa: E402

rank, local = int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0))
world = int(os.environ.get("WORLD_SIZE", 1))
torch.cuda.set_device(local)
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("gloo")
n = 4096
ec = EnvConfig(tracks=("austria",), n_envs=n, action_repeat=8, auto_reset=True, reset_mode="random", seed=1,
               env_id_offset=rank * n, time_limit_steps=250)
a = np.stack([np.full(n, 0.6), 0.8 * np.sin(np.random.RandomState(rank).uniform(0, 6.28, n))], 1).astype(np.float32)
all_cpus = sorted(os.sched_getaffinity(0))
for bind in (False, True, False, True):
    os.sched_setaffinity(0, all_cpus)
    env = HostSteppedEnv(ec, device=f"cuda:{local}", n_shards=8, bind_cpu=bind)
    env.reset()
    for _ in range(20):
        env.step(a)
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(200):
        env.step(a)
    dt = (time.perf_counter() - t0) / 200 * 1e3
    print(f"rank {rank} bind={int(bind)} cpus={len(env.cpus) if env.cpus else len(all_cpus)} e2e {dt:.3f} ms/step", flush=True)
    env.close()
    if world > 1:
        dist.barrier()
