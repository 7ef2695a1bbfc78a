#!/usr/bin/env python
"""Where the host-facing step's time goes: rd_step_host ms/step over shard counts and the zero-copy option, next to the
PCIe floor measured with the same chunking (18 MB device->host in k back-to-back pinned copies)."""
import os
import sys
import time
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from racing_dreamer_b200 import EnvConfig  # noqa: E402
from racing_dreamer_b200.host import HostSteppedEnv  # noqa: E402

torch.cuda.set_device(0)
n = 4096
nbytes = n * 1080 * 4
d = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
h = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
for k in (1, 8, 16):
    b = [nbytes * i // k for i in range(k + 1)]
    for rep in range(3):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(50):
            for i in range(k):
                h[b[i]:b[i + 1]].copy_(d[b[i]:b[i + 1]], non_blocking=True)
            torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / 50
    print(f"d2h 17.7 MB in {k:2d} copies + sync: {dt * 1e3:.3f} ms ({nbytes / dt / 1e9:.1f} GB/s)", flush=True)

a = np.stack([np.full(n, 0.6), 0.8 * np.sin(np.random.RandomState(0).uniform(0, 6.28, n))], 1).astype(np.float32)
for spec in ("", "1,15", "1,3,12", "1,4,11", "1,2,4,8", "2,6,8", "1,2,5,8", "1,3,6,6", "1,2,3,4,6", "1,7,8", "1,5,10", "3,13", "1,1,2,4,8"):
    os.environ.pop("RD_HOST_CHUNKS", None)
    if spec:
        os.environ["RD_HOST_CHUNKS"] = spec
    ec = EnvConfig(tracks=("austria",), n_envs=n, action_repeat=8, auto_reset=True, reset_mode="random", seed=1,
                   time_limit_steps=250)
    env = HostSteppedEnv(ec, device="cuda:0", n_shards=8)
    env.reset()
    for _ in range(30):
        env.step(a)
    best = 1e9
    for rep in range(4):
        t0 = time.perf_counter()
        for _ in range(200):
            env.step(a)
        best = min(best, (time.perf_counter() - t0) / 200 * 1e3)
    print(f"chunks {spec or 'default(8)':12s}: e2e {best:.3f} ms/step", flush=True)
    env.close()

for dt in ("float32", "float16"):
    os.environ.pop("RD_HOST_CHUNKS", None)
    ec = EnvConfig(tracks=("austria",), n_envs=n, action_repeat=8, auto_reset=True, reset_mode="random", seed=1,
                   time_limit_steps=250, lidar_dtype=dt)
    env = HostSteppedEnv(ec, device="cuda:0", n_shards=8)
    env.reset()
    for _ in range(30):
        env.step(a)
    best = 1e9
    for rep in range(4):
        t0 = time.perf_counter()
        for _ in range(200):
            env.step(a)
        best = min(best, (time.perf_counter() - t0) / 200 * 1e3)
    print(f"lidar_dtype {dt}: e2e {best:.3f} ms/step = {n / best / 1e3:.2f} M env-steps/s, {env.d2h_bytes_per_step / 1e6:.1f} MB d2h per step", flush=True)
    env.close()
