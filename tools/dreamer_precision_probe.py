#!/usr/bin/env python
"""Dreamer agent step time at both precisions (tf32x3 = three TF32 passes, float32-grade; tf32 = one pass): tells
whether k_dense is bound by operand traffic / MMA work (time ~ passes) or by launch and pipeline latency (time flat).
usage (GPU box): python tools/dreamer_precision_probe.py"""
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import bench  # noqa: E402
from racing_dreamer_b200 import BatchedRaceEnv, DreamerPolicy  # noqa: E402


def main():
    wl = bench.workload_of(2)
    for n in (4096, 16384):
        for prec in ("tf32x3", "tf32"):
            env = BatchedRaceEnv(bench.env_config(wl, n, 0), device="cuda:0")
            pol = DreamerPolicy(env, "austria_dreamer", noise="philox", precision=prec)
            env.reset()
            pol.rollout(20)
            torch.cuda.synchronize()
            env.enable_timing(True)
            env.read_timing(reset=True)
            c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            c0.record()
            pol.rollout(100)
            c1.record()
            torch.cuda.synchronize()
            t = env.read_timing(reset=True)
            print(f"n={n} {prec}: policy {t['policy_ms'] / max(1, t['policy_launches']):.4f} ms per agent step, "
                  f"closed loop {c0.elapsed_time(c1) / 100:.4f} ms per step", flush=True)
            env.close()


if __name__ == "__main__":
    main()
