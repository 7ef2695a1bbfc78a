#!/usr/bin/env python
"""BASELINE config 1 / BASELINE.md C1: the reference's UNMODIFIED wrapper stack over the one-tick oracle env, 1 env,
Columbia, 1000 random-action steps, 1 core.  Needs /root/reference (build container only) -- the GPU box has no reference
tree, so bench.py reports the file this writes (profiles/r2_reference_stack_cpu.json) when it cannot run the stack live.
usage: python tools/time_reference_stack.py [steps]"""
import json
import os
import platform
import sys
from pathlib import Path

os.environ.setdefault("OMP_NUM_THREADS", "1")
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import bench  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
best = None
for rep in range(3):
    r = bench.reference_stack_baseline(steps)
    assert "value" in r and r["where"].startswith("live"), r
    if best is None or r["value"] > best["value"]:
        best = r
best["host"] = f"{platform.processor() or platform.machine()}, {os.cpu_count()} cores (build container)"
best["note"] = "best of 3 runs; the stack is single-threaded Python + NumPy/SciPy/Pillow (OccupancyMapObs dominates)"
(ROOT / "profiles" / "r2_reference_stack_cpu.json").write_text(json.dumps(best, indent=1) + "\n")
print(json.dumps(best))
