#!/bin/bash
# what does a k_lidar launch over 64 envs spend its 34 us on?  ncu duration + full capture with source view
OUT=gpurun_out/${1:-lsmall}; mkdir -p $OUT
cat > /tmp/small.py <<'PY'
import sys, numpy as np, torch
sys.path.insert(0, '.')
from racing_dreamer_b200 import BatchedRaceEnv, EnvConfig
n = int(sys.argv[1])
env = BatchedRaceEnv(EnvConfig(tracks=("austria",), n_envs=n, action_repeat=8, auto_reset=True, reset_mode="random", seed=1), device="cuda:0")
a = torch.zeros((n, 2), device="cuda"); a[:, 0] = 0.6
env.reset()
for _ in range(12): env.step(a)
torch.cuda.synchronize()
PY
timeout 600 ncu --metrics gpu__time_duration.sum,sm__cycles_elapsed.max,smsp__inst_executed.sum --clock-control none -k regex:k_lidar -s 6 -c 3 python /tmp/small.py 64 2>&1 | grep -E "k_lidar|gpu__time|cycles_elapsed|inst_executed" | head -12
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_lidar -s 8 -c 1 -o $OUT/prof_lidar64 -f python /tmp/small.py 64 > $OUT/ncu64.log 2>&1; echo "ncu rc=$?"
