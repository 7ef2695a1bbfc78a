#!/bin/bash
# k_occupancy CTA-size variants (variants/occ*.so): parity tests of the occupancy stage, then its per-launch time
for v in variants/occ*.so; do
  echo "== $v"
  RD_ENV_LIB=$PWD/$v timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "occupancy or golden_replay or mixed_maps" 2>&1 | tail -2
  for rep in 1 2; do
    RD_ENV_LIB=$PWD/$v timeout 300 python bench.py --obs lidar_occupancy --envs 16384 --steps 40 --warmup 5 --no-cpu-baseline --no-closed-loop --no-multi-agent --e2e-steps 3 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('  ', round(d['value']/1e6,3), 'M env-steps/s', round(d['ms_per_step'],3), 'ms', d['kernel_ms'])"
  done
done
