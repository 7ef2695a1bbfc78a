#!/bin/bash
# L2 persistence of the track A/B (RD_L2_PERSIST=0/1) on the flushed bench, then the GPU parity suite
for rep in 1 2; do for v in 0 1; do
  RD_L2_PERSIST=$v timeout 300 python bench.py --steps 1000 --warmup 200 --no-cpu-baseline --no-closed-loop --no-multi-agent 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('l2_persist $v', round(d['value']/1e6,2), 'M env-steps/s', round(d['ms_per_step'],4), 'ms', d['kernel_ms'], 'b2b', round(d['ms_per_step_back_to_back'],4), 'e2e', round(d['e2e']['ms_per_step'],4))"
done; done
timeout 700 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
