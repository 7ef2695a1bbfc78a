#!/bin/bash
# k_step tuning round: GPU tests, then the config-2 bench per k_step block size.  usage: bash tools/gpu_stepopt.sh tag
TAG=${1:-s}; OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 900 python -m pytest tests -x -q -m gpu > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 $OUT/pytest_gpu.log
for B in 32 64 128; do
  RD_STEP_BLOCK=$B timeout 600 python bench.py --no-cpu-baseline --steps 300 --warmup 20 > $OUT/bench_b$B.json 2> $OUT/bench_b$B.err
  python - $OUT/bench_b$B.json $B <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1]))
    print('block',sys.argv[2],{k:d[k] for k in ('value','ms_per_step','kernel_ms','ms_per_step_back_to_back')}, 'e2e', d['e2e']['value'], 'closed', d['closed_loop'] and (d['closed_loop']['value'], d['closed_loop']['kernel_ms']))
except Exception as e: print('bench parse failed', e)
PY
  tail -2 $OUT/bench_b$B.err
done
timeout 600 python bench.py --no-cpu-baseline --config 4 --steps 100 --warmup 10 > $OUT/bench_c4.json 2> $OUT/bench_c4.err; python -c "
import json; d=json.load(open('$OUT/bench_c4.json')); print('config4', d['value'], d['kernel_ms'], d['closed_loop'] and d['closed_loop']['value'])"
