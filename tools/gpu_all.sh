#!/bin/bash
# tests + all four BASELINE configs on one GPU
OUT=gpurun_out/${1:-all}; mkdir -p $OUT
timeout 900 python -m pytest tests -x -q -m gpu > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 $OUT/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
for c in 2 3 4 5; do
  st=300; [ $c -ge 3 ] && st=60; [ $c -eq 5 ] && st=30
  python bench.py --config $c --steps $st --warmup 10 --no-cpu-baseline > $OUT/bench_c$c.json 2>$OUT/err_c$c.log
  python -c "import json,sys; d=json.load(open('$OUT/bench_c$c.json')); print('config$c', round(d['value']), 'env-steps/s', d['ms_per_step'], d['kernel_ms'], 'e2e', round(d['e2e']['value']), d['e2e']['ms_per_step'], 'floor', d['e2e']['d2h_floor_ms'], 'launches', d['gpu_launches'])"; tail -2 $OUT/err_c$c.log
done
