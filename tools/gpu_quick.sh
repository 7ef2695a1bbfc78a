#!/bin/bash
# quick GPU iteration: parity tests + one bench line (+ optional ncu of one kernel).  usage: bash tools/gpu_quick.sh tag [kernel-regex]
TAG=${1:-q}; K=$2
OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 900 python -m pytest tests -x -q -m gpu > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 $OUT/pytest_gpu.log
timeout 600 python bench.py --no-cpu-baseline > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?"; python - $OUT/bench.json <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1]))
    print({k:d[k] for k in ('value','ms_per_step','kernel_ms','ms_per_step_back_to_back')}, 'e2e', d['e2e']['value'], 'roof', d['roofline']['frac'])
except Exception as e: print('bench parse failed', e)
PY
tail -3 $OUT/bench.err
if [ -n "$K" ]; then
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$K -s 5 -c 1 -o $OUT/prof_$K -f \
     python bench.py --steps 10 --warmup 3 --no-cpu-baseline --e2e-steps 2 > $OUT/ncu_$K.log 2>&1; echo "ncu rc=$?"
fi
