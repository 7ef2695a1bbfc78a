#!/bin/bash
# round-2 GPU iteration: parity tests, smoke, bench line, optional library variants A/B, optional ncu of one kernel.
# usage: bash tools/gpu_r2.sh tag [kernel-regex] [bench args...]
TAG=${1:-r2}; K=$2; shift 2
OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 1500 python -m pytest tests -x -q -m gpu > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -6 $OUT/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 $OUT/smoke.log
show() { python - "$1" <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1]))
    print({k:d.get(k) for k in ('value','ms_per_step','kernel_ms','ms_per_step_back_to_back')}, 'e2e', d['e2e']['value'], d['e2e'].get('ms_per_step'), 'roof', d['roofline']['frac'])
except Exception as e: print('bench parse failed', e)
PY
}
timeout 900 python bench.py --no-cpu-baseline "$@" > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?"; show $OUT/bench.json; tail -3 $OUT/bench.err
for v in variants/*.so; do
  [ -f "$v" ] || continue
  b=$(basename $v .so)
  RD_ENV_LIB=$PWD/$v timeout 600 python bench.py --no-cpu-baseline --no-closed-loop --no-multi-agent --steps 300 --warmup 20 --e2e-steps 20 > $OUT/bench_$b.json 2> $OUT/bench_$b.err; echo "== variant $b rc=$?"; show $OUT/bench_$b.json
  timeout 600 python bench.py --no-cpu-baseline --no-closed-loop --no-multi-agent --steps 300 --warmup 20 --e2e-steps 20 > $OUT/bench_base_vs_$b.json 2> /dev/null; echo "== base"; show $OUT/bench_base_vs_$b.json
done
if [ -n "$K" ] && [ "$K" != "-" ]; then
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$K -s 5 -c 1 -o $OUT/prof_$K -f \
     python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-closed-loop --no-multi-agent --e2e-steps 2 > $OUT/ncu_$K.log 2>&1; echo "ncu rc=$?"
fi
ls $OUT
