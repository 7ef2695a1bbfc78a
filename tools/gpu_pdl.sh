#!/bin/bash
# k_lidar variants: programmatic dependent launch on/off (RD_LIDAR_PDL), after the GPU parity suite.  usage: bash tools/gpu_pdl.sh tag
TAG=${1:-pdl}; OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 900 python -m pytest tests -x -q -m gpu > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 $OUT/pytest_gpu.log
for v in 0 1 0 1; do
  RD_LIDAR_PDL=$v timeout 300 python bench.py --steps 1000 --warmup 200 --no-cpu-baseline --no-closed-loop --no-multi-agent > $OUT/bench_pdl$v.json 2> $OUT/bench_pdl$v.err
  python - $OUT/bench_pdl$v.json $v <<'PY'
import json,sys
d=json.load(open(sys.argv[1]))
print('pdl',sys.argv[2],{k:d[k] for k in ('value','ms_per_step','kernel_ms','ms_per_step_back_to_back')}, 'e2e', d['e2e']['value'], d['e2e']['ms_per_step'])
PY
done
timeout 600 python bench.py --no-cpu-baseline > $OUT/bench_full.json 2> $OUT/bench_full.err; echo "bench rc=$?"; cat $OUT/bench_full.json | python -c "
import json,sys
d=json.load(sys.stdin); print(d['value'], d['ms_per_step'], d['kernel_ms']); print('multi', d['multi_agent']); print('closed', {k:(v['value'], v['ms_per_step']) for k,v in d['closed_loop'].items()})"
tail -3 $OUT/bench_full.err
