#!/bin/bash
# what the driver runs at round end: both arms with its flags, plus the GPU tests.  usage: bash tools/gpu_bench_driver.sh tag [N]
TAG=${1:-drv}; N=${2:-1}
OUT=gpurun_out/$TAG; mkdir -p $OUT
if [ "$N" = "1" ]; then
  [ -n "$SKIP_TESTS" ] || { timeout 1500 python -m pytest tests -x -q -m gpu > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 $OUT/pytest_gpu.log; }
  time python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > $OUT/bench_ref.json 2> $OUT/bench_ref.err; echo "ref rc=$?"
  time python bench.py --gpus 1 --steps 20 --warmup 5 > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?"
else
  time python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 > $OUT/bench_n$N.json 2> $OUT/bench_n$N.err; echo "bench N=$N rc=$?"
  cp $OUT/bench_n$N.json $OUT/bench.json
fi
python - $OUT/bench.json <<'PY'
import json,sys
d=json.load(open(sys.argv[1]))
print('config2', round(d['value']), d['ms_per_step'], d['kernel_ms'], 'e2e', round(d['e2e']['value']), d['e2e']['ms_per_step'], 'floor', d['e2e']['d2h_floor_ms'])
print('e2e_f16', d['e2e_f16'] and round(d['e2e_f16']['value']), 'two_groups', d['e2e_two_groups_async'] and round(d['e2e_two_groups_async']['value']))
for k,v in (d.get('configs') or {}).items():
    print(k, round(v['value']), v['ms_per_step'], v['kernel_ms'], 'roof', v['roofline']['kernel'], round(v['roofline']['frac'],4), 'e2e', round(v['e2e']['value']), v.get('terminations_per_s'))
print('cpu', d.get('cpu_baseline'))
print('closed', {k:round(v['value']) for k,v in (d.get('closed_loop') or {}).items()}, 'multi', d['multi_agent'] and round(d['multi_agent']['value']))
PY
tail -3 $OUT/bench*.err 2>/dev/null | tail -3
