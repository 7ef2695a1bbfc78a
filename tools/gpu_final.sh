#!/bin/bash
# Round-end validation in one gpurun call: GPU tests, smoke, both bench arms with the driver's flags, the ncu launch
# list of the same command, one full capture per hot kernel.  usage: bash tools/gpu_final.sh tag
TAG=${1:-final}; OUT=gpurun_out/$TAG; mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > $OUT/gpu.csv 2>&1
echo "== pytest -m gpu"; timeout 1200 python -m pytest tests -x -q -m gpu > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 $OUT/pytest_gpu.log
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 $OUT/smoke.log
echo "== bench reference arm"; timeout 600 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > $OUT/bench_ref.json 2> $OUT/bench_ref.err; echo "rc=$?"
echo "== bench"; timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?"; tail -2 $OUT/bench.err
python - $OUT/bench.json <<'PY'
import json,sys
d=json.load(open(sys.argv[1]))
print('config2', round(d['value']), d['ms_per_step'], d['kernel_ms'], 'e2e', round(d['e2e']['value']), d['e2e']['ms_per_step'], 'floor', d['e2e']['d2h_floor_ms'])
print('e2e_f16', d['e2e_f16'] and round(d['e2e_f16']['value']), 'two_groups', d['e2e_two_groups_async'] and round(d['e2e_two_groups_async']['value']))
for k,v in (d.get('configs') or {}).items():
    print(k, round(v['value']), v['ms_per_step'], v['kernel_ms'], 'e2e', round(v['e2e']['value']), v.get('terminations_per_s'))
print('closed', {k: round(v['value']) for k,v in (d.get('closed_loop') or {}).items() if isinstance(v, dict) and 'value' in v})
PY
echo "== ncu launch list"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches.csv \
   python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline --no-closed-loop --no-multi-agent --no-configs --e2e-steps 5 > $OUT/ncu_launches.log 2>&1; echo "rc=$?"
for K in k_lidar k_step_split; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$K -s 6 -c 1 -o $OUT/prof_$K -f \
     python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-closed-loop --no-multi-agent --no-configs --e2e-steps 2 > $OUT/ncu_$K.log 2>&1; echo "ncu $K rc=$?"
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_occupancy -s 2 -c 1 -o $OUT/prof_k_occupancy -f \
   python bench.py --config 3 --envs 4096 --steps 4 --warmup 3 --no-cpu-baseline --no-closed-loop --no-multi-agent --no-configs --e2e-steps 2 > $OUT/ncu_k_occupancy.log 2>&1; echo "ncu k_occupancy rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_step_ma -s 5 -c 1 -o $OUT/prof_k_step_ma -f \
   python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-closed-loop --no-configs --e2e-steps 2 > $OUT/ncu_k_step_ma.log 2>&1; echo "ncu k_step_ma rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_dense_chain -s 5 -c 1 -o $OUT/prof_k_dense_chain -f \
   python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-multi-agent --no-configs --no-e2e-variants --e2e-steps 2 > $OUT/ncu_k_dense_chain.log 2>&1; echo "ncu k_dense_chain rc=$?"
echo "== Dreamer agent step: per-kernel launch list and step time"
bash tools/gpu_dreamer_launches.sh $TAG > $OUT/dreamer_launches.txt 2>&1; tail -12 $OUT/dreamer_launches.txt
timeout 300 python tools/dreamer_precision_probe.py > $OUT/dreamer_probe.txt 2>&1; cat $OUT/dreamer_probe.txt
ls -la $OUT | head -40
