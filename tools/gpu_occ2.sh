#!/bin/bash
# k_occupancy iteration: parity tests, config-3 bench for the in-tree library and every variants/*.so, ncu of k_occupancy.
# usage: bash tools/gpu_occ2.sh tag [ncu]
TAG=${1:-occ}; OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 1500 python -m pytest tests -x -q -m gpu > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 $OUT/pytest_gpu.log
B="--config 3 --steps 30 --warmup 5 --no-closed-loop --no-multi-agent --no-configs --no-cpu-baseline --no-e2e-variants --e2e-steps 3"
show() { python -c "import json,sys; d=json.load(open('$1')); print('$2', round(d['value']), d['ms_per_step'], d['kernel_ms'])"; }
for rep in 1 2; do
  python bench.py $B > $OUT/bench_c3.json 2> $OUT/bench_c3.err; show $OUT/bench_c3.json base
  for v in variants/*.so; do
    [ -f "$v" ] || continue
    b=$(basename $v .so)
    RD_ENV_LIB=$PWD/$v python bench.py $B > $OUT/bench_c3_$b.json 2> $OUT/bench_c3_$b.err; show $OUT/bench_c3_$b.json $b
  done
done
if [ -n "$2" ]; then
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_occupancy -s 2 -c 1 -o $OUT/prof_occ -f \
     python bench.py --config 3 --envs 4096 --steps 4 --warmup 3 --no-closed-loop --no-multi-agent --no-configs --no-cpu-baseline --no-e2e-variants --e2e-steps 2 > $OUT/ncu_occ.log 2>&1; echo "ncu rc=$?"
fi
