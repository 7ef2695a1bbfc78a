#!/bin/bash
# k_step_split (two warps per 32 envs) against k_step.  usage: bash tools/gpu_step_split.sh tag
TAG=${1:-ss}; OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 900 python -m pytest tests -x -q -m gpu > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 $OUT/pytest_gpu.log
B="--steps 200 --warmup 20 --no-closed-loop --no-multi-agent --no-configs --no-cpu-baseline --no-e2e-variants --e2e-steps 100"
show() { python -c "import json,sys; d=json.loads(open('$1').read()); print('$2', round(d['value']), d['ms_per_step'], d['kernel_ms'], 'e2e', d['e2e']['ms_per_step'])"; }
for rep in 1 2; do
  for sp in 1 0; do
    RD_STEP_SPLIT=$sp python bench.py $B > $OUT/b_$sp.json 2>$OUT/err.log; show $OUT/b_$sp.json "config2 split=$sp"
    RD_STEP_SPLIT=$sp python bench.py --config 4 $B > $OUT/b4_$sp.json 2>>$OUT/err.log; show $OUT/b4_$sp.json "config4 split=$sp"
    RD_STEP_SPLIT=$sp python bench.py --envs 16384 $B > $OUT/b16_$sp.json 2>>$OUT/err.log; show $OUT/b16_$sp.json "16384 envs split=$sp"
    RD_STEP_SPLIT=$sp python bench.py --envs 8192 $B > $OUT/b8_$sp.json 2>>$OUT/err.log; show $OUT/b8_$sp.json "8192 envs split=$sp"
  done
done
tail -3 $OUT/err.log
