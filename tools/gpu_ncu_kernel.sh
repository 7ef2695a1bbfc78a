#!/bin/bash
# one `ncu --set full` capture per kernel regex from the config-2 bench.  usage: bash tools/gpu_ncu_kernel.sh tag regex [regex...]
TAG=$1; shift; OUT=gpurun_out/$TAG; mkdir -p $OUT
for K in "$@"; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$K -s 6 -c 1 -o $OUT/prof_$K -f \
     python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-closed-loop --no-multi-agent --e2e-steps 2 > $OUT/ncu_$K.log 2>&1; echo "ncu $K rc=$?"
done
