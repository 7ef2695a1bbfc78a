#!/bin/bash
# ncu --set full captures of k_lidar (config 2) and k_occupancy (config 3 at 4096 envs).  usage: bash tools/gpu_ncu2.sh tag
TAG=$1; OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_lidar -s 6 -c 1 -o $OUT/prof_k_lidar -f \
   python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-closed-loop --no-multi-agent --no-configs --e2e-steps 2 > $OUT/ncu_k_lidar.log 2>&1; echo "ncu k_lidar rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_occupancy -s 2 -c 1 -o $OUT/prof_k_occ -f \
   python bench.py --config 3 --envs 4096 --steps 4 --warmup 3 --no-cpu-baseline --no-closed-loop --no-multi-agent --no-configs --e2e-steps 2 > $OUT/ncu_k_occ.log 2>&1; echo "ncu k_occ rc=$?"
