#!/bin/bash
# where does k_lidar's fixed cost come from?  kernel time vs batch size, with the launch attributes toggled
for n in 64 512 4096 16384; do
  for cfg in "" "RD_L2_PERSIST=0" "RD_LIDAR_PDL=0" "RD_L2_PERSIST=0 RD_LIDAR_PDL=0"; do
    echo "== n=$n $cfg"; env $cfg RD_SWEEP=austria:$n:1 python tools/lidar_sweep.py 2>&1 | tail -1 | cut -c1-120
  done
done
