#!/bin/bash
# k_lidar CTA shape A/B (RD_LIDAR_WARPS) on the config-2 / config-3 / config-4 shapes.  usage: bash tools/gpu_lidar_warps.sh tag
TAG=${1:-lw}; OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 600 python -m pytest tests -x -q -m gpu -k "lidar or parity or smoke or fuzz" > $OUT/pytest.log 2>&1; echo "pytest rc=$?"; tail -2 $OUT/pytest.log
for rep in 1 2; do
  for spec in austria:4096:1 austria:16384:1 treitlstrasse_v2:65536:1 columbia:16384:1 barcelona:16384:2; do
    for w in auto 16 24 32; do
      echo "== warps=$w $spec" | tee -a $OUT/ab.log
      if [ $w = auto ]; then RD_SWEEP=$spec python tools/lidar_sweep.py 2>&1 | tail -1 | tee -a $OUT/ab.log
      else RD_LIDAR_WARPS=$w RD_SWEEP=$spec python tools/lidar_sweep.py 2>&1 | tail -1 | tee -a $OUT/ab.log; fi
    done
  done
done
