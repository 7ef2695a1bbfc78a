#!/bin/bash
# k_lidar work-chunk size (variants/chunk*.so) and draw-ahead A/B.  usage: bash tools/gpu_lidar_chunk.sh tag
TAG=${1:-lc}; OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "split_step" > $OUT/pytest_split.log 2>&1; echo "pytest(split) rc=$?"; tail -2 $OUT/pytest_split.log
for rep in 1 2; do
  for spec in austria:4096:1 austria:16384:1 treitlstrasse_v2:65536:1 columbia:16384:1; do
    echo "== base ahead=auto $spec" | tee -a $OUT/ab.log; RD_SWEEP=$spec python tools/lidar_sweep.py 2>&1 | tail -1 | tee -a $OUT/ab.log
    for a in 0 1; do
      echo "== base ahead=$a $spec" | tee -a $OUT/ab.log; RD_LIDAR_AHEAD=$a RD_SWEEP=$spec python tools/lidar_sweep.py 2>&1 | tail -1 | tee -a $OUT/ab.log
    done
    for v in variants/chunk*.so; do
      echo "== $(basename $v .so) $spec" | tee -a $OUT/ab.log; RD_ENV_LIB=$PWD/$v RD_SWEEP=$spec python tools/lidar_sweep.py 2>&1 | tail -1 | tee -a $OUT/ab.log
    done
  done
done
