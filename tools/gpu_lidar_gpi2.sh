#!/bin/bash
# threshold of two-groups-per-item: forced 1 / 2 at intermediate batch sizes.  usage: bash tools/gpu_lidar_gpi2.sh tag
TAG=${1:-gpi2}; OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 600 python -m pytest tests -x -q -m gpu -k "lidar or closed_loop or mixed_maps or host or multi" > $OUT/pytest.log 2>&1; echo "pytest rc=$?"; tail -1 $OUT/pytest.log
RD_LIDAR_GPI=2 timeout 600 python -m pytest tests -x -q -m gpu -k "lidar or closed_loop or mixed_maps or host or multi" > $OUT/pytest2.log 2>&1; echo "pytest(gpi=2) rc=$?"; tail -1 $OUT/pytest2.log
for spec in austria:4096:1 austria:6144:1 austria:8192:1 austria:12288:1 treitlstrasse_v2:8192:1 treitlstrasse_v2:16384:1 columbia:4096:1 columbia:8192:1; do
  for g in 1 2; do
    echo "== gpi=$g $spec" | tee -a $OUT/ab.log; RD_LIDAR_GPI=$g RD_SWEEP=$spec python tools/lidar_sweep.py 2>&1 | tail -1 | tee -a $OUT/ab.log
  done
done
