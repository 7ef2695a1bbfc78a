#!/bin/bash
# k_dense with / without the activation multicast across a cluster of N tiles.  usage: bash tools/gpu_dense_cluster.sh tag
OUT=gpurun_out/${1:-dc}; mkdir -p $OUT
timeout 300 python -m pytest tests/test_gpu_dreamer.py tests/test_gpu_policy.py -x -q -m gpu > $OUT/pytest_dreamer.log 2>&1; echo "pytest(cluster) rc=$?"; tail -3 $OUT/pytest_dreamer.log
for c in 1 0 1 0; do echo "== RD_DENSE_CLUSTER=$c"; RD_DENSE_CLUSTER=$c timeout 200 python tools/dreamer_precision_probe.py 2>&1 | tail -4; done
