#!/bin/bash
# k_dense_chain launched as thread-block clusters (the N tiles of a row block = one cluster) against the plain launch
OUT=gpurun_out/${1:-chaincl}; mkdir -p $OUT
RD_DREAMER_CLUSTER=1 RD_DREAMER_DEBUG=1 timeout 900 python -m pytest tests/test_gpu_dreamer.py tests/test_gpu_policy.py -m gpu -x -q -s 2>&1 | grep -v "^$" | tail -8 | tee $OUT/pytest_cluster.log
for c in 1 0 1 0; do
  echo "RD_DREAMER_CLUSTER=$c"; RD_DREAMER_CLUSTER=$c RD_DREAMER_DEBUG=1 timeout 300 python tools/dreamer_precision_probe.py 2>&1 | grep "tf32x3\|clusters"
done | tee $OUT/probe.txt
