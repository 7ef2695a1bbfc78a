#!/bin/bash
# k_dense_chain with the fused output head (RD_DREAMER_HEAD=1) against the head as a k_dense launch of its own
OUT=gpurun_out/${1:-head}; mkdir -p $OUT
RD_DREAMER_HEAD=1 timeout 900 python -m pytest tests/test_gpu_dreamer.py tests/test_gpu_policy.py -m gpu -q -k "not bitwise" 2>&1 | tail -15 | tee $OUT/pytest_head.log
for h in 1 0 1 0; do
  echo "RD_DREAMER_HEAD=$h"; RD_DREAMER_HEAD=$h timeout 300 python tools/dreamer_precision_probe.py 2>&1 | grep "tf32x3"
done | tee $OUT/probe.txt
