#!/bin/bash
# actor trunk in one launch (k_dense_chain) and copy-engine output tiles against the per-layer / register-store paths:
# tests, then the agent-step time each way
OUT=gpurun_out/${1:-chain}; mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_dreamer.py tests/test_gpu_policy.py -m gpu -x -q > $OUT/pytest_dreamer.log 2>&1; echo "pytest rc=$?"; tail -5 $OUT/pytest_dreamer.log
for cfg in "1 1" "1 0" "0 1" "0 0" "1 1"; do
  set -- $cfg
  echo "RD_DREAMER_CHAIN=$1 RD_DREAMER_TMA_OUT=$2"; RD_DREAMER_CHAIN=$1 RD_DREAMER_TMA_OUT=$2 timeout 300 python tools/dreamer_precision_probe.py 2>&1 | grep -v "^$" | grep "n=4096\|tf32x3"
done | tee $OUT/probe.txt
