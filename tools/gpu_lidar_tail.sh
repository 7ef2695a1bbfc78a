#!/bin/bash
# k_lidar work list: the last T items per resident warp drawn one at a time (RD_LIDAR_TAIL=T, 0 = whole chunks only)
OUT=gpurun_out/${1:-tail}; mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_full_size.py -m gpu -x -q > $OUT/pytest.log 2>&1; echo "pytest rc=$?"; tail -2 $OUT/pytest.log
for rep in 1 2; do
for t in 0 1 2 4 8; do
  for c in 2 4; do
    echo -n "RD_LIDAR_TAIL=$t config $c: "
    RD_LIDAR_TAIL=$t timeout 300 python bench.py --config $c --steps 40 --warmup 5 --no-cpu-baseline --no-closed-loop --no-multi-agent --no-configs --no-e2e-variants --e2e-steps 3 2>/dev/null \
      | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(round(d['ms_per_step'],4), 'ms', {k: round(v*1e3,1) for k,v in d['kernel_ms'].items()})"
  done
done
done | tee $OUT/tail.txt
