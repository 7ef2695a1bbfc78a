#!/bin/bash
# two-track batches (config 5): each track's ray casting on its own stream against both on the caller's stream;
# full per-GPU size and one of eight shards of it (strong scaling).  Also the two-track parity tests.
OUT=gpurun_out/${1:-fork}; mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_full_size.py tests/test_gpu_config_fuzz.py -m gpu -x -q > $OUT/pytest.log 2>&1; echo "pytest rc=$?"; tail -2 $OUT/pytest.log
for f in 1 0 1 0; do
  for n in 16384 131072; do
    echo -n "RD_FORK_MAPS=$f envs=$n: "
    RD_FORK_MAPS=$f timeout 300 python bench.py --config 5 --envs $n --steps 30 --warmup 5 --no-cpu-baseline --no-closed-loop --no-multi-agent --no-configs --no-e2e-variants --e2e-steps 3 2>/dev/null \
      | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(round(d['ms_per_step'],4), 'ms', round(d['value']/1e6,2), 'M', d['kernel_ms'])"
  done
done | tee $OUT/fork.txt
