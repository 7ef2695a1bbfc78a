#!/bin/bash
OUT=gpurun_out/${1:-occ}; mkdir -p $OUT
timeout 900 python -m pytest tests -x -q -m gpu > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 $OUT/pytest_gpu.log
RD_ENV_LIB=$PWD/variants/occ1024.so timeout 600 python -m pytest tests -x -q -m gpu -k occupancy 2>&1 | tail -2
for lib in librd_env.so ../variants/occ1024.so; do
  echo "== lib=$lib"
  RD_ENV_LIB=$PWD/racing_dreamer_b200/$lib python bench.py --config 3 --steps 50 --warmup 5 --no-cpu-baseline 2>$OUT/err_c3.log | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('config3', d['value'], d['kernel_ms'], 'e2e', d['e2e']['value'])"; tail -2 $OUT/err_c3.log
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_occupancy -s 2 -c 1 -o $OUT/prof_occ -f \
   python bench.py --config 3 --envs 4096 --steps 4 --warmup 3 --no-cpu-baseline --e2e-steps 2 > $OUT/ncu_occ.log 2>&1; echo "ncu rc=$?"
