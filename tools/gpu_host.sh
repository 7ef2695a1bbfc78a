#!/bin/bash
OUT=gpurun_out/${1:-host}; mkdir -p $OUT
timeout 900 python -m pytest tests -x -q -m gpu > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -6 $OUT/pytest_gpu.log
for sh in 2 4 6 8; do python bench.py --no-cpu-baseline --e2e-shards $sh --steps 300 --warmup 50 2>$OUT/err_$sh.log | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('shards', d['e2e']['shards'], 'e2e', d['e2e']['value'], d['e2e']['ms_per_step'], 'floor', d['e2e']['d2h_floor_ms'], 'dev', d['value'], d['kernel_ms'])"; tail -2 $OUT/err_$sh.log; done
for lib in "" variants/occ1024.so; do
python bench.py --config 3 --steps 50 --warmup 5 --no-cpu-baseline 2>$OUT/err_c3.log | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('config3', d['value'], d['kernel_ms'], 'e2e', d['e2e']['value'])"; tail -2 $OUT/err_c3.log
done
RD_ENV_LIB=$PWD/variants/occ1024.so timeout 600 python -m pytest tests -x -q -m gpu -k occupancy 2>&1 | tail -2
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_occupancy -s 2 -c 1 -o $OUT/prof_occ -f \
   python bench.py --config 3 --envs 4096 --steps 4 --warmup 3 --no-cpu-baseline --e2e-steps 2 > $OUT/ncu_occ.log 2>&1; echo "ncu rc=$?"
