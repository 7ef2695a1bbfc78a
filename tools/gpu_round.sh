#!/bin/bash
# One gpurun call: GPU parity tests, smoke, a bench line, the ncu launch list and one full capture per hot kernel.
# Usage (from the repo root, on the GPU box): bash tools/gpu_round.sh [tag]
TAG=${1:-r01}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > $OUT/gpu.csv 2>&1
echo "== pytest -m gpu" ; timeout 900 python -m pytest tests -x -q -m gpu > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 $OUT/pytest_gpu.log
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; echo "smoke rc=$?"; tail -3 $OUT/smoke.log
echo "== bench"; timeout 900 python bench.py > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?"; cat $OUT/bench.json; tail -3 $OUT/bench.err
echo "== bench reference"; timeout 600 python bench.py --impl reference --steps 20 --warmup 3 > $OUT/bench_ref.json 2> $OUT/bench_ref.err; echo "rc=$?"; cat $OUT/bench_ref.json
echo "== bench occupancy (config 3 shape, 16384 envs)"; timeout 900 python bench.py --obs lidar_occupancy --envs 16384 --steps 50 --warmup 5 --no-cpu-baseline > $OUT/bench_occ.json 2> $OUT/bench_occ.err; echo "rc=$?"; cat $OUT/bench_occ.json; tail -3 $OUT/bench_occ.err
echo "== ncu launch list"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches.csv \
   python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-closed-loop --no-multi-agent --e2e-steps 5 > $OUT/ncu_launches.log 2>&1; echo "rc=$?"
echo "== ncu full k_lidar"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_lidar -s 5 -c 2 -o $OUT/prof_lidar -f \
   python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-closed-loop --no-multi-agent --e2e-steps 2 > $OUT/ncu_lidar.log 2>&1; echo "rc=$?"
echo "== ncu full k_step"
timeout 900 ncu --set full --clock-control none --import-source on -k k_step -s 5 -c 2 -o $OUT/prof_step -f \
   python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-closed-loop --no-multi-agent --e2e-steps 2 > $OUT/ncu_step.log 2>&1; echo "rc=$?"
echo "== ncu full k_step_ma (worlds of 4 cars)"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_step_ma -s 5 -c 1 -o $OUT/prof_step_ma -f \
   python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-closed-loop --e2e-steps 2 > $OUT/ncu_step_ma.log 2>&1; echo "rc=$?"
echo "== ncu full k_occupancy"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_occupancy -s 2 -c 1 -o $OUT/prof_occ -f \
   python bench.py --obs lidar_occupancy --envs 4096 --steps 4 --warmup 3 --no-cpu-baseline --no-closed-loop --no-multi-agent --e2e-steps 2 > $OUT/ncu_occ.log 2>&1; echo "rc=$?"
ls -la $OUT
