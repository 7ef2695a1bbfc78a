#!/bin/bash
# NOTE: kept for the record -- the RD_HOST_PIPE / RD_HOST_TRACE / RD_HOST_ACT_COPY switches these runs used were removed
# together with the rejected schedules (profiles/r2p_host_pipeline_trace.txt); the script no longer runs as is.
# host-facing step: "flags" schedule (one ray-cast launch, chunk counters + cuStreamWaitValue32) against "streams".
# usage: bash tools/gpu_e2e3.sh tag
TAG=${1:-e2e3}; OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 900 python -m pytest tests -x -q -m gpu -k "host or compat or episodes or smoke or lidar or two_groups" > $OUT/pytest_host.log 2>&1; echo "pytest(host) rc=$?"; tail -3 $OUT/pytest_host.log
RD_HOST_PIPE=streams timeout 600 python -m pytest tests -x -q -m gpu -k "host or two_groups" > $OUT/pytest_host_streams.log 2>&1; echo "pytest(host, streams) rc=$?"; tail -1 $OUT/pytest_host_streams.log
B="--steps 30 --warmup 5 --no-closed-loop --no-multi-agent --no-configs --no-cpu-baseline --no-e2e-variants --e2e-steps 60"
for pipe in flags streams; do
  echo "== trace $pipe"; RD_HOST_PIPE=$pipe RD_HOST_TRACE=40 python bench.py $B 2>&1 >/dev/null | grep trace
done
for ch in "1,4,16,43" "1,3,9,27" "1,2,6,18,37" "2,6,18,38" "1,4,12,24,23"; do
  echo "== trace flags chunks $ch"; RD_HOST_CHUNKS=$ch RD_HOST_TRACE=40 python bench.py $B 2>&1 >/dev/null | grep trace
done
B="--steps 100 --warmup 10 --no-closed-loop --no-multi-agent --no-configs --no-cpu-baseline --e2e-steps 200"
show() { python -c "import json; d=json.load(open('$1')); e=d['e2e']; print('$2', 'e2e', round(e['value']), e['ms_per_step'], 'floor', round(e['d2h_floor_ms'],4), 'ratio', round(e['d2h_floor_ms']/e['ms_per_step'],3), 'f16', d['e2e_f16'] and d['e2e_f16']['ms_per_step'], 'two', d['e2e_two_groups_async'] and d['e2e_two_groups_async']['ms_per_step'], 'dev', d['ms_per_step'])"; }
for rep in 1 2; do
  python bench.py $B > $OUT/b_flags.json 2>$OUT/err.log; show $OUT/b_flags.json flags
  RD_HOST_PIPE=streams python bench.py $B > $OUT/b_streams.json 2>>$OUT/err.log; show $OUT/b_streams.json streams
  RD_HOST_PIPE=streams python bench.py $B --e2e-shards 4 > $OUT/b_streams4.json 2>>$OUT/err.log; show $OUT/b_streams4.json streams4
done
tail -3 $OUT/err.log
