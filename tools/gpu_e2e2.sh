#!/bin/bash
# NOTE: kept for the record -- the RD_HOST_PIPE / RD_HOST_TRACE / RD_HOST_ACT_COPY switches these runs used were removed
# together with the rejected schedules (profiles/r2p_host_pipeline_trace.txt); the script no longer runs as is.
# host-facing step A/B: default vs explicit action copy, shard counts.  usage: bash tools/gpu_e2e2.sh tag
TAG=${1:-e2e}; OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 900 python -m pytest tests -x -q -m gpu -k "host or compat or episodes or smoke" > $OUT/pytest_host.log 2>&1; echo "pytest(host) rc=$?"; tail -3 $OUT/pytest_host.log
B="--steps 100 --warmup 10 --no-closed-loop --no-multi-agent --no-configs --no-cpu-baseline --e2e-steps 200"
show() { python -c "import json; d=json.load(open('$1')); e=d['e2e']; print('$2', 'e2e', round(e['value']), e['ms_per_step'], 'floor', round(e['d2h_floor_ms'],4), 'ratio', round(e['d2h_floor_ms']/e['ms_per_step'],3), 'f16', d['e2e_f16'] and d['e2e_f16']['ms_per_step'], 'two', d['e2e_two_groups_async'] and d['e2e_two_groups_async']['ms_per_step'])"; }
for rep in 1 2; do
  python bench.py $B > $OUT/b_default.json 2>$OUT/err.log; show $OUT/b_default.json default
  RD_HOST_ACT_COPY=1 python bench.py $B > $OUT/b_actcopy.json 2>>$OUT/err.log; show $OUT/b_actcopy.json act_copy
  python bench.py $B --e2e-shards 4 > $OUT/b_sh4.json 2>>$OUT/err.log; show $OUT/b_sh4.json shards4
  python bench.py $B --e2e-shards 12 > $OUT/b_sh12.json 2>>$OUT/err.log; show $OUT/b_sh12.json shards12
  RD_HOST_ZEROCOPY=1 python bench.py $B > $OUT/b_zc.json 2>>$OUT/err.log; show $OUT/b_zc.json zerocopy
done
tail -3 $OUT/err.log
