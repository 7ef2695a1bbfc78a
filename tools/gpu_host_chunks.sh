#!/bin/bash
# chunk weights of the host-facing step (RD_HOST_CHUNKS) with the round-2 kernels.  usage: bash tools/gpu_host_chunks.sh
B="--steps 30 --warmup 5 --no-closed-loop --no-multi-agent --no-configs --no-cpu-baseline --no-e2e-variants --e2e-steps 300"
for rep in 1 2; do
for ch in "" "1,2,4,8" "1,3,9,27" "1,2,4,8,16" "1,2,4,8,16,16" "1,4,16,32" "1,2,6,12,24" "2,4,8,16,16" "1,1,2,4,8,8,8"; do
  if [ -z "$ch" ]; then python bench.py $B > /tmp/b.json 2>/dev/null; else RD_HOST_CHUNKS=$ch python bench.py $B > /tmp/b.json 2>/dev/null; fi
  python -c "import json; d=json.load(open('/tmp/b.json')); e=d['e2e']; print('chunks=%-18s e2e %.4f ms  floor %.4f  ratio %.3f' % ('$ch' or 'default', e['ms_per_step'], e['d2h_floor_ms'], e['d2h_floor_ms']/e['ms_per_step']))"
done
done
