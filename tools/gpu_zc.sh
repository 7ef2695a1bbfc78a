#!/bin/bash
OUT=gpurun_out/${1:-zc}; mkdir -p $OUT
for zc in 0 1; do for sh in 1 4; do RD_HOST_ZEROCOPY=$zc python bench.py --no-cpu-baseline --e2e-shards $sh --steps 300 --warmup 50 2>$OUT/err.log | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('zerocopy $zc shards', d['e2e']['shards'], 'e2e', d['e2e']['value'], d['e2e']['ms_per_step'], 'floor', d['e2e']['d2h_floor_ms'])"; tail -2 $OUT/err.log; done; done
RD_HOST_ZEROCOPY=1 timeout 600 python -m pytest tests -x -q -m gpu -k host_stepped 2>&1 | tail -2
python tools/lidar_sweep.py 2>&1 | grep -E "barcelona|columbia" 
