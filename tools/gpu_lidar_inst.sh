#!/bin/bash
# warp instructions and duration of the k_lidar launches of one bench run (config 2): shipped library against a variant
# (usage: bash tools/gpu_lidar_inst.sh tag [variants/lib.so])
OUT=gpurun_out/${1:-linst}; mkdir -p $OUT
run() {  # label, env assignments...
  local label=$1; shift
  env "$@" timeout 600 ncu --metrics smsp__inst_executed.sum,gpu__time_duration.sum --clock-control none -k regex:k_lidar -s 6 -c 3 --csv --log-file $OUT/inst_$label.csv \
    python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-closed-loop --no-multi-agent --no-configs --no-e2e-variants --e2e-steps 2 > /dev/null 2>&1
  echo "$label"; python - $OUT/inst_$label.csv <<'PY'
import csv,sys
rows=list(csv.reader(open(sys.argv[1])))
hdr=[r for r in rows if 'Metric Name' in r][0]
for r in rows:
    if len(r)==len(hdr) and r!=hdr:
        d=dict(zip(hdr,r)); print('  ', d['Metric Name'], d['Metric Value'], d['Grid Size'])
PY
}
(run shipped RD_NOTHING=0; if [ -n "$2" ]; then run variant RD_ENV_LIB=$2; fi) | tee $OUT/inst.txt
