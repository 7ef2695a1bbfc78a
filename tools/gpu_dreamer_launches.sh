#!/bin/bash
# per-kernel durations of one Dreamer agent step (ncu launch list; cold-cache, serialised: shares, not absolutes)
OUT=gpurun_out/${1:-dr}; mkdir -p $OUT
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -s 400 -c 60 --csv --log-file $OUT/launches_dreamer.csv \
   python tools/dreamer_precision_probe.py > $OUT/probe.log 2>&1; echo "rc=$?"
python - $OUT/launches_dreamer.csv <<'PY'
import csv,sys,collections
rows=list(csv.reader(open(sys.argv[1])))
hdr=None; agg=collections.OrderedDict()
for r in rows:
    if 'Kernel Name' in r: hdr=r; continue
    if hdr and len(r)==len(hdr):
        d=dict(zip(hdr,r))
        if d.get('Metric Name')=='gpu__time_duration.sum':
            k=(d['Kernel Name'][:58], d['Grid Size'])
            agg.setdefault(k,[]).append(float(d['Metric Value'])/1000.0)
for k,v in agg.items(): print('%-60s %-14s n=%2d  %.1f us'%(k[0],k[1],len(v),sum(v)/len(v)))
PY
