#!/bin/bash
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,launch__registers_per_thread --clock-control none --csv --log-file gpurun_out/r2v_fused_launches.csv python tools/prof_fused.py > gpurun_out/r2v.log 2>&1
python - <<'PY'
import csv
rows=list(csv.reader(open('gpurun_out/r2v_fused_launches.csv')))
hdr=None
for r in rows:
    if r and r[0]=='ID': hdr=r; continue
    if hdr and len(r)==len(hdr):
        d=dict(zip(hdr,r))
        if d['Metric Name'] in ('gpu__time_duration.sum','smsp__inst_executed.sum'):
            print(d['ID'], d['Kernel Name'][:60], d['Grid Size'], d['Metric Name'], d['Metric Value'])
PY
