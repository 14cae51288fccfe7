#!/bin/bash
mkdir -p gpurun_out
python scripts/ncu_step.py cnn 128 8 > /dev/null 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none --csv --log-file gpurun_out/r02_launches_warm.csv python scripts/ncu_step.py cnn 128 8 > gpurun_out/ncu.log 2>&1
tail -3 gpurun_out/ncu.log
python - <<'P'
import csv
rows=list(csv.reader(open('gpurun_out/r02_launches_warm.csv')))
hdr=None; out=[]
for r in rows:
    if len(r)>5 and r[0]=='ID': hdr=r; continue
    if hdr and len(r)==len(hdr): out.append(dict(zip(hdr,r)))
n=len(out)//8
last=out[-n:]
tot=0
for d in last:
    v=float(d['Metric Value'].replace(',',''));
    if d['Metric Unit']=='ns': v/=1e3
    tot+=v
    print(f"{v:7.2f} us  {d['Kernel Name'][:70]:70s} grid {d['Grid Size']}")
print('kernels per step',n,'sum',round(tot,1))
P
