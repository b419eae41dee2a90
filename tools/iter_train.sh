#!/bin/bash
# development iteration of the reverse pass: gradient tests, step breakdown, warm per-kernel launch list (T=4)
tag=${1:-x}
python -m pytest tests/test_gpu_train.py -m gpu -x -q 2>&1 | tail -3
python tools/train_breakdown.py 32 2>&1 | tail -5
ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none --csv --log-file gpurun_out/train_launches_${tag}.csv python tools/train_profile.py 4 > /dev/null 2>&1
python tools/ncu_by_kernel.py gpurun_out/train_launches_${tag}.csv > gpurun_out/train_by_kernel_${tag}.txt
python - <<P
import csv
from collections import defaultdict
rows=[r for r in csv.reader(open('gpurun_out/train_launches_${tag}.csv')) if len(r)>5]
hdr=next(i for i,r in enumerate(rows) if "Kernel Name" in r)
H=rows[hdr]; ki,vi,gi=H.index("Kernel Name"),H.index("Metric Value"),H.index("Grid Size")
d=defaultdict(list)
for r in rows[hdr+1:]:
    d[(r[ki][:48], r[gi])].append(float(r[vi].replace(",",""))/1e3)
for k,v in sorted(d.items(), key=lambda kv:-sum(kv[1])):
    if sum(v)>150: print("  %-50s %-14s n=%3d avg %7.1f min %7.1f"%(k[0],k[1],len(v),sum(v)/len(v),min(v)))
P
