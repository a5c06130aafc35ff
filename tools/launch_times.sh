#!/bin/bash
# usage: tools/launch_times.sh <tag> <log2_build> <log2_probe>   (runs on the GPU box; per-kernel durations via ncu)
set -e
SWEEP_ONLY=16:2 python tools/sweep_probe.py $2 $3 > gpurun_out/p.log 2>&1
SWEEP_ONLY=16:2 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/launches_$1.csv python tools/sweep_probe.py $2 $3 > gpurun_out/ncu_$1.log 2>&1
python - <<PY
import csv
from collections import OrderedDict
rows=[r for r in csv.reader(open('gpurun_out/launches_$1.csv')) if len(r)>10 and r[0].isdigit()]
d=OrderedDict()
for r in rows: d.setdefault((r[0],r[4][:44]),{})[r[-3]]=r[-1]
for k,v in list(d.items())[-5:]: print(f"{k[1]:46s} {float(v['gpu__time_duration.sum'])/1e6:8.3f} ms  inst {float(v['smsp__inst_executed.sum'])/1e9:6.3f} G  dram R {float(v['dram__bytes_read.sum'])/1e9:6.2f} W {float(v['dram__bytes_write.sum'])/1e9:6.2f} GB")
PY
