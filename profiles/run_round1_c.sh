#!/bin/bash
mkdir -p gpurun_out
python bench.py --steps 2000 --warmup 10 > gpurun_out/bench_r1d.json 2> gpurun_out/bench_r1d.err; python -c "
import json; d=json.load(open('gpurun_out/bench_r1d.json')); print('value',d['value'],'e2e',d['e2e']['value'],'frac',d['roofline']['frac'],'kernel_ms',d['roofline']['kernel_ms'],'cfg',d['config']['tile_bins'],d['config']['cluster'],d['config']['consumer_warps'],'cpu',d['cpu_baseline']['value'],d['cpu_baseline']['cores'], d['clocks'])"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 100 --csv --log-file gpurun_out/launches_r1d.csv python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launch_d.log 2>&1
python - <<'PY'
import csv
from collections import defaultdict
rows=[r for r in csv.reader(open('gpurun_out/launches_r1d.csv')) if len(r)>5]
h=rows[0]; ik,iv=h.index('Kernel Name'),h.index('Metric Value')
d=defaultdict(list)
for r in rows[1:]:
    try: d[r[ik][:70]].append(float(r[iv].replace(',','')))
    except: pass
for k,v in d.items(): print(f"{k:72s} n={len(v):3d} mean={sum(v)/len(v)/1000:9.2f} us")
PY
