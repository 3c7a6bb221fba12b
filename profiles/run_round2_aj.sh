#!/bin/bash
# Round 2, call aj (1 GPU): ncu evidence for the closing default (config 3 on 8-bin tiles / CTA pairs): --set full capture of the stream
# kernel, launch list of the bench command, steady-state DRAM bytes (application replay, caches untouched).
mkdir -p gpurun_out
T="timeout -k 5"
$T 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_final2_launches.csv python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-config5 > gpurun_out/r2aj_ncu_launch.log 2>&1
$T 600 ncu --set full --clock-control none --import-source on -k regex:sfh_fg_fused2 -s 3 -c 1 -o gpurun_out/r2_final2_config3 python profiles/one_config.py 0 0 0 0 5 > gpurun_out/r2aj_ncu1.log 2>&1
$T 300 ncu --replay-mode application --cache-control none --clock-control none --metrics gpu__time_duration.sum,dram__bytes_read.sum,lts__t_sector_hit_rate.pct -k regex:sfh_fg_fused2 -s 20 -c 3 --csv --log-file gpurun_out/r2_final2_warm.csv python profiles/one_config.py 0 0 0 0 30 > gpurun_out/r2aj_ncu2.log 2>&1
python - <<'PY'
import csv
rows = [r for r in csv.reader(open('gpurun_out/r2_final2_warm.csv')) if len(r) > 5]
h = rows[0]; im, iv = h.index("Metric Name"), h.index("Metric Value")
print('warm', [(r[im][:24], r[iv]) for r in rows[1:]])
PY
ls -la gpurun_out | tail -6
