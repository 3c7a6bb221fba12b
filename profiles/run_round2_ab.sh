#!/bin/bash
# Round 2, call ab (1 GPU): experiment -- persisting access-policy window (L2 set-aside) for the head of the stack vs evict_last hints
mkdir -p gpurun_out
T="timeout -k 5"
for mode in 0 1 0 1; do
  echo -n "window=$mode " ; SFH_L2_WINDOW=$mode $T 90 python profiles/one_config.py 0 0 0 4 60 2>&1 | tail -2 | tr '\n' ' '; echo
done | tee gpurun_out/r2ab_window.txt
for mode in 0 1; do
  SFH_L2_WINDOW=$mode $T 300 ncu --replay-mode application --cache-control none --clock-control none --metrics gpu__time_duration.sum,dram__bytes_read.sum,lts__t_sector_hit_rate.pct -k regex:sfh_fg_fused2 -s 20 -c 3 --csv --log-file gpurun_out/r2ab_warm_window_$mode.csv python profiles/one_config.py 0 0 0 4 30 > gpurun_out/r2ab_ncu_$mode.log 2>&1
done
python - <<'PY'
import csv
for k in ('0', '1'):
    try:
        rows = [r for r in csv.reader(open(f'gpurun_out/r2ab_warm_window_{k}.csv')) if len(r) > 5]
        h = rows[0]; im, iv = h.index("Metric Name"), h.index("Metric Value")
        print('window', k, [(r[im][:24], r[iv]) for r in rows[1:]])
    except Exception as e:
        print(k, 'FAILED', e)
PY
