#!/bin/bash
# Round 2, call z (1 GPU): DRAM bytes per launch in steady state (application replay: no save / restore between passes, caches untouched)
mkdir -p gpurun_out
T="timeout -k 5"
for keep in default 0; do
  if [ $keep = default ]; then unset SFH_L2_KEEP_MB; else export SFH_L2_KEEP_MB=$keep; fi
  $T 300 ncu --replay-mode application --cache-control none --clock-control none --metrics gpu__time_duration.sum,dram__bytes_read.sum,lts__t_sector_hit_rate.pct -k regex:sfh_fg_fused2 -s 20 -c 4 --csv --log-file gpurun_out/r2z_warm_keep_$keep.csv python profiles/one_config.py 0 0 0 4 30 > gpurun_out/r2z_ncu_$keep.log 2>&1
  $T 300 ncu --cache-control none --clock-control none --metrics dram__bytes_read.sum -k regex:sfh_fg_fused2 -s 20 -c 4 --csv --log-file gpurun_out/r2z_warm1_keep_$keep.csv python profiles/one_config.py 0 0 0 4 30 > gpurun_out/r2z_ncu1_$keep.log 2>&1
done
python - <<'PY'
import csv
for pre in ('warm', 'warm1'):
    for k in ('default', '0'):
        try:
            rows = [r for r in csv.reader(open(f'gpurun_out/r2z_{pre}_keep_{k}.csv')) if len(r) > 5]
            h = rows[0]; im, iv = h.index("Metric Name"), h.index("Metric Value")
            print(pre, k, [(r[im], r[iv]) for r in rows[1:]])
        except Exception as e:
            print(pre, k, 'FAILED', e)
PY
tail -3 gpurun_out/r2z_ncu_default.log
