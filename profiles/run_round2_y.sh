#!/bin/bash
# Round 2, call y (1 GPU): L2-resident head as the default (budget min(3/4 L2, 8 % of the stack)): full suite, DRAM bytes per launch with
# warm caches (ncu --cache-control none) with and without it, the driver's invocation.
mkdir -p gpurun_out
T="timeout -k 5"
$T 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 | tee gpurun_out/r2y_gpu_tests.log
for keep in default 0; do
  if [ $keep = default ]; then unset SFH_L2_KEEP_MB; else export SFH_L2_KEEP_MB=$keep; fi
  $T 300 ncu --cache-control none --clock-control none --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,lts__t_bytes.sum -k regex:sfh_fg_fused2 -s 20 -c 4 --csv --log-file gpurun_out/r2y_warm_keep_$keep.csv python profiles/one_config.py 0 0 0 4 30 > gpurun_out/r2y_ncu_$keep.log 2>&1
done
unset SFH_L2_KEEP_MB
python - <<'PY'
import csv
for k in ('default', '0'):
    rows = [r for r in csv.reader(open(f'gpurun_out/r2y_warm_keep_{k}.csv')) if len(r) > 5]
    h = rows[0]; im, iv, iu = h.index("Metric Name"), h.index("Metric Value"), h.index("Metric Unit")
    print(k, [(r[im], r[iv], r[iu]) for r in rows[1:6]])
PY
$T 600 python bench.py --gpus 1 --steps 20 --warmup 5 2> gpurun_out/r2y_bench_driver.err > gpurun_out/r2y_bench_driver.json
python - <<'PY'
import json
d = json.load(open('gpurun_out/r2y_bench_driver.json'))
print('value', round(d['value'], 1), 'ms', round(d['ms_per_step'], 5), 'e2e_ms', round(d['e2e']['ms_per_step'], 5), 'frac', round(d['roofline']['frac'], 4), 'kernel_ms', round(d['roofline']['kernel_ms'], 5), d['clocks'], 'hier', round(d['fg_hier']['ms_per_eval'], 5), 'cpu', d['cpu_baseline']['value'], 'config5', d['config5']['ms_per_eval'], d['config']['l2_resident_mb'], d.get('parity'))
PY
