#!/bin/bash
# Round 2, call af (1 GPU): chooser prefers CTA pairs on ties: what it now picks for the usual shapes, the pair vs single choice on the
# one other shape it changes (60000 x 1200), full suite, the driver's invocation.
mkdir -p gpurun_out
T="timeout -k 5"
for shape in "60000 2400 float64" "60000 1200 float64" "40000 500 float64" "11250 2000 float64" "200000 1000 float32" "125000 10000 float32" "10000 100 float64"; do
  $T 90 python profiles/one_config.py 0 0 0 0 60 $shape 2>&1 | tail -1
done | tee gpurun_out/r2af_auto_choice.txt
for cfg in "0 8 1 4" "0 16 2 4"; do
  $T 90 python profiles/one_config.py $cfg 60 60000 1200 float64 2>&1 | tail -1
done | tee -a gpurun_out/r2af_auto_choice.txt
$T 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 | tee gpurun_out/r2af_gpu_tests.log
$T 600 python bench.py --gpus 1 --steps 20 --warmup 5 2> gpurun_out/r2af_bench_driver.err > gpurun_out/r2af_bench_driver.json
python - <<'PY'
import json
d = json.load(open('gpurun_out/r2af_bench_driver.json'))
print('value', round(d['value'], 1), 'ms', round(d['ms_per_step'], 5), 'e2e_ms', round(d['e2e']['ms_per_step'], 5), 'frac', round(d['roofline']['frac'], 4), 'kernel_ms', round(d['roofline']['kernel_ms'], 5), d['clocks'], 'hier', round(d['fg_hier']['ms_per_eval'], 5), 'cpu', d['cpu_baseline']['value'], 'config5', d['config5']['ms_per_eval'], {k: d['config'][k] for k in ('tile_bins', 'cluster', 'n_clusters', 'ring_slots', 'l2_resident_mb')}, d.get('parity'))
PY
