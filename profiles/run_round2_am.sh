#!/bin/bash
# Round 2, call am (1 GPU): the closing build as the driver will see it: full suite, smoke(), both bench arms in the driver's invocation.
mkdir -p gpurun_out
T="timeout -k 5"
$T 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 | tee gpurun_out/r2am_gpu_tests.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 | tee gpurun_out/r2am_smoke.txt
$T 600 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 2> gpurun_out/r2am_bench_reference.err > gpurun_out/r2am_bench_reference.json
$T 600 python bench.py --gpus 1 --steps 20 --warmup 5 2> gpurun_out/r2am_bench_driver.err > gpurun_out/r2am_bench_driver.json
python - <<'PY'
import json
d = json.load(open('gpurun_out/r2am_bench_driver.json'))
print('value', round(d['value'], 1), 'ms', round(d['ms_per_step'], 5), 'e2e', round(d['e2e']['value'], 1), 'e2e_ms', round(d['e2e']['ms_per_step'], 5), 'frac', round(d['roofline']['frac'], 4), 'kernel_ms', round(d['roofline']['kernel_ms'], 5), d['clocks'], 'hier', round(d['fg_hier']['ms_per_eval'], 5), 'cpu', round(d['cpu_baseline']['value'], 1), 'config5', round(d['config5']['ms_per_eval'], 3), d.get('parity'))
r = json.load(open('gpurun_out/r2am_bench_reference.json'))
print('reference', round(r['value'], 1), r['cpu_baseline']['cores'], r['cpu_baseline']['kind'], 'e2e ratio', round(d['e2e']['value'] / r['value'], 1), 'device ratio', round(d['value'] / r['value'], 1))
PY
