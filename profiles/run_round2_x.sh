#!/bin/bash
# Round 2, call x (1 GPU): L2-resident head, second sweep: intermediate stack sizes (180 MB ... 1.15 GB) x budget, then the bench line with
# the budget off / 64 / 96 MB in the driver's invocation and in a 2000-step run.
mkdir -p gpurun_out
T="timeout -k 5"
for shape in "11250 2000 float64" "100000 500 float64" "60000 1200 float64" "200000 1000 float32" "60000 2400 float64"; do
  for keep in 0 16 32 48 64 96; do
    echo -n "keep_mb=$keep " ; SFH_L2_KEEP_MB=$keep $T 90 python profiles/one_config.py 0 0 0 4 60 $shape 2>&1 | tail -1
  done
done | tee gpurun_out/r2x_l2keep_sweep.txt
for keep in 0 64 96; do
  SFH_L2_KEEP_MB=$keep $T 300 python bench.py --gpus 1 --steps 20 --warmup 5 --no-config5 --no-cpu-baseline 2> gpurun_out/r2x_bench20_keep$keep.err > gpurun_out/r2x_bench20_keep$keep.json
  SFH_L2_KEEP_MB=$keep $T 300 python bench.py --gpus 1 --steps 2000 --warmup 10 --no-config5 --no-cpu-baseline 2> gpurun_out/r2x_bench2000_keep$keep.err > gpurun_out/r2x_bench2000_keep$keep.json
done
python - <<'PY'
import json
for k in (0, 64, 96):
    for n in ('20', '2000'):
        try:
            d = json.load(open(f'gpurun_out/r2x_bench{n}_keep{k}.json'))
            print('keep', k, 'steps', n, 'value', round(d['value'], 1), 'ms', round(d['ms_per_step'], 5), 'e2e_ms', round(d['e2e']['ms_per_step'], 5), 'frac', round(d['roofline']['frac'], 4),
                  'kernel_ms', round(d['roofline']['kernel_ms'], 5), d['clocks']['sm_mhz'], 'hier', round(d['fg_hier']['ms_per_eval'], 5))
        except Exception as e:
            print(k, n, 'FAILED', e)
PY
