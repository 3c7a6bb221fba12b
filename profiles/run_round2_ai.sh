#!/bin/bash
# Round 2, call ai (1 GPU): L2 budget re-swept on the pair tiling of config 3 (bench.py in the driver's invocation + isolated kernel)
mkdir -p gpurun_out
T="timeout -k 5"
for keep in 64 80 87 96 104; do
  echo -n "keep_mb=$keep " ; SFH_L2_KEEP_MB=$keep $T 90 python profiles/one_config.py 0 0 0 0 60 2>&1 | tail -1
done | tee gpurun_out/r2ai_keep_pairs.txt
for keep in 80 96; do
  SFH_L2_KEEP_MB=$keep $T 300 python bench.py --gpus 1 --steps 20 --warmup 5 --no-config5 --no-cpu-baseline 2> /dev/null > gpurun_out/r2ai_bench20_keep$keep.json
  SFH_L2_KEEP_MB=$keep $T 300 python bench.py --gpus 1 --steps 2000 --warmup 10 --no-config5 --no-cpu-baseline 2> /dev/null > gpurun_out/r2ai_bench2000_keep$keep.json
done
$T 300 python bench.py --gpus 1 --steps 2000 --warmup 10 --no-config5 --no-cpu-baseline 2> /dev/null > gpurun_out/r2ai_bench2000_default.json
python - <<'PY'
import json
for k in ('keep80', 'keep96', 'default'):
    for n in ('20', '2000'):
        try:
            d = json.load(open(f'gpurun_out/r2ai_bench{n}_{k}.json'))
            print(k, 'steps', n, 'ms', round(d['ms_per_step'], 5), 'e2e_ms', round(d['e2e']['ms_per_step'], 5), 'kernel_ms', round(d['roofline']['kernel_ms'], 5), d['clocks']['sm_mhz'], 'hier', round(d['fg_hier']['ms_per_eval'], 5))
        except Exception as e:
            print(k, n, 'FAILED', type(e).__name__)
PY
