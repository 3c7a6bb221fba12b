#!/bin/bash
# Round 2, call ae (1 GPU): config 3 with 8-bin tiles on CTA pairs (74 clusters) vs the default 4-bin tiles on single CTAs, in bench.py
mkdir -p gpurun_out
T="timeout -k 5"
for rep in 1 2; do
for cfg in "4 1" "8 2"; do
  set -- $cfg
  $T 300 python bench.py --gpus 1 --steps 20 --warmup 5 --tile $1 --cluster $2 --no-config5 --no-cpu-baseline 2> /dev/null > gpurun_out/r2ae_bench20_t$1c$2_$rep.json
  $T 300 python bench.py --gpus 1 --steps 2000 --warmup 10 --tile $1 --cluster $2 --no-config5 --no-cpu-baseline 2> /dev/null > gpurun_out/r2ae_bench2000_t$1c$2_$rep.json
done
done
python - <<'PY'
import json
for rep in (1, 2):
  for t in ('t4c1', 't8c2'):
    for n in ('20', '2000'):
        try:
            d = json.load(open(f'gpurun_out/r2ae_bench{n}_{t}_{rep}.json'))
            print(rep, t, 'steps', n, 'ms', round(d['ms_per_step'], 5), 'e2e_ms', round(d['e2e']['ms_per_step'], 5), 'kernel_ms', round(d['roofline']['kernel_ms'], 5), d['clocks']['sm_mhz'], 'hier', round(d['fg_hier']['ms_per_eval'], 5), 'ring', d['config']['ring_slots'], 'kt', d['config']['chunks_per_tile'], 'ncl', d['config']['n_clusters'])
        except Exception as e:
            print(t, n, 'FAILED', e)
PY
