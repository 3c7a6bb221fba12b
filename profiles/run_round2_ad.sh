#!/bin/bash
# Round 2, call ad (1 GPU): does the tiling choice for config 3 still hold with the L2-resident head and the reworked launch path?
# tile width x cluster size, isolated back-to-back kernel, then the two best in bench.py (20 steps and 2000 steps).
mkdir -p gpurun_out
T="timeout -k 5"
for cfg in "0 2 1 4" "0 4 1 4" "0 8 1 4" "0 16 1 4" "0 4 2 4" "0 8 2 4" "0 16 2 4"; do
  $T 60 python profiles/one_config.py $cfg 60 2>&1 | tail -1
done | tee gpurun_out/r2ad_config3_tiles.txt
for tile in 2 4 8; do
  $T 300 python bench.py --gpus 1 --steps 20 --warmup 5 --tile $tile --cluster 1 --no-config5 --no-cpu-baseline 2> /dev/null > gpurun_out/r2ad_bench20_tile$tile.json
  $T 300 python bench.py --gpus 1 --steps 2000 --warmup 10 --tile $tile --cluster 1 --no-config5 --no-cpu-baseline 2> /dev/null > gpurun_out/r2ad_bench2000_tile$tile.json
done
python - <<'PY'
import json
for t in (2, 4, 8):
    for n in ('20', '2000'):
        try:
            d = json.load(open(f'gpurun_out/r2ad_bench{n}_tile{t}.json'))
            print('tile', t, 'steps', n, 'ms', round(d['ms_per_step'], 5), 'e2e_ms', round(d['e2e']['ms_per_step'], 5), 'kernel_ms', round(d['roofline']['kernel_ms'], 5), d['clocks']['sm_mhz'], 'hier', round(d['fg_hier']['ms_per_eval'], 5), 'ring', d['config']['ring_slots'], 'kt', d['config']['chunks_per_tile'])
        except Exception as e:
            print(t, n, 'FAILED', e)
PY
