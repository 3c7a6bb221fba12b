#!/bin/bash
# Round 2, call m (2 GPUs): logL pushed by block 0 at kernel start, no ticket on one GPU; suite + bench + warm-cache per-kernel durations
mkdir -p gpurun_out
T="timeout -k 5"
$T 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 | tee gpurun_out/r2m_gpu_tests.log
$T 400 python bench.py --steps 2000 --warmup 10 --no-config5 --no-cpu-baseline 2> gpurun_out/r2m_bench_n1.err > gpurun_out/r2m_bench_n1.json
$T 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29721 bench.py --gpus 2 --steps 2000 --warmup 10 --no-config5 2> gpurun_out/r2m_bench_n2.err > gpurun_out/r2m_bench_n2.json
python - <<'PY'
import json
for n in ('n1', 'n2'):
    try:
        d = json.load(open(f'gpurun_out/r2m_bench_{n}.json'))
        print(n, 'value', round(d['value'], 1), 'ms', round(d['ms_per_step'], 5), 'e2e_ms', round(d['e2e']['ms_per_step'], 5), 'frac', round(d['roofline']['frac'], 4),
              'kernel_ms', round(d['roofline']['kernel_ms'], 5), d['clocks'], 'launches', d['gpu_launches'], 'hier', round(d['fg_hier']['ms_per_eval'], 5), 'tile', d['config']['tile_bins'])
    except Exception as e:
        print(n, 'FAILED', e)
PY
$T 200 python profiles/bench_latency.py 2>&1 | tee gpurun_out/r2m_latency.txt
$T 300 ncu --cache-control none --clock-control none --metrics gpu__time_duration.sum,lts__t_sector_hit_rate.pct,dram__bytes_read.sum -k regex:"prologue2|finalize" -s 30 -c 8 --csv --log-file gpurun_out/r2m_hier_warm.csv python profiles/hier_once.py > gpurun_out/r2m_hier_warm.log 2>&1
python - <<'PY'
import csv
rows = [r for r in csv.reader(open('gpurun_out/r2m_hier_warm.csv')) if len(r) > 5]
h = rows[0]; ik, im, iv = h.index("Kernel Name"), h.index("Metric Name"), h.index("Metric Value")
for r in rows[1:]:
    print(r[ik][:40], r[im], r[iv])
PY
