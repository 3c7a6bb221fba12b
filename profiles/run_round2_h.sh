#!/bin/bash
# Round 2, call h (2 GPUs): where do the N = 2 step's extra ~16 us come from (per-rank kernel times; N independent shards with
# no exchange), per-kernel durations of the hierarchical call, hierarchical parity errors in units of the backward-error scale.
mkdir -p gpurun_out
T="timeout -k 5"
$T 600 python -m pytest tests/test_gpu_hier.py tests/test_gpu_core.py -m gpu -x -q 2>&1 | tail -6 | tee gpurun_out/r2h_gpu_tests.log
$T 300 python profiles/hier_parity_errors.py 2>&1 | tee gpurun_out/r2h_hier_parity_errors.txt | tail -3
$T 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r2h_hier_launches.csv python profiles/hier_once.py > gpurun_out/r2h_hier_once.log 2>&1
python - <<'PY'
import csv
rows = [r for r in csv.reader(open('gpurun_out/r2h_hier_launches.csv')) if len(r) > 5]
h = rows[0]; ik, iv = h.index("Kernel Name"), h.index("Metric Value")
for r in rows[-12:]:
    print(r[ik][:60], r[iv])
PY
$T 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29721 bench.py --gpus 2 --steps 2000 --warmup 10 --no-config5 2> gpurun_out/r2h_bench_n2.err > gpurun_out/r2h_bench_n2.json
SFH_BENCH_NO_EXCHANGE=1 $T 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29722 bench.py --gpus 2 --steps 2000 --warmup 10 --no-config5 2> gpurun_out/r2h_bench_n2_noex.err > gpurun_out/r2h_bench_n2_noex.json
CUDA_VISIBLE_DEVICES=1 $T 400 python bench.py --steps 2000 --warmup 10 --no-config5 --no-cpu-baseline 2> gpurun_out/r2h_bench_gpu1.err > gpurun_out/r2h_bench_gpu1.json
python - <<'PY'
import json
for n in ('n2', 'n2_noex', 'gpu1'):
    try:
        d = json.load(open(f'gpurun_out/r2h_bench_{n}.json'))
        print(n, 'value', round(d['value'], 1), 'ms', round(d['ms_per_step'], 5), 'e2e_ms', round(d['e2e']['ms_per_step'], 5), 'kernel_ms_per_rank', d['roofline'].get('kernel_ms_per_rank'),
              d['clocks'], 'hier', round(d['fg_hier']['ms_per_eval'], 5))
    except Exception as e:
        print(n, 'FAILED', e)
PY
