#!/bin/bash
# Round 2, call q (2 GPUs): sharded gradient warps push all their entries before polling; hierarchical exchange polled by all 8 warps.
# Full suite on the 2-GPU box (multi-GPU tests included), bench at N = 2.
mkdir -p gpurun_out
T="timeout -k 5"
$T 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 | tee gpurun_out/r2q_gpu_tests.log
$T 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29751 bench.py --gpus 2 --steps 2000 --warmup 10 --no-config5 2> gpurun_out/r2q_bench_n2.err > gpurun_out/r2q_bench_n2.json
python - <<'PY'
import json
for n in ('n2',):
    try:
        d = json.load(open(f'gpurun_out/r2q_bench_{n}.json'))
        print(n, 'value', round(d['value'], 1), 'ms', round(d['ms_per_step'], 5), 'e2e_ms', round(d['e2e']['ms_per_step'], 5), 'frac', round(d['roofline']['frac'], 4),
              'kernel_ms', [round(v, 4) for v in d['roofline']['kernel_ms_per_rank']], d['clocks'], 'launches', d['gpu_launches'], 'hier', round(d['fg_hier']['ms_per_eval'], 5), d.get('parity', {}).get('ok'))
    except Exception as e:
        print(n, 'FAILED', e)
PY
tail -3 gpurun_out/r2q_bench_n2.err
$T 300 python profiles/bench_group.py 2 config5half 2>&1 | tail -1 | tee gpurun_out/r2q_group.txt
