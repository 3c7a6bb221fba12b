#!/bin/bash
# Round 2, call v (8 GPUs): closing multi-GPU run -- parity script on 8 ranks, the driver's own N = 8 invocation (weak config 3 + strong
# config 5 + sharded == whole), a 1000-step run of the same, config 5 from ONE process over 8 GPUs.
mkdir -p gpurun_out
T="timeout -k 5"
$T 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29761 tests/mgpu_check.py 2>&1 | tail -2 | tee gpurun_out/r2v_mgpu_check.txt
$T 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29762 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r2v_bench_n8_driver.json 2> gpurun_out/r2v_bench_n8_driver.err
$T 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29763 bench.py --gpus 8 --steps 1000 --warmup 10 --no-config5 > gpurun_out/r2v_bench_n8_1000.json 2> gpurun_out/r2v_bench_n8_1000.err
python - <<'PY'
import json
for n in ('driver', '1000'):
    try:
        d = json.load(open(f'gpurun_out/r2v_bench_n8_{n}.json'))
        print(n, 'value', round(d['value'], 1), 'ms', round(d['ms_per_step'], 5), 'e2e_ms', round(d['e2e']['ms_per_step'], 5), 'kernel_ms_per_rank', [round(v, 4) for v in d['roofline']['kernel_ms_per_rank']], d['clocks'], 'hier', round(d['fg_hier']['ms_per_eval'], 5))
        print('  parity', d['parity'])
        if 'config5' in d: print('  config5', {k: d['config5'][k] for k in ('ms_per_eval', 'aggregate_GBps', 'per_gpu', 'exchange', 'parity')})
    except Exception as e:
        print(n, 'FAILED', e)
PY
tail -3 gpurun_out/r2v_bench_n8_driver.err
$T 300 python profiles/bench_group.py 8 config5 2>&1 | tail -1 | tee gpurun_out/r2v_group.txt
