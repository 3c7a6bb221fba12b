#!/bin/bash
# Round 2, call l (8 GPUs): the multi-GPU parity script on 8 ranks, the bench line at N = 8 (weak config 3 + strong config 5 + sharded == whole),
# and config 5 / config 3 from ONE process over 8 GPUs.
mkdir -p gpurun_out
T="timeout -k 5"
$T 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29711 tests/mgpu_check.py 2>&1 | tail -3 | tee gpurun_out/r2l_mgpu_check.txt
$T 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29712 bench.py --gpus 8 --steps 1000 --warmup 10 > gpurun_out/r2l_bench_n8.json 2> gpurun_out/r2l_bench_n8.err
python - <<'PY'
import json
try:
    d = json.load(open('gpurun_out/r2l_bench_n8.json'))
    print('n8 value', round(d['value'], 1), 'ms', round(d['ms_per_step'], 5), 'e2e_ms', round(d['e2e']['ms_per_step'], 5), 'kernel_ms_per_rank', [round(v, 4) for v in d['roofline']['kernel_ms_per_rank']], d['clocks'], 'hier', round(d['fg_hier']['ms_per_eval'], 5))
    print('  parity', d['parity']); print('  config5', {k: d['config5'][k] for k in ('ms_per_eval', 'aggregate_GBps', 'per_gpu', 'exchange', 'parity')})
except Exception as e:
    print('n8 FAILED', e)
PY
tail -3 gpurun_out/r2l_bench_n8.err
$T 300 python profiles/bench_group.py 8 config5 2>&1 | tail -1 | tee gpurun_out/r2l_group.txt
$T 300 python profiles/bench_group.py 8 config3 2>&1 | tail -1 | tee -a gpurun_out/r2l_group.txt
$T 300 python -m pytest tests/test_gpu_group.py -m gpu -x -q 2>&1 | tail -3 | tee gpurun_out/r2l_group_tests.txt
