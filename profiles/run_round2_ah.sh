#!/bin/bash
# Round 2, call ah (8 GPUs): the closing state at N = 8: parity script, the driver's invocation.
mkdir -p gpurun_out
T="timeout -k 5"
$T 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29781 tests/mgpu_check.py 2>&1 | tail -2 | tee gpurun_out/r2ah_mgpu_check.txt
$T 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29782 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r2ah_bench_n8_driver.json 2> gpurun_out/r2ah_bench_n8_driver.err
python - <<'PY'
import json
d = json.load(open('gpurun_out/r2ah_bench_n8_driver.json'))
print('n8 value', round(d['value'], 1), 'ms', round(d['ms_per_step'], 5), 'e2e_ms', round(d['e2e']['ms_per_step'], 5), 'kernel_ms_per_rank', [round(v, 4) for v in d['roofline']['kernel_ms_per_rank']], d['clocks'], 'hier', round(d['fg_hier']['ms_per_eval'], 5))
print('  parity', d['parity']['ok'], d['parity']['grad_max_err_over_tol'], 'config5', d['config5']['ms_per_eval'], d['config5']['per_gpu']['roofline_frac_kernel'], d['config5']['parity']['ok'])
PY
tail -2 gpurun_out/r2ah_bench_n8_driver.err
