#!/bin/bash
# Round 2, call ag (2 GPUs): the closing state on a 2-GPU box: multi-GPU tests, parity script, the driver's invocation at N = 2 and N = 1.
mkdir -p gpurun_out
T="timeout -k 5"
$T 600 python -m pytest tests/test_gpu_multi.py tests/test_gpu_group.py -m gpu -x -q 2>&1 | tail -3 | tee gpurun_out/r2ag_multi_tests.txt
$T 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29771 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2ag_bench_n2_driver.json 2> gpurun_out/r2ag_bench_n2_driver.err
$T 500 python bench.py --gpus 1 --steps 20 --warmup 5 --no-config5 --no-cpu-baseline > gpurun_out/r2ag_bench_n1_driver.json 2> gpurun_out/r2ag_bench_n1_driver.err
python - <<'PY'
import json
for n in ('n1', 'n2'):
    try:
        d = json.load(open(f'gpurun_out/r2ag_bench_{n}_driver.json'))
        print(n, 'value', round(d['value'], 1), 'ms', round(d['ms_per_step'], 5), 'e2e_ms', round(d['e2e']['ms_per_step'], 5), 'kernel_ms', [round(v, 4) for v in d['roofline']['kernel_ms_per_rank']], d['clocks'], 'hier', round(d['fg_hier']['ms_per_eval'], 5), (d.get('parity') or {}).get('ok'), (d.get('config5') or {}).get('ms_per_eval'))
    except Exception as e:
        print(n, 'FAILED', e)
PY
tail -2 gpurun_out/r2ag_bench_n2_driver.err
