#!/bin/bash
# Round 2, call ak (2 GPUs): sharded flat finalize without ticket / last block (the stream kernel bumps the exchange epoch): multi-GPU
# tests, parity script, the driver's invocation at N = 2 (twice) and N = 1.
mkdir -p gpurun_out
T="timeout -k 5"
$T 600 python -m pytest tests/test_gpu_multi.py tests/test_gpu_group.py -m gpu -x -q 2>&1 | tail -3 | tee gpurun_out/r2ak_multi_tests.txt
for i in 1 2; do
$T 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2979$i bench.py --gpus 2 --steps 20 --warmup 5 --no-config5 > gpurun_out/r2ak_bench_n2_driver$i.json 2> gpurun_out/r2ak_bench_n2_driver$i.err
done
$T 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29793 bench.py --gpus 2 --steps 2000 --warmup 10 --no-config5 > gpurun_out/r2ak_bench_n2_2000.json 2> gpurun_out/r2ak_bench_n2_2000.err
$T 500 python bench.py --gpus 1 --steps 20 --warmup 5 --no-config5 --no-cpu-baseline > gpurun_out/r2ak_bench_n1_driver.json 2> gpurun_out/r2ak_bench_n1_driver.err
$T 500 python bench.py --gpus 1 --steps 2000 --warmup 10 --no-config5 --no-cpu-baseline > gpurun_out/r2ak_bench_n1_2000.json 2> gpurun_out/r2ak_bench_n1_2000.err
python - <<'PY'
import json
for n in ('n1_driver', 'n2_driver1', 'n2_driver2', 'n1_2000', 'n2_2000'):
    try:
        d = json.load(open(f'gpurun_out/r2ak_bench_{n}.json'))
        print(n, 'value', round(d['value'], 1), 'ms', round(d['ms_per_step'], 5), 'e2e_ms', round(d['e2e']['ms_per_step'], 5), 'kernel_ms', [round(v, 4) for v in d['roofline']['kernel_ms_per_rank']], d['clocks']['sm_mhz'], 'hier', round(d['fg_hier']['ms_per_eval'], 5), (d.get('parity') or {}).get('ok'))
    except Exception as e:
        print(n, 'FAILED', e)
PY
