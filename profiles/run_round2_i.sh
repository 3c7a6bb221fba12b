#!/bin/bash
# Round 2, call i (2 GPUs): per-warp packet exchange, product-based hier tail, lean prologue; bench at N = 1 (2-bin vs 4-bin tiles, sustained) and N = 2.
mkdir -p gpurun_out
T="timeout -k 5"
$T 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -12 | tee gpurun_out/r2i_gpu_tests.log
$T 400 python bench.py --steps 2000 --warmup 10 --no-config5 --no-cpu-baseline 2> gpurun_out/r2i_bench_n1.err > gpurun_out/r2i_bench_n1.json
$T 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29721 bench.py --gpus 2 --steps 2000 --warmup 10 --no-config5 2> gpurun_out/r2i_bench_n2.err > gpurun_out/r2i_bench_n2.json
SFH_NO_P2P=1 $T 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29723 bench.py --gpus 2 --steps 2000 --warmup 10 --no-config5 2> gpurun_out/r2i_bench_n2_nccl.err > gpurun_out/r2i_bench_n2_nccl.json
python - <<'PY'
import json
for n in ('n1', 'n2', 'n2_nccl'):
    try:
        d = json.load(open(f'gpurun_out/r2i_bench_{n}.json'))
        print(n, 'value', round(d['value'], 1), 'ms', round(d['ms_per_step'], 5), 'e2e_ms', round(d['e2e']['ms_per_step'], 5), 'frac', round(d['roofline']['frac'], 4),
              'kernel_ms', round(d['roofline']['kernel_ms'], 5), d['clocks'], 'launches', d['gpu_launches'], 'hier', round(d['fg_hier']['ms_per_eval'], 5), 'tile', d['config']['tile_bins'])
    except Exception as e:
        print(n, 'FAILED', e)
PY
$T 200 python profiles/bench_latency.py 2>&1 | tee gpurun_out/r2i_latency.txt
$T 300 python profiles/bench_group.py 2 config5 2>&1 | tail -1 | tee gpurun_out/r2i_group.txt
$T 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r2i_hier_launches.csv python profiles/hier_once.py > gpurun_out/r2i_hier_once.log 2>&1
tail -9 gpurun_out/r2i_hier_launches.csv | cut -d, -f5,15 
