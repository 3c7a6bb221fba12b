#!/bin/bash
# 8-GPU closing check: multi-GPU parity script on 8 ranks, weak-scaling bench line and strong-scaling config 5 at N = 8
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29711 tests/mgpu_check.py 2>&1 | tail -3
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29712 bench.py --gpus 8 --steps 1000 --warmup 10 > gpurun_out/scale8_r1f.json 2> gpurun_out/scale8_r1f.err
cut -c1-400 gpurun_out/scale8_r1f.json
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29713 profiles/bench_config5.py --steps 30 2>/dev/null | grep workload | tee gpurun_out/config5_n8_r1f.json
