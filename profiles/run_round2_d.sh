#!/bin/bash
# Round 2, call d (1 GPU): folded hierarchical path, conversion-free F32 unpack, back-to-back kernel timing; suite, sweeps, bench.
mkdir -p gpurun_out
T="timeout -k 5"
$T 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 | tee gpurun_out/r2d_gpu_tests.log
for shape in "60000 2400 float64" "40000 500 float64" "10000 100 float64" "11250 2000 float64" "11250 2000 float32" "200000 1000 float32" "125000 10000 float32"; do
  for cfg in "0 0 0 1" "0 0 0 4"; do
    $T 90 python profiles/one_config.py $cfg 50 $shape 2>&1 | tail -1
  done
done | tee gpurun_out/r2d_sweep_shapes.txt
for cfg in "0 2 1 4" "0 4 1 4" "0 8 1 4" "8 16 4 1"; do $T 60 python profiles/one_config.py $cfg 50 2>&1 | tail -1; done | tee gpurun_out/r2d_config3.txt
SFH_F32_FAST=0 $T 90 python profiles/one_config.py 0 0 0 4 50 125000 10000 float32 2>&1 | tail -1 | tee gpurun_out/r2d_f32_slow_unpack.txt
$T 200 python profiles/bench_latency.py 2>&1 | tee gpurun_out/r2d_latency.txt
$T 600 python bench.py --steps 2000 --warmup 10 2> gpurun_out/r2d_bench.err | tee gpurun_out/r2d_bench.json | cut -c1-1200
python -c "
import json; d=json.load(open('gpurun_out/r2d_bench.json'))
print('value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e'],'frac',d['roofline']['frac'],'kernel_ms',d['roofline']['kernel_ms'],d['clocks'])
print('hier',d['fg_hier']); print('config5',d['config5']['ms_per_eval'],d['config5']['per_gpu'])"
