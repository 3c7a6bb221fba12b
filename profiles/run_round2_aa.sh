#!/bin/bash
# Round 2, call aa (1 GPU): evict_last marks on the gradient partials and on the hierarchical tables / factors (what the finalize and
# prologue kernels read) vs the previous build, C caller, alternating processes on one box; hierarchical tests first.
mkdir -p gpurun_out
T="timeout -k 5"
$T 900 python -m pytest tests/test_gpu_hier.py tests/test_gpu_core.py -m gpu -x -q 2>&1 | tail -3 | tee gpurun_out/r2aa_tests.log
gcc -O2 -std=c99 -Iinclude profiles/latency_c.c -Lstarformationhistories.jl_b200 -lsfhcuda -Wl,-rpath,$PWD/starformationhistories.jl_b200 -lm -o /tmp/latency_c || exit 1
mkdir -p /tmp/prev && cp profiles/ab/libsfhcuda_l2head.so /tmp/prev/libsfhcuda.so
gcc -O2 -std=c99 -Iinclude profiles/latency_c.c -L/tmp/prev -lsfhcuda -Wl,-rpath,/tmp/prev -lm -o /tmp/latency_c_prev || exit 1
for i in 1 2; do
  echo "== previous build (L2 head only)" | tee -a gpurun_out/r2aa_latency_c.txt
  timeout 300 /tmp/latency_c_prev 2>&1 | tee -a gpurun_out/r2aa_latency_c.txt
  echo "== tree (partials + hierarchical tables evict_last)" | tee -a gpurun_out/r2aa_latency_c.txt
  timeout 300 /tmp/latency_c 2>&1 | tee -a gpurun_out/r2aa_latency_c.txt
done
