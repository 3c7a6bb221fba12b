#!/bin/bash
# Round 2, call t (1 GPU): host packets gathered per block (8 packets = 128 aligned bytes per store) vs one 16-byte store per warp
# (library of commit 3f2065b), same box, alternating processes; packet stress test first.
mkdir -p gpurun_out
T="timeout -k 5"
$T 600 python -m pytest tests/test_gpu_core.py -m gpu -x -q -k "packet or determinism or kats or full_size" 2>&1 | tail -3 | tee gpurun_out/r2t_tests.log
for i in 1 2; do
  SFH_LIB=$PWD/profiles/ab/libsfhcuda_3f2065b.so $T 200 python profiles/bench_e2e_quick.py 2>&1 | tee -a gpurun_out/r2t_ab.txt
  $T 200 python profiles/bench_e2e_quick.py 2>&1 | tee -a gpurun_out/r2t_ab.txt
done
lscpu | grep -E "Model name|Socket|NUMA node\(s\)" | tee -a gpurun_out/r2t_ab.txt
