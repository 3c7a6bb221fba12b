#!/bin/bash
# Round 2, call b: first run of the warp-specialised stream kernel (sfh_fused2.cuh, variant 4): parity, then sweeps vs v1.
mkdir -p gpurun_out
T="timeout -k 5"
$T 420 python -m pytest tests/test_gpu_core.py -m gpu -x -q -k "stream_kernel or reference_kats or fg_parity" 2>&1 | tail -15 | tee gpurun_out/r2b_v2_tests.log
rc=${PIPESTATUS[0]}
if [ "$rc" -ge 124 ]; then echo "v2 parity tests hung (rc=$rc): sweeps skipped" | tee gpurun_out/r2b_sweep_config3.txt; exit 0; fi
# config 3: v1 default, then v2 across (bins per tile, cluster)
for cfg in "8 8 2 1" "8 16 4 1" "0 2 1 4" "0 2 2 4" "0 4 1 4" "0 4 2 4" "0 8 1 4" "0 16 1 4"; do
  $T 60 python profiles/one_config.py $cfg 20 2>&1 | tail -1
done | tee gpurun_out/r2b_sweep_config3.txt
# config-5 shard (one eighth of 10^6 x 10^4 F32)
for cfg in "8 8 4 1" "0 4 2 4" "0 4 4 4" "0 8 4 4" "0 8 8 4"; do
  $T 90 python profiles/one_config.py $cfg 10 125000 10000 float32 2>&1 | tail -1
done | tee gpurun_out/r2b_sweep_config5_shard.txt
# other shapes: config 2 (40000 x 500 F64), config 1 (10000 x 100 F64), the reference's CI shape (11250 x 2000 F32 / F64)
for shape in "40000 500 float64" "10000 100 float64" "11250 2000 float64" "11250 2000 float32" "200000 1000 float32"; do
  for cfg in "0 0 0 1" "0 0 0 4"; do
    $T 60 python profiles/one_config.py $cfg 20 $shape 2>&1 | tail -1
  done
done | tee gpurun_out/r2b_sweep_shapes.txt
# the whole 40 GB config-5 stack on one GPU
for cfg in "0 0 0 1" "0 0 0 4"; do
  $T 240 python profiles/one_config.py $cfg 5 1000000 10000 float32 2>&1 | tail -1
done | tee gpurun_out/r2b_config5_full.txt
# the full GPU suite with v2 as the default
$T 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/r2b_gpu_tests.log
