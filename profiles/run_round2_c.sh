#!/bin/bash
# Round 2, call c (1 GPU): v2 as the default (8 exchange slots), new finalize tail; full GPU suite, sweeps, ncu captures, new bench line.
mkdir -p gpurun_out
T="timeout -k 5"
$T 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 | tee gpurun_out/r2c_gpu_tests.log
for shape in "60000 2400 float64" "40000 500 float64" "10000 100 float64" "11250 2000 float64" "11250 2000 float32" "200000 1000 float32" "125000 10000 float32"; do
  for cfg in "0 0 0 1" "0 0 0 4"; do
    $T 90 python profiles/one_config.py $cfg 20 $shape 2>&1 | tail -1
  done
done | tee gpurun_out/r2c_sweep_shapes.txt
# config 3: where do the missing ~15 us go?  logL-only instantiation (A warps release the ring; no B pass), other tilings
for cfg in "0 2 1 4" "0 4 1 4" "0 8 1 4"; do
  SFH_WANT_G=0 $T 60 python profiles/one_config.py $cfg 20 2>&1 | tail -1
  $T 60 python profiles/one_config.py $cfg 20 2>&1 | tail -1
done | tee gpurun_out/r2c_config3_experiments.txt
# ncu: the stream kernel on config 3 and on the config-5 shard (full set + source), launch list of the bench command
$T 600 ncu --set full --clock-control none --import-source on -k regex:sfh_fg_fused2 -s 3 -c 1 -o gpurun_out/prof_v2_config3 python profiles/one_config.py 0 0 0 4 5 > gpurun_out/r2c_ncu1.log 2>&1
$T 600 ncu --set full --clock-control none --import-source on -k regex:sfh_fg_fused2 -s 3 -c 1 -o gpurun_out/prof_v2_config5shard python profiles/one_config.py 0 0 0 4 5 125000 10000 float32 > gpurun_out/r2c_ncu2.log 2>&1
$T 600 python bench.py --steps 2000 --warmup 10 2> gpurun_out/r2c_bench.err | tee gpurun_out/r2c_bench.json | cut -c1-1500
$T 300 python bench.py --impl reference --steps 20 --warmup 3 2> gpurun_out/r2c_bench_reference.err | tee gpurun_out/r2c_bench_reference.json | cut -c1-600
$T 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_r2c.csv python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-config5 > gpurun_out/r2c_ncu_launch.log 2>&1
ls -la gpurun_out | tail -12
