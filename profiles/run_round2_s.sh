#!/bin/bash
# Round 2, call s (1 GPU): A/B of two library builds on ONE box (commit 3f2065b vs the tree: did the finalize restructure cost the flat
# call anything, or was r2r's 201 us the box?), then the closing ncu evidence with the reports summarised on the box (64 MiB limit).
mkdir -p gpurun_out
T="timeout -k 5"
for i in 1 2; do
  SFH_LIB=$PWD/profiles/ab/libsfhcuda_3f2065b.so $T 200 python profiles/bench_e2e_quick.py 2>&1 | tee -a gpurun_out/r2s_ab.txt
  $T 200 python profiles/bench_e2e_quick.py 2>&1 | tee -a gpurun_out/r2s_ab.txt
done
nvidia-smi --query-gpu=name,pcie.link.gen.current,pcie.link.width.current --format=csv | tee -a gpurun_out/r2s_ab.txt
lscpu | grep -E "Model name|Socket|NUMA node\(s\)" | tee -a gpurun_out/r2s_ab.txt
$T 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_final_launches.csv python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-config5 > gpurun_out/r2s_ncu_launch.log 2>&1
$T 600 ncu --set full --clock-control none --import-source on -k regex:sfh_fg_fused2 -s 3 -c 1 -o gpurun_out/r2_final_config3 python profiles/one_config.py 0 0 0 4 5 > gpurun_out/r2s_ncu1.log 2>&1
$T 600 ncu --set full --clock-control none -k regex:sfh_fg_fused2 -s 3 -c 1 -o /tmp/r2_final_config5shard python profiles/one_config.py 0 0 0 4 5 125000 10000 float32 > gpurun_out/r2s_ncu2.log 2>&1
$T 600 ncu --set full --clock-control none -k regex:"sfh_finalize|sfh_copy_in|prologue2" -s 2 -c 8 -o /tmp/r2_final_small python profiles/hier_once.py > gpurun_out/r2s_ncu3.log 2>&1
ncu -i /tmp/r2_final_config5shard.ncu-rep --page raw --csv > gpurun_out/r2_final_config5shard_raw.csv 2>/dev/null
ncu -i /tmp/r2_final_small.ncu-rep --page raw --csv > gpurun_out/r2_final_small_raw.csv 2>/dev/null
ls -la gpurun_out | tail -12; du -sh gpurun_out
