#!/bin/bash
mkdir -p gpurun_out
python profiles/bench_batched.py 2>&1 | tee gpurun_out/batched_r1b.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 100 --csv --log-file gpurun_out/launches_r1b.csv python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launch_b.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:sfh_fg_fused -s 3 -c 1 -o gpurun_out/prof_fused_r1b python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full_b.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:sfh_batched_logl -s 1 -c 1 -o gpurun_out/prof_batched_r1b python profiles/bench_batched.py > gpurun_out/ncu_full_b2.log 2>&1
ls -la gpurun_out | tail -8
