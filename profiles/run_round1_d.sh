#!/bin/bash
# round-1 final measurement set (1 GPU): bench line, ncu launch list, ncu full capture of fused + batched kernels, latency, batched
mkdir -p gpurun_out
python bench.py > gpurun_out/bench_r1_final.json 2> gpurun_out/bench_r1_final.err
python -c "
import json; d=json.load(open('gpurun_out/bench_r1_final.json')); print('value',d['value'],'e2e',d['e2e']['value'],'frac',d['roofline']['frac'],'kernel_ms',d['roofline']['kernel_ms'],'cfg',d['config']['tile_bins'],d['config']['cluster'],d['config']['consumer_warps'],'cpu',d['cpu_baseline']['value'],d['cpu_baseline']['cores'], d['clocks'], 'launches', d['gpu_launches'])"
python bench.py --impl reference --steps 100 --warmup 3 > gpurun_out/bench_r1_reference.json 2>/dev/null; cut -c1-300 gpurun_out/bench_r1_reference.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/launches_r1_final.csv python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launch_final.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:sfh_fg_fused -s 3 -c 1 -o gpurun_out/prof_fused_r1_final python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full_final.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:sfh_batched_logl_mma -s 1 -c 1 -o gpurun_out/prof_batched_r1_final python profiles/bench_batched.py > gpurun_out/ncu_full_final2.log 2>&1
python profiles/bench_batched.py 2>&1 | tee gpurun_out/batched_r1_final.txt
python profiles/bench_latency.py 2>&1 | tee gpurun_out/latency_r1_final.txt
python profiles/bench_config5.py --steps 20 2>&1 | tail -1 | tee gpurun_out/config5_n1.txt
ls -la gpurun_out | tail -12
