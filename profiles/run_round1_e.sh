#!/bin/bash
# round-1 closing measurement set (1 GPU): full GPU test suite, smoke, bench lines, ncu launch list + full captures, side benchmarks
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -4 | tee gpurun_out/pytest_gpu_r1e.txt
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
python bench.py > gpurun_out/bench_r1e.json 2> gpurun_out/bench_r1e.err
python -c "
import json; d=json.load(open('gpurun_out/bench_r1e.json')); print('value',d['value'],'e2e',d['e2e']['value'],'frac',d['roofline']['frac'],'kernel_ms',d['roofline']['kernel_ms'],'cpu',d['cpu_baseline']['value'],d['cpu_baseline']['cores'], d['clocks'], 'launches', d['gpu_launches'])"
python bench.py --impl reference --steps 100 --warmup 3 > gpurun_out/bench_r1e_reference.json 2>/dev/null; cut -c1-300 gpurun_out/bench_r1e_reference.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/launches_r1e.csv python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launch_r1e.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:sfh_fg_fused -s 3 -c 1 -o gpurun_out/prof_fused_r1e python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full_r1e.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:sfh_batched_logl_mma -s 1 -c 1 -o gpurun_out/prof_batched_r1e python profiles/bench_batched.py k6 > gpurun_out/ncu_full_r1e2.log 2>&1
python profiles/bench_batched.py k6 2>&1 | tee gpurun_out/batched_r1e.txt
python profiles/bench_latency.py 2>&1 | tee gpurun_out/latency_r1e.txt
ls -la gpurun_out | tail -14
