#!/bin/bash
# First GPU call of round 2: everything built after round 1's GPU budget ran out.
#   gpurun --timeout 600 -- 'bash profiles/run_round2_a.sh'
mkdir -p gpurun_out
python -m pytest tests/test_zz_gpu_nuts.py tests/test_zz_gpu_native.py tests/test_zz_gpu_file.py -q -m gpu 2>&1 | tail -15 | tee gpurun_out/r2a_zz_tests.log
timeout 200 python profiles/bench_nuts.py 2>&1 | tee gpurun_out/r2a_nuts.txt
timeout 100 python profiles/bench_native.py 2>&1 | tee gpurun_out/r2a_native.txt
python bench.py --steps 2000 --warmup 10 2> gpurun_out/r2a_bench.err | tee gpurun_out/r2a_bench.json
