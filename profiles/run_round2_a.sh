#!/bin/bash
# First GPU call of round 2: everything built after round 1's GPU budget ran out.
#   gpurun --timeout 2700 -- 'bash profiles/run_round2_a.sh'
# (typical run: 6-8 minutes; 2700 s covers the sum of the per-step limits below, so that gpurun's own limit never fires first --
#  box time is charged for what is used, not for the limit)
# Every step carries its own `timeout -k` (a step that hangs must not take the box with it), and the one kernel that has never
# run on a device -- the two-tiles-in-flight fused kernel, a cluster exchange that could deadlock -- comes LAST, its sweeps
# only if its parity tests finished.
mkdir -p gpurun_out
T="timeout -k 5"
$T 240 python -m pytest tests/test_zz_gpu_nuts.py tests/test_zz_gpu_native.py tests/test_zz_gpu_file.py -q -m gpu 2>&1 | tail -15 | tee gpurun_out/r2a_zz_tests.log
$T 200 python profiles/bench_nuts.py 2>&1 | tee gpurun_out/r2a_nuts.txt
$T 100 python profiles/bench_native.py 2>&1 | tee gpurun_out/r2a_native.txt
$T 240 python bench.py --steps 2000 --warmup 10 2> gpurun_out/r2a_bench.err | tee gpurun_out/r2a_bench.json
# the CPU arm now times both restatements of the reference's fg! (OpenBLAS gemv route / OpenMP nest) and the one-thread BLAS setting:
# the box's 16 host cores have only ever been measured with the (then slower) OpenMP nest
$T 120 python bench.py --impl reference --steps 100 --warmup 3 2> gpurun_out/r2a_bench_reference.err | tee gpurun_out/r2a_bench_reference.json
# fit_templates at 2400 templates: host-resident vs device-resident inverse Hessian (the host BFGS update is ~17 ms per iteration
# there).  Its three kernels are plain grid-stride loops (no barriers): parity first (tests/experimental_gpu_pipe.py -k hessian).
$T 120 python -m pytest tests/experimental_gpu_pipe.py -m gpu -q -k "hessian" 2>&1 | tail -15 | tee gpurun_out/r2a_device_hessian_tests.log
SFH_BENCH_DEVICE_HESSIAN=1 $T 300 python profiles/bench_native.py 2>&1 | tee gpurun_out/r2a_native_device_hessian.txt
# the two-tiles-in-flight fused kernel (variant 3): parity first, then timing against the default on config 3 and on a config-5 shard
$T 180 python -m pytest tests/experimental_gpu_pipe.py -m gpu -x -q -k "not hessian" 2>&1 | tail -15 | tee gpurun_out/r2a_pipe_tests.log
rc=${PIPESTATUS[0]}
if [ "$rc" -ge 124 ]; then
  echo "pipelined parity tests did not finish (rc=$rc): sweeps skipped" | tee gpurun_out/r2a_pipe_config3.txt
  exit 0
fi
for cfg in "0 0 0 0" "0 0 0 3" "8 8 4 3" "8 16 4 3" "8 16 8 3" "16 8 4 3" "16 16 8 3"; do
  $T 60 python profiles/one_config.py $cfg 20 2>&1 | tail -1
done | tee gpurun_out/r2a_pipe_config3.txt
for cfg in "0 0 0 0" "0 0 0 3" "8 8 8 3" "8 16 8 3" "16 8 8 3" "16 16 8 3" "16 16 16 3"; do
  $T 90 python profiles/one_config.py $cfg 10 125000 10000 float32 2>&1 | tail -1
done | tee gpurun_out/r2a_pipe_config5_shard.txt
