#!/bin/bash
# Round 2, call ac (1 GPU): compute-sanitizer over the paths touched this round (upload kernel, finalize kernels with the gathered host
# packets, hierarchical finalize, stream kernel with the L2 policies): memcheck on the packet / KAT / hierarchical tests, racecheck on the
# finalize and prologue kernels' shared memory.
mkdir -p gpurun_out
T="timeout -k 5"
$T 900 compute-sanitizer --tool memcheck --error-exitcode 77 python -m pytest tests/test_gpu_core.py -m gpu -x -q -k "packet or kats or edge or determinism" 2>&1 | tail -6 | tee gpurun_out/r2ac_memcheck_core.txt
$T 900 compute-sanitizer --tool memcheck --error-exitcode 77 python -m pytest tests/test_gpu_hier.py -m gpu -x -q -k "fg_hier or free_masks" 2>&1 | tail -6 | tee gpurun_out/r2ac_memcheck_hier.txt
$T 900 compute-sanitizer --tool racecheck --error-exitcode 77 --kernel-regex kns=sfh_finalize --kernel-regex kns=sfh_hier_prologue2 --kernel-regex kns=sfh_copy_in python -m pytest tests/test_gpu_hier.py tests/test_gpu_core.py -m gpu -x -q -k "fg_hier_parity or packet or kats" 2>&1 | tail -6 | tee gpurun_out/r2ac_racecheck.txt
