#!/bin/bash
# Round 2, call j (2 GPUs): source-level profiles of the hierarchical finalize / prologue kernels; where the group's host-API overhead comes from
mkdir -p gpurun_out
T="timeout -k 5"
$T 300 ncu --set full --clock-control none --import-source on -k regex:sfh_finalize -s 12 -c 1 -o gpurun_out/prof_finalize_hier python profiles/hier_once.py > gpurun_out/r2j_ncu1.log 2>&1
$T 300 ncu --set full --clock-control none --import-source on -k regex:prologue2 -s 8 -c 1 -o gpurun_out/prof_prologue2 python profiles/hier_once.py > gpurun_out/r2j_ncu2.log 2>&1
for a in "1 config5half" "2 config5half" "2 config5" "2 config3" "1 config3"; do $T 300 python profiles/bench_group.py $a 2>&1 | tail -1; done | tee gpurun_out/r2j_group.txt
