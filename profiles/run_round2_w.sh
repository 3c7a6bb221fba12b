#!/bin/bash
# Round 2, call w (1 GPU): L2-resident head of every CTA's tile sequence (evict_last for the first keep_stages stages): sweep of the budget
mkdir -p gpurun_out
T="timeout -k 5"
for keep in 0 24 48 64 80 96 112; do
  for shape in "60000 2400 float64" "40000 500 float64" "125000 10000 float32"; do
    echo -n "keep_mb=$keep " ; SFH_L2_KEEP_MB=$keep $T 90 python profiles/one_config.py 0 0 0 4 60 $shape 2>&1 | tail -1
  done
done | tee gpurun_out/r2w_l2keep_sweep.txt
for keep in 0 64; do
  SFH_L2_KEEP_MB=$keep $T 200 python profiles/bench_e2e_quick.py 2>&1 | sed "s/^/keep_mb=$keep /" | tee -a gpurun_out/r2w_l2keep_e2e.txt
done
