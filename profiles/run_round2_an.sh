#!/bin/bash
# Round 2, call an (1 GPU): CTA pairs on the narrower stacks (where the chooser keeps single CTAs because a pair's tile is < 72 KB per CTA)
mkdir -p gpurun_out
T="timeout -k 5"
run() { $T 90 python profiles/one_config.py "$@" 2>&1 | tail -1; }
{
run 0 16 1 4 60 40000 500 float64;   run 0 32 2 4 60 40000 500 float64
run 0 4 1 4 60 11250 2000 float64;   run 0 8 2 4 60 11250 2000 float64
run 0 16 1 4 60 200000 1000 float32; run 0 32 2 4 60 200000 1000 float32
run 0 16 1 4 60 100000 500 float64;  run 0 32 2 4 60 100000 500 float64
run 0 64 1 4 60 10000 100 float64;   run 0 32 1 4 60 10000 100 float64; run 0 16 1 4 60 10000 100 float64
} | tee gpurun_out/r2an_pairs_narrow.txt
