#!/bin/bash
# Round 2, call u (1 GPU): latency of the host-synchronous calls from a C caller (no ctypes in the loop), graphs on / off, packets on / off
mkdir -p gpurun_out
gcc -O2 -std=c99 -Iinclude profiles/latency_c.c -Lstarformationhistories.jl_b200 -lsfhcuda -Wl,-rpath,$PWD/starformationhistories.jl_b200 -lm -o /tmp/latency_c || exit 1
for v in "1 0" "0 0" "1 1"; do
  set -- $v
  echo "== SFH_HOST_PACKETS=$1 SFH_NO_GRAPH=$2" | tee -a gpurun_out/r2u_latency_c.txt
  SFH_HOST_PACKETS=$1 SFH_NO_GRAPH=$2 timeout 300 /tmp/latency_c 2>&1 | tee -a gpurun_out/r2u_latency_c.txt
done
