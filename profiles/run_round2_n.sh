#!/bin/bash
# Round 2, call n (1 GPU): early programmatic launch (dependents released at kernel start, producer streams ahead of griddepcontrol.wait),
# upload kernel instead of the copy node, completion by self-validating packets in pinned memory instead of cudaStreamSynchronize.
# A/B by environment switch in the same call; suite first.
mkdir -p gpurun_out
T="timeout -k 5"
$T 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 | tee gpurun_out/r2n_gpu_tests.log
for v in "0 0" "1 0" "0 1" "1 1"; do
  set -- $v
  echo "== SFH_PDL_EARLY=$1 SFH_HOST_PACKETS=$2" | tee -a gpurun_out/r2n_latency.txt
  SFH_PDL_EARLY=$1 SFH_HOST_PACKETS=$2 $T 200 python profiles/bench_latency.py 2>&1 | tee -a gpurun_out/r2n_latency.txt
done
for v in 0 1; do
  SFH_PDL_EARLY=$v $T 400 python bench.py --steps 2000 --warmup 10 --no-config5 --no-cpu-baseline 2> gpurun_out/r2n_bench_pdl$v.err > gpurun_out/r2n_bench_pdl$v.json
done
python - <<'PY'
import json
for n in ('pdl0', 'pdl1'):
    try:
        d = json.load(open(f'gpurun_out/r2n_bench_{n}.json'))
        print(n, 'value', round(d['value'], 1), 'ms', round(d['ms_per_step'], 5), 'e2e_ms', round(d['e2e']['ms_per_step'], 5), 'frac', round(d['roofline']['frac'], 4),
              'kernel_ms', round(d['roofline']['kernel_ms'], 5), d['clocks'], 'launches', d['gpu_launches'], 'hier', round(d['fg_hier']['ms_per_eval'], 5), 'tile', d['config']['tile_bins'])
    except Exception as e:
        print(n, 'FAILED', e)
PY
tail -5 gpurun_out/r2n_bench_pdl1.err
