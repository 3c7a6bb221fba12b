#!/bin/bash
# round-1 measurement script (run under gpurun): bench line, tile/cluster sweep, ncu launch list + full capture
mkdir -p gpurun_out
python bench.py --steps 200 --warmup 10 > gpurun_out/bench_r1a.json 2> gpurun_out/bench_r1a.err
tail -c 3000 gpurun_out/bench_r1a.json
for cfg in "64 8" "32 4" "32 8" "16 2" "16 4" "16 8"; do
  set -- $cfg
  echo "== tile=$1 cluster=$2" >> gpurun_out/sweep_r1a.txt
  timeout 300 python bench.py --steps 100 --warmup 5 --no-cpu-baseline --tile $1 --cluster $2 >> gpurun_out/sweep_r1a.txt 2>&1
done
python - <<'PY'
import json
for ln in open('gpurun_out/sweep_r1a.txt'):
    if ln.startswith('=='): print(ln.strip())
    elif ln.startswith('{'):
        d=json.loads(ln); print('   value=%.1f e2e=%.1f kernel_ms=%.4f frac=%.3f' % (d['value'], d['e2e']['value'], d['roofline']['kernel_ms'], d['roofline']['frac']))
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches_r1a.csv python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:sfh_fg_fused -s 3 -c 2 -o gpurun_out/prof_fused_r1a python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out
