#!/bin/bash
# 1 -> N GPU scaling on ONE box: the weak-scaling bench line (bench.py) and the strong-scaling config-5 run.
NMAX=${1:-8}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
for n in 1 2 4 8; do
  [ $n -gt $NMAX ] && break
  if [ $n = 1 ]; then
    python bench.py --steps 1000 --warmup 10 --no-cpu-baseline > gpurun_out/scale_n$n.json 2> gpurun_out/scale_n$n.err
    python profiles/bench_config5.py --steps 20 2>/dev/null | tail -1 > gpurun_out/config5_n$n.json
  else
    timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500+n)) bench.py --gpus $n --steps 1000 --warmup 10 > gpurun_out/scale_n$n.json 2> gpurun_out/scale_n$n.err
    timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600+n)) profiles/bench_config5.py --steps 30 2>/dev/null | grep workload > gpurun_out/config5_n$n.json
  fi
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/scale_n$n.json')); print('N=$n weak: value %.1f evals/s  e2e %.1f  ms/step %.4f  frac %.3f' % (d['value'], d['e2e']['value'], d['ms_per_step'], d['roofline']['frac']))
except Exception as e: print('N=$n bench failed', e); print(open('gpurun_out/scale_n$n.err').read()[-1500:])
try:
    d=json.load(open('gpurun_out/config5_n$n.json')); print('N=$n config5 strong: %.3f ms/eval  %.1f evals/s  aggregate %.0f GB/s  per-GPU frac %.3f' % (d['ms_per_eval'], d['evals_per_s'], d['aggregate_GBps'], d['per_gpu_roofline_frac']))
except Exception as e: print('N=$n config5 failed', e)
PY
done
