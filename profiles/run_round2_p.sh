#!/bin/bash
# Round 2, call p (2 GPUs): the packet completion + early dependent release under the one-shot exchange: parity script on 2 ranks,
# group tests, bench at N = 2 with the finalize kernel's early release on (default) and off, one-process group timing.
mkdir -p gpurun_out
T="timeout -k 5"
$T 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29731 tests/mgpu_check.py 2>&1 | tail -3 | tee gpurun_out/r2p_mgpu_check.txt
$T 400 python -m pytest tests/test_gpu_group.py tests/test_gpu_multi.py -m gpu -x -q 2>&1 | tail -3 | tee gpurun_out/r2p_group_tests.txt
for m in 6 4; do
  SFH_PDL_EARLY=$m $T 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2974$m bench.py --gpus 2 --steps 2000 --warmup 10 --no-config5 2> gpurun_out/r2p_bench_n2_pdl$m.err > gpurun_out/r2p_bench_n2_pdl$m.json
done
SFH_PDL_EARLY=6 $T 400 python bench.py --steps 2000 --warmup 10 --no-config5 --no-cpu-baseline 2> gpurun_out/r2p_bench_n1.err > gpurun_out/r2p_bench_n1.json
python - <<'PY'
import json
for n in ('n1', 'n2_pdl6', 'n2_pdl4'):
    try:
        d = json.load(open(f'gpurun_out/r2p_bench_{n}.json'))
        print(n, 'value', round(d['value'], 1), 'ms', round(d['ms_per_step'], 5), 'e2e_ms', round(d['e2e']['ms_per_step'], 5), 'frac', round(d['roofline']['frac'], 4),
              'kernel_ms', [round(v, 4) for v in d['roofline']['kernel_ms_per_rank']], d['clocks'], 'launches', d['gpu_launches'], 'hier', round(d['fg_hier']['ms_per_eval'], 5), d.get('parity'))
    except Exception as e:
        print(n, 'FAILED', e)
PY
tail -3 gpurun_out/r2p_bench_n2_pdl6.err
$T 300 python profiles/bench_group.py 2 config3 2>&1 | tail -1 | tee gpurun_out/r2p_group.txt
