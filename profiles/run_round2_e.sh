#!/bin/bash
# Round 2, call e (2 GPUs): full GPU suite incl. the two-process NCCL/one-shot checks and the single-process groups; chooser check;
# bench line at N = 1 and N = 2; config 5 from one process.
mkdir -p gpurun_out
T="timeout -k 5"
$T 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -12 | tee gpurun_out/r2e_gpu_tests.log
for shape in "60000 2400 float64" "40000 500 float64" "11250 2000 float64" "125000 10000 float32"; do
  $T 90 python profiles/one_config.py 0 0 0 0 50 $shape 2>&1 | tail -1
done | tee gpurun_out/r2e_auto_choice.txt
for cfg in "0 2 1 4" "0 4 1 4"; do $T 60 python profiles/one_config.py $cfg 200 2>&1 | tail -1; done | tee gpurun_out/r2e_config3_tiles.txt
$T 400 python bench.py --steps 1000 --warmup 10 2> gpurun_out/r2e_bench_n1.err > gpurun_out/r2e_bench_n1.json
$T 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29721 bench.py --gpus 2 --steps 1000 --warmup 10 2> gpurun_out/r2e_bench_n2.err > gpurun_out/r2e_bench_n2.json
python - <<'PY'
import json
for n in (1, 2):
    try:
        d = json.load(open(f'gpurun_out/r2e_bench_n{n}.json'))
        print(n, 'value', round(d['value'], 1), 'ms', round(d['ms_per_step'], 5), 'e2e_ms', round(d['e2e']['ms_per_step'], 5), 'frac', round(d['roofline']['frac'], 4),
              'kernel_ms', round(d['roofline']['kernel_ms'], 5), d['clocks'], 'launches', d['gpu_launches'])
        print('  hier', round(d['fg_hier']['ms_per_eval'], 5), 'config5', round(d['config5']['ms_per_eval'], 4), d['config5']['per_gpu'], d['config5'].get('parity'))
        print('  parity', d.get('parity'), d['config']['sharding'], d['config']['tile_bins'])
    except Exception as e:
        print(n, 'FAILED', e)
PY
tail -5 gpurun_out/r2e_bench_n2.err
$T 300 python profiles/bench_group.py 2 config5 2>&1 | tail -2 | tee gpurun_out/r2e_group.txt
$T 300 python profiles/bench_group.py 2 config3 2>&1 | tail -2 | tee -a gpurun_out/r2e_group.txt
$T 300 python profiles/bench_group.py 1 config3 2>&1 | tail -2 | tee -a gpurun_out/r2e_group.txt
