#!/bin/bash
# Round 2, call r (1 GPU): the driver's own invocations (bench.py --gpus 1 --steps 20 --warmup 5, both arms), smoke(), and the closing
# ncu evidence: launch list of the bench command, --set full captures of the stream kernel (config 3, config-5 shard) and of the
# finalize kernels.
mkdir -p gpurun_out
T="timeout -k 5"
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee gpurun_out/r2r_smoke.txt
for i in 1 2; do
  $T 600 python bench.py --gpus 1 --steps 20 --warmup 5 2> gpurun_out/r2r_bench_driver$i.err > gpurun_out/r2r_bench_driver$i.json
done
$T 600 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 2> gpurun_out/r2r_bench_reference.err > gpurun_out/r2r_bench_reference.json
$T 600 python bench.py --steps 2000 --warmup 10 --no-config5 --no-cpu-baseline 2> gpurun_out/r2r_bench_2000.err > gpurun_out/r2r_bench_2000.json
python - <<'PY'
import json
for n in ('driver1', 'driver2', '2000'):
    try:
        d = json.load(open(f'gpurun_out/r2r_bench_{n}.json'))
        print(n, 'value', round(d['value'], 1), 'ms', round(d['ms_per_step'], 5), 'e2e_ms', round(d['e2e']['ms_per_step'], 5), 'frac', round(d['roofline']['frac'], 4),
              'kernel_ms', round(d['roofline']['kernel_ms'], 5), d['clocks'], 'launches', d['gpu_launches'], 'hier', round(d['fg_hier']['ms_per_eval'], 5),
              'cpu', (d.get('cpu_baseline') or {}).get('value'), 'config5', (d.get('config5') or {}).get('ms_per_eval'))
    except Exception as e:
        print(n, 'FAILED', e)
try:
    d = json.load(open('gpurun_out/r2r_bench_reference.json')); print('reference', d['value'], d['cpu_baseline'])
except Exception as e:
    print('reference FAILED', e)
PY
$T 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_final_launches.csv python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-config5 > gpurun_out/r2r_ncu_launch.log 2>&1
$T 600 ncu --set full --clock-control none --import-source on -k regex:sfh_fg_fused2 -s 3 -c 1 -o gpurun_out/r2_final_config3 python profiles/one_config.py 0 0 0 4 5 > gpurun_out/r2r_ncu1.log 2>&1
$T 600 ncu --set full --clock-control none --import-source on -k regex:sfh_fg_fused2 -s 3 -c 1 -o gpurun_out/r2_final_config5shard python profiles/one_config.py 0 0 0 4 5 125000 10000 float32 > gpurun_out/r2r_ncu2.log 2>&1
$T 600 ncu --set full --clock-control none --import-source on -k regex:"sfh_finalize|sfh_copy_in|prologue2" -s 2 -c 8 -o gpurun_out/r2_final_small python profiles/hier_once.py > gpurun_out/r2r_ncu3.log 2>&1
ls -la gpurun_out | tail -8
