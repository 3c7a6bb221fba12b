"""Single-process multi-GPU (sfh_group_*): BASELINE config 5 (10^6 bins x 10^4 templates F32, 40 GB) and config 3 (weak) sharded over
the visible GPUs from ONE process.  usage: bench_group.py [ndev] [config5|config3]"""
import json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sfh_b200 as S

ndev = int(sys.argv[1]) if len(sys.argv) > 1 else S.device_count()
which = sys.argv[2] if len(sys.argv) > 2 else "config5"
if which == "config5":
    nb, nt, dt, seed, scale = 1000 * 1000, 10000, np.float32, 58392, 1.0
    x = np.random.Generator(np.random.Philox(seed)).random(nt)
elif which == "config5half":      # the shard one of two GPUs holds, on however many GPUs were asked for
    nb, nt, dt, seed, scale = 500 * 1000, 10000, np.float32, 58392, 1.0
    x = np.random.Generator(np.random.Philox(seed)).random(nt)
else:
    nb, nt, dt, seed, scale = 60000 * ndev, 2400, np.float64, 94823, 1e-5
    x = np.random.Generator(np.random.Philox(seed)).random(nt) * 1e4
t0 = time.perf_counter()
g = S.DeviceStackGroup.synthetic(nb, nt, dt, seed, scale, x, ndev=ndev)
t_build = time.perf_counter() - t0
xe = x * 1.02
ms_dev, _ = g.time_fg(xe, reps=20)
for _ in range(5):
    g.eval_fg(xe)
n = 50
t0 = time.perf_counter()
for _ in range(n):
    nl, G, _ = g.eval_fg(xe)
wall_ms = (time.perf_counter() - t0) / n * 1e3
t0 = time.perf_counter()
for _ in range(n):
    g.eval_fg(xe, want_G=False)
wall_f_ms = (time.perf_counter() - t0) / n * 1e3
i = g.infos()
bytes_shard = (i[0].row_end - i[0].row_begin) * nt * np.dtype(dt).itemsize
print(json.dumps({"what": f"{which} from ONE process over {ndev} GPU(s)", "nb": nb, "nt": nt, "ndev": ndev, "build_s": t_build,
                  "ms_per_eval_device": ms_dev, "ms_per_eval_wall_host_api": wall_ms, "ms_per_eval_wall_logl_only": wall_f_ms, "per_gpu_GBps_device": bytes_shard / ms_dev / 1e6,
                  "neg_logL": nl, "tiling": [i[0].variant, i[0].tile_bins, i[0].cluster, i[0].chunks_per_tile, i[0].ring_slots, i[0].n_clusters]}))
g.close()
