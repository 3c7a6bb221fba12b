"""Kernel-configuration sweep of the fused fg kernel (run under gpurun).  Prints per-config kernel time and
fraction of the measured HBM roofline; used to fit the scoring model in csrc/sfh_api.cu::choose_config."""
import itertools, json, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sfh_b200 as S

PEAK = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"]

def run(nb, nt, dtype, combos, reps=30, label=""):
    rng = np.random.default_rng(0)
    x = 100 * rng.random(nt)
    es = np.dtype(dtype).itemsize
    bytes_alg = nb * nt * es + nb * 8 + 16 * nt + 8
    print(f"## {label} nb={nb} nt={nt} dtype={np.dtype(dtype).name} bytes_alg={bytes_alg/1e6:.1f} MB", flush=True)
    for combo in combos:
        nw, bt, c = combo[:3]
        var = combo[3] if len(combo) > 3 else 0
        try:
            ds = S.DeviceStack.synthetic(nb, nt, dtype, seed=1, scale=1.0, x_true=x, tile_bins=bt, cluster=c, consumer_warps=nw, variant=var)
        except Exception as e:
            print(f"nw={nw:2d} bt={bt:3d} c={c:2d} var={var}  unavailable ({str(e)[:60]})"); continue
        i = ds.info()
        if not i.fused:
            print(f"nw={nw:2d} bt={bt:3d} c={c:2d} var={var}  not fused"); ds.close(); continue
        ds.time_fg(x, reps=5, flush_l2=False)
        ms, msk = ds.time_fg(x, reps=reps, flush_l2=(nb * nt * es < 400e6))
        print(f"nw={i.consumer_warps:2d} variant={i.variant} bt={i.tile_bins:3d} c={i.cluster:2d} kt={i.chunks_per_tile:2d} ring={i.ring_slots:2d} ncl={i.n_clusters:3d}  "
              f"eval={ms*1e3:8.1f} us kernel={msk*1e3:8.1f} us  {bytes_alg/msk/1e6:7.0f} GB/s  frac={bytes_alg/msk/1e6/PEAK:.3f}", flush=True)
        ds.close()

which = sys.argv[1] if len(sys.argv) > 1 else "c3"
if which == "c3":
    combos = [(8, 16, 4, 1), (16, 16, 2, 1), (8, 16, 4, 2), (8, 8, 2, 2), (8, 32, 8, 2), (12, 16, 4, 2), (12, 32, 8, 2), (12, 8, 2, 2),
              (12, 16, 8, 2), (8, 16, 8, 2), (12, 8, 4, 2), (0, 0, 0, 0)]
    run(60000, 2400, np.float64, combos, label="config3")
elif which == "panel":
    combos = [(8, 16, 4, 1), (16, 16, 2, 1), (8, 8, 2, 1), (16, 8, 1, 1), (8, 8, 4, 1), (8, 32, 8, 1), (16, 32, 4, 1), (16, 64, 8, 1), (8, 64, 16, 1),
              (8, 8, 2, 2), (12, 8, 2, 2), (8, 16, 4, 2), (12, 16, 4, 2)]
    run(60000, 2400, np.float64, combos, label="config3")
elif which == "rt":
    run(60000, 2400, np.float64, [(0, 0, 0, 0)], label="config3 auto")
    run(40000, 500, np.float64, [(0, 0, 0, 0), (8, 64, 1, 2), (8, 32, 1, 2), (12, 64, 1, 2), (12, 32, 1, 2), (8, 64, 2, 2), (8, 64, 4, 1)], label="config2 stack")
    run(10000, 100, np.float64, [(0, 0, 0, 0), (8, 64, 1, 2), (12, 64, 1, 2), (8, 64, 1, 1)], reps=50, label="config1")
    run(11250, 2000, np.float64, [(0, 0, 0, 0)], label="reference CI shape f64")
    run(11250, 2000, np.float32, [(0, 0, 0, 0)], label="reference CI shape f32")
    run(125000, 10000, np.float32, [(0, 0, 0, 0), (8, 16, 8, 2), (12, 16, 8, 2), (8, 32, 16, 2), (12, 32, 16, 2), (16, 32, 8, 1)], reps=10, label="config5 shard (1/8)")
elif which == "all":
    run(60000, 2400, np.float64, [(0, 0, 0)], label="config3 auto")
    run(10000, 100, np.float64, [(0, 0, 0), (16, 16, 1), (8, 16, 1), (8, 32, 1), (16, 64, 1), (8, 64, 1)], reps=50, label="config1")
    run(40000, 500, np.float64, [(0, 0, 0), (16, 64, 2), (8, 64, 4), (8, 32, 2), (16, 32, 1), (8, 16, 1), (16, 16, 1)], label="config2 stack")
    run(11250, 2000, np.float64, [(0, 0, 0)], label="reference CI shape f64")
    run(11250, 2000, np.float32, [(0, 0, 0)], label="reference CI shape f32")
    run(125000, 10000, np.float32, [(0, 0, 0), (16, 32, 8), (8, 32, 16), (16, 64, 16), (8, 64, 16), (16, 128, 16)], reps=10, label="config5 shard (1/8)")
