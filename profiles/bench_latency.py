"""Host-synchronous call latency of the reference-facing entry points (wall clock per call, Python ctypes caller).
Run twice: default (CUDA graphs + mapped result buffer) and SFH_NO_GRAPH=1."""
import ctypes as C, os, sys, time, json
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sfh_b200 as S
L = S._lib
dp = C.POINTER(C.c_double)

def time_call(fn, n=300, warm=20):
    for _ in range(warm): fn()
    t0 = time.perf_counter()
    for _ in range(n): fn()
    return (time.perf_counter() - t0) / n * 1e6

def flat(nb, nt, label, dt=np.float64):
    x = 100 * np.random.default_rng(0).random(nt)
    ds = S.DeviceStack.synthetic(nb, nt, dt, 1, 1.0, x)
    ctx = ds.ctx(); G = np.empty(nt); nl = C.c_double(); xx = np.ascontiguousarray(x)
    t_fg = time_call(lambda: L.lib.sfh_eval_fg(ctx.handle, xx.ctypes.data_as(dp), C.byref(nl), G.ctypes.data_as(dp), None))
    t_f = time_call(lambda: L.lib.sfh_eval_fg(ctx.handle, xx.ctypes.data_as(dp), C.byref(nl), None, None))
    ms, msk = ds.time_fg(x, reps=20, flush_l2=False)
    print(json.dumps({"case": label, "nb": nb, "nt": nt, "eval_fg_us": t_fg, "eval_f_only_us": t_f, "device_eval_us": ms * 1e3, "fused_kernel_us": msk * 1e3}), flush=True)

def hier(nb, nj, nk, label):
    rng = np.random.default_rng(1)
    la = np.repeat(np.linspace(10.1, 6.6, nj), nk); mh = np.tile(np.linspace(-2.5, 0, nk), nj)
    R = rng.random(nj) * 1e6
    mz, dpm = S.PowerLawMZR(1.0, -2.0, 6.0), S.GaussianDispersion(0.2)
    xt = S.calculate_coeffs(mz, dpm, R, la, mh)
    ds = S.DeviceStack.synthetic(nb, nj * nk, np.float64, 2, 1e-5, xt)
    v = np.concatenate([R, [1.0, -2.0, 0.2]]) * 1.03
    G = np.empty(nj + 3)
    t = time_call(lambda: S.fg_(True, G, mz, dpm, v, ds, None, None, la, mh), n=200)
    ctx = ds.ctx(); nl = C.c_double(); fx = mz.fixed(); free = np.array([1, 1, 1, 0], dtype=np.uint8)
    t_raw = time_call(lambda: L.lib.sfh_eval_fg_hier(ctx.handle, 0, fx.ctypes.data_as(dp), 0, v.ctypes.data_as(dp), free.ctypes.data_as(C.POINTER(C.c_uint8)), C.byref(nl), G.ctypes.data_as(dp)), n=300)
    print(json.dumps({"case": label, "nb": nb, "nt": nj * nk, "fg_hier_python_us": t, "sfh_eval_fg_hier_us": t_raw}), flush=True)

print("graphs:", "off" if os.environ.get("SFH_NO_GRAPH") == "1" else "on", flush=True)
flat(10000, 100, "config1 100x100 bins x 100")
flat(9801, 142, "notebook 99x99 x 142")
flat(40000, 500, "config2 stack")
flat(60000, 2400, "config3")
flat(125000, 10000, "config5 shard (1/8) F32", np.float32)
hier(10000, 21, 26, "mzr_test 100x100 x 546")
hier(60000, 60, 40, "config3 hier")
