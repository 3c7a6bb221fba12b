"""Median wall clock per host-synchronous call (packets on / off interleaved) for one library build: SFH_LIB=... python bench_e2e_quick.py"""
import ctypes as C, os, sys, time, json
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sfh_b200 as S
L = S._lib
dp = C.POINTER(C.c_double)

def block(fn, n, warm=15):
    for _ in range(warm): fn()
    ts = np.empty(n)
    for i in range(n):
        t0 = time.perf_counter(); fn(); ts[i] = time.perf_counter() - t0
    return ts * 1e6

def flat(nb, nt, label, n=300):
    x = 100 * np.random.default_rng(0).random(nt)
    ds = S.DeviceStack.synthetic(nb, nt, np.float64, 1, 1.0, x)
    ctx = ds.ctx(); G = np.empty(nt); nl = C.c_double(); xx = np.ascontiguousarray(x)
    call = lambda: L.lib.sfh_eval_fg(ctx.handle, xx.ctypes.data_as(dp), C.byref(nl), G.ctypes.data_as(dp), None)
    callf = lambda: L.lib.sfh_eval_fg(ctx.handle, xx.ctypes.data_as(dp), C.byref(nl), None, None)
    res = {}
    for rnd in range(3):
        for k in (0, 1):
            os.environ["SFH_HOST_PACKETS"] = str(k)
            res.setdefault(("fg", k), []).append(block(call, n))
            res.setdefault(("f_only", k), []).append(block(callf, n))
    out = {"lib": os.path.basename(os.environ.get("SFH_LIB", "in-tree")), "case": label}
    for (what, k), ts in res.items():
        out[f"{what}_{'packets' if k else 'sync'}_median_us"] = round(float(np.median(np.concatenate(ts))), 2)
    print(json.dumps(out), flush=True)

flat(10000, 100, "config1", n=500)
flat(60000, 2400, "config3")
