"""A/B of the latency switches in ONE process (the library reads SFH_PDL_EARLY / SFH_HOST_PACKETS on every call):
per-call wall clock of sfh_eval_fg / sfh_eval_fg_hier (median and mean over blocks of calls, modes interleaved round-robin so
that clock / power drift hits all of them alike) and the device-resident loop (CUDA events)."""
import ctypes as C, os, sys, time, json
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sfh_b200 as S
import torch
L = S._lib
dp = C.POINTER(C.c_double)
MODES = [(m, k) for k in (0, 1) for m in (0, 1, 2, 4, 6, 7)]
ROUNDS = int(os.environ.get("ROUNDS", "4"))

def setmode(m, k):
    os.environ["SFH_PDL_EARLY"] = str(m); os.environ["SFH_HOST_PACKETS"] = str(k)

def block(fn, n, warm=15):
    for _ in range(warm): fn()
    ts = np.empty(n)
    for i in range(n):
        t0 = time.perf_counter(); fn(); ts[i] = time.perf_counter() - t0
    return ts * 1e6

def report(label, res):
    for mode, ts in res.items():
        t = np.concatenate(ts)
        print(json.dumps({"case": label, "pdl_early": mode[0], "host_packets": mode[1], "median_us": round(float(np.median(t)), 2),
                          "mean_us": round(float(t.mean()), 2), "p10_us": round(float(np.percentile(t, 10)), 2), "p90_us": round(float(np.percentile(t, 90)), 2), "calls": int(t.size)}), flush=True)

def flat(nb, nt, label, dt=np.float64, n=300, modes=MODES):
    x = 100 * np.random.default_rng(0).random(nt)
    ds = S.DeviceStack.synthetic(nb, nt, dt, 1, 1.0, x)
    ctx = ds.ctx(); G = np.empty(nt); nl = C.c_double(); xx = np.ascontiguousarray(x)
    call = lambda: L.lib.sfh_eval_fg(ctx.handle, xx.ctypes.data_as(dp), C.byref(nl), G.ctypes.data_as(dp), None)
    res = {m: [] for m in modes}
    for _ in range(ROUNDS):
        for m in modes:
            setmode(*m); res[m].append(block(call, n))
    report(label, res)
    # device-resident loop: back-to-back enqueued evaluations, events on the context's stream
    st = torch.cuda.Stream(); ctx2 = ds.new_ctx(st.cuda_stream)
    d_x = torch.tensor(x, dtype=torch.float64, device="cuda"); d_out = torch.zeros(1 + nt, dtype=torch.float64, device="cuda")
    steps = 400
    dev = {m: [] for m in modes if m[1] == 1}
    for _ in range(ROUNDS):
        for m in dev:
            setmode(*m)
            for _ in range(10): L.lib.sfh_enqueue_fg(ctx2.handle, d_x.data_ptr(), d_out.data_ptr(), 1)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(st)
            for _ in range(steps): L.lib.sfh_enqueue_fg(ctx2.handle, d_x.data_ptr(), d_out.data_ptr(), 1)
            e1.record(st); st.synchronize()
            dev[m].append(e0.elapsed_time(e1) * 1e3 / steps)
    for m, v in dev.items():
        print(json.dumps({"case": label + " device loop", "pdl_early": m[0], "us_per_step": [round(t, 2) for t in v]}), flush=True)

def hier(nb, nj, nk, label, n=300, modes=MODES):
    rng = np.random.default_rng(1)
    la = np.repeat(np.linspace(10.1, 6.6, nj), nk); mh = np.tile(np.linspace(-2.5, 0, nk), nj)
    R = rng.random(nj) * 1e6
    mz, dpm = S.PowerLawMZR(1.0, -2.0, 6.0), S.GaussianDispersion(0.2)
    xt = S.calculate_coeffs(mz, dpm, R, la, mh)
    ds = S.DeviceStack.synthetic(nb, nj * nk, np.float64, 2, 1e-5, xt)
    v = np.concatenate([R, [1.0, -2.0, 0.2]]) * 1.03
    G = np.empty(nj + 3)
    S.fg_(True, G, mz, dpm, v, ds, None, None, la, mh)     # binds the grid
    ctx = ds.ctx(); nl = C.c_double(); fx = mz.fixed(); free = np.array([1, 1, 1, 0], dtype=np.uint8)
    call = lambda: L.lib.sfh_eval_fg_hier(ctx.handle, 0, fx.ctypes.data_as(dp), 0, v.ctypes.data_as(dp), free.ctypes.data_as(C.POINTER(C.c_uint8)), C.byref(nl), G.ctypes.data_as(dp))
    res = {m: [] for m in modes}
    for _ in range(ROUNDS):
        for m in modes:
            setmode(*m); res[m].append(block(call, n))
    report(label, res)

flat(10000, 100, "config1 100x100 x 100", n=600)
flat(40000, 500, "config2 40000 x 500", n=400)
flat(60000, 2400, "config3 60000 x 2400")
hier(10000, 21, 26, "mzr_test 10000 x 546 hier", n=400)
hier(60000, 60, 40, "config3 hier")
