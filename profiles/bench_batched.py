"""K6 batched-walker benchmark (BASELINE config 2: 200x200 bins x 500 templates F64, 1024 walkers per step)."""
import ctypes as C, os, sys, time, json
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import sfh_b200 as S
L = S._lib

def run(nb, nt, W, dtype=np.float64, reps=10):
    rng = np.random.default_rng(0)
    x = 100 * rng.random(nt)
    ds = S.DeviceStack.synthetic(nb, nt, dtype, seed=2, scale=1.0, x_true=x)
    stream = torch.cuda.Stream(); torch.cuda.set_stream(stream)
    ctx = ds.new_ctx(stream.cuda_stream)
    X = np.asfortranarray(np.maximum(0.0, x[:, None] + rng.standard_normal((nt, W))))
    dX = torch.tensor(np.ascontiguousarray(X.T), dtype=torch.float64, device="cuda")  # (W, nt) row-major == (nt, W) col-major
    dL = torch.zeros(W, dtype=torch.float64, device="cuda")
    for _ in range(3):
        L.check(L.lib.sfh_enqueue_logl_batched(ctx.handle, dX.data_ptr(), W, dL.data_ptr()))
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(reps):
        L.check(L.lib.sfh_enqueue_logl_batched(ctx.handle, dX.data_ptr(), W, dL.data_ptr()))
    e1.record(stream); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    flops = 2.0 * nb * nt * W
    # host-API (e2e) timing
    out = np.empty(W)
    t0 = time.perf_counter()
    for _ in range(reps):
        got = ds.eval_logl_batched(X)
    e2e = (time.perf_counter() - t0) / reps * 1e3
    print(json.dumps({"nb": nb, "nt": nt, "W": W, "dtype": np.dtype(dtype).name, "ms_per_batch": ms, "walker_evals_per_s": W / ms * 1e3,
                      "fp64_tflops": flops / ms / 1e9, "e2e_ms_per_batch": e2e, "e2e_walker_evals_per_s": W / e2e * 1e3}), flush=True)
    return ds, X, got

def run_mcmc_engines(nb=40000, nt=500, W=1024, nsteps=20):
    """mcmc_sample at BASELINE config 2: the device-resident ensemble sampler vs numpy proposals around one K6 call per half."""
    from sfh_b200 import solvers as V
    rng = np.random.default_rng(9)
    x = 100 * rng.random(nt)
    ds = S.DeviceStack.synthetic(nb, nt, np.float64, seed=9, scale=1.0, x_true=x)
    data = ds.download_data()
    X0 = np.maximum(0.0, x[:, None] + rng.standard_normal((nt, W)))
    for engine in ("host", "device"):
        V.mcmc_sample(ds, data, X0, 2, rng=np.random.default_rng(1), engine=engine)
        t0 = time.perf_counter()
        chain, lps, acc = V.mcmc_sample(ds, data, X0, nsteps, rng=np.random.default_rng(1), engine=engine)
        t = time.perf_counter() - t0
        print(json.dumps({"mcmc_sample": engine, "nb": nb, "nt": nt, "W": W, "nsteps": nsteps, "s": t, "ms_per_step": t / nsteps * 1e3,
                          "walker_evals_per_s": W * nsteps / t, "accept": acc}), flush=True)
    t0 = time.perf_counter()
    ds.mcmc_run(X0, nsteps, 1, 2.0, seed=5, store=False)
    t = time.perf_counter() - t0
    print(json.dumps({"mcmc_run_no_store": True, "ms_per_step": t / nsteps * 1e3, "walker_evals_per_s": W * nsteps / t}), flush=True)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "mcmc":
        run_mcmc_engines(); sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "k6":
        run(40000, 500, 1024); run(40000, 500, 512); run(60000, 2400, 256); sys.exit(0)
    ds, X, got = run(40000, 500, 1024)
    # spot parity of 3 walkers against the fused single-vector path
    for w in (0, 511, 1023):
        nl, _, _ = ds.eval_fg(X[:, w], want_G=False)
        assert abs(got[w] + nl) <= 1e-12 * abs(nl), (got[w], nl)
    run(40000, 500, 512)
    run(40000, 500, 128)
    run(60000, 2400, 256)


def run_fg_batched(nb, nt, Cs=(1, 8, 16, 32, 64), reps=5):
    """sfh_eval_fg_batched (multi-chain fg): host-API time per call and per chain-evaluation, vs the single-vector path."""
    rng = np.random.default_rng(3)
    x = 100 * rng.random(nt)
    ds = S.DeviceStack.synthetic(nb, nt, np.float64, seed=3, scale=1.0, x_true=x)
    ds.eval_fg(x)
    t0 = time.perf_counter()
    for _ in range(20): ds.eval_fg(x)
    t_single = (time.perf_counter() - t0) / 20
    for C in Cs:
        X = np.asfortranarray(x[:, None] * (1 + 0.05 * rng.standard_normal((nt, C))))
        ds.eval_fg_batched(X)
        t0 = time.perf_counter()
        for _ in range(reps): ds.eval_fg_batched(X)
        t = (time.perf_counter() - t0) / reps
        print(json.dumps({"fg_batched": True, "nb": nb, "nt": nt, "C": C, "ms_per_call": t * 1e3, "us_per_chain_eval": t / C * 1e6,
                          "single_vector_us_per_eval": t_single * 1e6, "fp64_tflops": 4.0 * nb * nt * C / t / 1e12}), flush=True)


def run_mcmc_engines(nb=40000, nt=500, W=1024, nsteps=20):
    """mcmc_sample at BASELINE config 2: the device-resident ensemble sampler vs numpy proposals around one K6 call per half."""
    from sfh_b200 import solvers as V
    rng = np.random.default_rng(9)
    x = 100 * rng.random(nt)
    ds = S.DeviceStack.synthetic(nb, nt, np.float64, seed=9, scale=1.0, x_true=x)
    data = ds.download_data()
    X0 = np.maximum(0.0, x[:, None] + rng.standard_normal((nt, W)))
    for engine in ("host", "device"):
        V.mcmc_sample(ds, data, X0, 2, rng=np.random.default_rng(1), engine=engine)
        t0 = time.perf_counter()
        chain, lps, acc = V.mcmc_sample(ds, data, X0, nsteps, rng=np.random.default_rng(1), engine=engine)
        t = time.perf_counter() - t0
        print(json.dumps({"mcmc_sample": engine, "nb": nb, "nt": nt, "W": W, "nsteps": nsteps, "s": t, "ms_per_step": t / nsteps * 1e3,
                          "walker_evals_per_s": W * nsteps / t, "accept": acc}), flush=True)
    t0 = time.perf_counter()
    ds.mcmc_run(X0, nsteps, 1, 2.0, seed=5, store=False)
    t = time.perf_counter() - t0
    print(json.dumps({"mcmc_run_no_store": True, "ms_per_step": t / nsteps * 1e3, "walker_evals_per_s": W * nsteps / t}), flush=True)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "mcmc":
        run_mcmc_engines(); sys.exit(0)
    run_fg_batched(60000, 2400)
