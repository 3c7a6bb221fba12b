"""Native driver loops and the stack container, measured on the device (round-1 closing check).

  * fit_sfh on BASELINE config 3 (200x300 bins x 2400 templates, PowerLawMZR + GaussianDispersion): the scipy-driven host loop
    vs the library's own BFGS loop (sfh_fit_sfh_bfgs) around the SAME device evaluations -- wall time, evaluations, us per
    evaluation end to end, agreement of the answers;
  * fit_templates (log-space MAP + MLE) at 100 and 500 templates, both engines;
  * sfh_stack_save / sfh_stack_create_from_file of the 1.15 GB config-3 stack: GB/s each way, bit-identical fg! afterwards.
Prints one JSON object per measurement."""
import json, os, sys, tempfile, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sfh_b200 as S

def out(**kw): print(json.dumps(kw), flush=True)

rng = np.random.Generator(np.random.Philox(94823))
uA = np.linspace(10.1, 6.6, 60); uM = np.linspace(-2.5, 0.0, 40)
la = np.repeat(uA, 40); mh = np.tile(uM, 60)
R = rng.random(60) * 1e6
mz, dp = S.PowerLawMZR(1.0, -2.0, 6.0), S.GaussianDispersion(0.2)
xt = S.calculate_coeffs(mz, dp, R, la, mh)
ds3 = S.DeviceStack.synthetic(60000, 2400, np.float64, 94823, 1e-5, xt)
truth = np.concatenate([R, [1.0, -2.0, 0.2]])
d3 = ds3.download_data()
start = (S.PowerLawMZR(1.2, -2.2, 6.0), S.GaussianDispersion(0.25))
res = {}
for engine in ("scipy", "native", "scipy", "native"):          # first pair warms up (graph capture, page-locking)
    t0 = time.perf_counter()
    r = S.fit_sfh(*start, ds3, d3, la, mh, x0=R * 1.5, g_abstol=1e-6, engine=engine)
    res[engine] = (time.perf_counter() - t0, r)
for engine, (t, r) in res.items():
    nf = int(r["map"].result.nfev) + int(r["mle"].result.nfev)
    z = np.abs(r["map"].mu - truth) / r["map"].sigma
    out(what="fit_sfh MZR (MAP+MLE BFGS), config 3, g_abstol 1e-6", engine=engine, wall_s=t, fevals=nf, us_per_eval_end_to_end=1e6 * t / nf,
        iterations=[int(r["map"].result.nit), int(r["mle"].result.nit)], converged=[bool(r["map"].result.success), bool(r["mle"].result.success)],
        frac_params_within_3sigma=float(np.mean(z < 3)), alpha_beta_sigma=[float(v) for v in r["mle"].mu[-3:]])
for ag in (2, 2):                                              # first-trial step from the previous decrease instead of InitialStatic
    t0 = time.perf_counter()
    r = S.fit_sfh(*start, ds3, d3, la, mh, x0=R * 1.5, g_abstol=1e-6, engine="native", alphaguess=ag)
    t = time.perf_counter() - t0
nf = int(r["map"].result.nfev) + int(r["mle"].result.nfev)
out(what="fit_sfh MZR config 3, native loop with alphaguess = 2 (previous-decrease first step)", wall_s=t, fevals=nf, us_per_eval_end_to_end=1e6 * t / nf,
    iterations=[int(r["map"].result.nit), int(r["mle"].result.nit)], converged=[bool(r["map"].result.success), bool(r["mle"].result.success)],
    max_rel_diff_mle_mu_vs_default=float(np.max(np.abs(r["mle"].mu / res["native"][1]["mle"].mu - 1))))
a, b = res["scipy"][1], res["native"][1]
out(what="fit_sfh native vs scipy", max_rel_diff_map_mu=float(np.max(np.abs(a["map"].mu / b["map"].mu - 1))),
    max_rel_diff_mle_mu=float(np.max(np.abs(a["mle"].mu / b["mle"].mu - 1))),
    median_sigma_ratio_map=float(np.median(b["map"].sigma[:60] / a["map"].sigma[:60])))

# ---- BASELINE config 1: fit_templates_lbfgsb, 100x100 bins x 100 templates -------------------------------------------------
g = np.random.Generator(np.random.Philox(58392))
x1 = 100 * g.random(100)
ds1 = S.DeviceStack.synthetic(10000, 100, np.float64, 58392, 1.0, x1)
d1 = ds1.download_data()
tt = {}
for engine in ("scipy", "native", "scipy", "native"):
    t0 = time.perf_counter()
    f, xf = S.fit_templates_lbfgsb(ds1, d1, x0=np.ones(100), engine=engine)
    tt[engine] = (time.perf_counter() - t0, f, xf)
out(what="fit_templates_lbfgsb config 1 (10000 bins x 100 templates)", scipy_s=tt["scipy"][0], native_s=tt["native"][0],
    speedup=tt["scipy"][0] / tt["native"][0], rel_diff_coeffs=float(np.linalg.norm(tt["native"][2] - tt["scipy"][2]) / np.linalg.norm(tt["scipy"][2])),
    rel_diff_nlogL=float(abs(tt["native"][1] - tt["scipy"][1]) / abs(tt["scipy"][1])))
ds1.close()

for nb, nt in ((10000, 100), (40000, 500)) + (((60000, 2400),) if os.environ.get("SFH_BENCH_DEVICE_HESSIAN") else ()):
    g = np.random.Generator(np.random.Philox(58392))
    x = 100 * g.random(nt)
    ds = S.DeviceStack.synthetic(nb, nt, np.float64, 58392, 1.0, x)
    d = ds.download_data()
    tt = {}
    # (scipy's dense BFGS update costs ~0.4 s per iteration at 2400 variables: not run there; iterations capped for the big case)
    engines = ("scipy", "native", "scipy", "native") if nt < 1000 else ("native",)
    kw = {} if nt < 1000 else {"iterations": 60}
    for engine in engines:
        t0 = time.perf_counter()
        r = S.fit_templates(ds, d, x0=np.ones(nt), engine=engine, **kw)
        tt[engine] = (time.perf_counter() - t0, r)
    for engine, (t, r) in tt.items():
        nf = int(r["map"].result.nfev) + int(r["mle"].result.nfev)
        out(what=f"fit_templates (MAP+MLE BFGS) {nb} bins x {nt} templates", engine=engine, wall_s=t, fevals=nf, us_per_eval_end_to_end=1e6 * t / nf,
            converged=[bool(r["map"].result.success), bool(r["mle"].result.success)],
            rel_err_vs_truth=float(np.linalg.norm(r["mle"].mu - x) / np.linalg.norm(x)))
    if os.environ.get("SFH_BENCH_DEVICE_HESSIAN"):               # experimental: the inverse Hessian kept in HBM
        for _ in range(2 if nt < 1000 else 1):
            t0 = time.perf_counter()
            r = S.fit_templates(ds, d, x0=np.ones(nt), engine="native", device_hessian=True, **kw)
            t = time.perf_counter() - t0
        nf = int(r["map"].result.nfev) + int(r["mle"].result.nfev)
        out(what=f"fit_templates (MAP+MLE BFGS) {nb} bins x {nt} templates", engine="native + device-resident inverse Hessian", wall_s=t, fevals=nf,
            us_per_eval_end_to_end=1e6 * t / nf, rel_diff_mle_vs_host_hessian=float(np.linalg.norm(r["mle"].mu - tt["native"][1]["mle"].mu) / np.linalg.norm(x)))
    if "scipy" in tt:
        out(what=f"fit_templates native vs scipy, {nt} templates",
                rel_diff_mle=float(np.linalg.norm(tt["native"][1]["mle"].mu - tt["scipy"][1]["mle"].mu) / np.linalg.norm(x)))
    ds.close()

# ---- container: the 1.15 GB config-3 stack to a file and back ---------------------------------------------------------
tmp = tempfile.mkdtemp(dir=os.environ.get("SFH_TMPDIR", None))
path = os.path.join(tmp, "config3.sfh")
f0, G0, _ = ds3.eval_fg(xt)
t0 = time.perf_counter(); ds3.save(path, logAge=la, MH=mh, hess_shape=(200, 300)); t_save = time.perf_counter() - t0
size = os.path.getsize(path)
t0 = time.perf_counter(); ds4 = S.DeviceStack.from_file(path); t_load = time.perf_counter() - t0
t0 = time.perf_counter(); sh = S.DeviceStack.from_file(path, rows=(7500, 15000)); t_shard = time.perf_counter() - t0
t0 = time.perf_counter()
with S.SFHFile(path) as f: f.verify()
t_verify = time.perf_counter() - t0
f1, G1, _ = ds4.eval_fg(xt)
out(what="stack container, config-3 stack", file_bytes=size, save_s=t_save, save_GBps=size / t_save / 1e9, load_s=t_load, load_GBps=size / t_load / 1e9,
    load_one_eighth_shard_s=t_shard, verify_s=t_verify, verify_GBps=size / t_verify / 1e9, bit_identical_fg=bool(f0 == f1 and np.array_equal(G0, G1)))
os.remove(path); os.rmdir(tmp)
