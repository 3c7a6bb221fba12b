"""Native multi-chain NUTS (sfh_sample_sfh_nuts / sfh_hmc_sample_nuts) vs the Python coroutine engine on the device
(prepared in round 1 after the GPU budget was spent; first thing to run in round 2).

  * tsample_sfh on BASELINE config 3/4 (2400-template MZR stack, dim 63): 8 / 32 / 64 chains x 20 draws, max_depth 5 --
    engine="host" batched (coroutines + one sfh_eval_fg_hier_batched pass per round) vs engine="native" (chain threads inside
    the library around the same pass): wall time, evaluations, share of the time spent outside the device pass;
  * hmc_sample on 60000 bins x 2400 templates, 8 / 16 chains, both engines.
Prints one JSON object per measurement."""
import json, os, sys, time
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
d3 = ds3.download_data()
fit = S.fit_sfh(S.PowerLawMZR(1.2, -2.2, 6.0), S.GaussianDispersion(0.25), ds3, d3, la, mh, x0=R * 1.5, g_abstol=1e-6, engine="native")
for nchains in (8, 32, 64):
    tt = {}
    for engine in ("host", "native", "host", "native"):                 # first pair warms up
        t0 = time.perf_counter()
        r = S.tsample_sfh(fit, ds3, d3, la, mh, 20 * nchains, eps=0.05, rng=np.random.default_rng(4), chain_length=20, max_depth=5, engine=engine)
        tt[engine] = (time.perf_counter() - t0, r)
    pm = {e: tt[e][1]["posterior_matrix"] for e in tt}
    out(what=f"tsample_sfh {nchains} chains x 20 NUTS draws (max_depth 5), config-3 stack, dim 63", host_batched_s=tt["host"][0],
        native_s=tt["native"][0], speedup=tt["host"][0] / tt["native"][0],
        mean_rel_diff_R=float(np.max(np.abs(pm["native"][:60].mean(axis=1) / pm["host"][:60].mean(axis=1) - 1))),
        shapes=[list(pm["host"].shape), list(pm["native"].shape)])
for nchains in (8, 16):
    tt = {}
    for engine in ("host", "native", "host", "native"):
        t0 = time.perf_counter()
        o = S.hmc_sample(ds3, d3, 12, nchains=nchains, nwarmup=12, rng=np.random.default_rng(5), x0=xt, max_depth=4, engine=engine)
        tt[engine] = (time.perf_counter() - t0, o)
    out(what=f"hmc_sample {nchains} chains, 12 warm-up + 12 draws, max_depth 4, 60000 bins x 2400 templates", host_batched_s=tt["host"][0],
        native_s=tt["native"][0], speedup=tt["host"][0] / tt["native"][0], all_positive=bool(np.all(tt["native"][1] > 0)))
