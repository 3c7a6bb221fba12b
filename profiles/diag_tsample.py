"""Where tsample_sfh's time goes: device calls vs the Python sampler, per chain count."""
import sys, os, time, json
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sfh_b200 as S
from sfh_b200 import solvers as V, hierarchical as H
rng = np.random.Generator(np.random.Philox(94823))
uA = np.linspace(10.1, 6.6, 60); uM = np.linspace(-2.5, 0.0, 40)
la = np.repeat(uA, 40); mh = np.tile(uM, 60)
R = rng.random(60) * 1e6
mz, dp = S.PowerLawMZR(1.0, -2.0, 6.0), S.GaussianDispersion(0.2)
xt = S.calculate_coeffs(mz, dp, R, la, mh)
ds3 = S.DeviceStack.synthetic(60000, 2400, np.float64, 94823, 1e-5, xt)
d3 = ds3.download_data()
res = S.fit_sfh(S.PowerLawMZR(1.2, -2.2, 6.0), S.GaussianDispersion(0.25), ds3, d3, la, mh, x0=R * 1.5, g_abstol=1e-6)
orig = H.HierarchicalOptimizer.logdensity_and_gradient_batched
stat = {"calls": 0, "evals": 0, "t": 0.0}
def timed(self, X):
    t0 = time.perf_counter(); r = orig(self, X); stat["t"] += time.perf_counter() - t0
    stat["calls"] += 1; stat["evals"] += X.shape[1]; return r
H.HierarchicalOptimizer.logdensity_and_gradient_batched = timed
for nch in (8, 16, 32, 64):
    for k in stat: stat[k] = 0
    t0 = time.perf_counter()
    S.tsample_sfh(res, ds3, d3, la, mh, 20 * nch, eps=0.05, rng=np.random.default_rng(4), chain_length=20, max_depth=5)
    t = time.perf_counter() - t0
    print(json.dumps({"chains": nch, "total_s": t, "device_call_s": stat["t"], "rounds": stat["calls"], "evals": stat["evals"],
                      "ms_per_round_device": stat["t"] / max(stat["calls"], 1) * 1e3, "us_per_eval_total": t / max(stat["evals"], 1) * 1e6}), flush=True)
