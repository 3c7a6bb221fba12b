"""A few sfh_eval_fg_hier calls on the config-3 stack (run under ncu for the per-kernel launch list)."""
import sys, os, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sfh_b200 as S
rng = np.random.Generator(np.random.Philox(94823))
uA = np.linspace(10.1, 6.6, 60); uM = np.linspace(-2.5, 0.0, 40)
la = np.repeat(uA, 40); mh = np.tile(uM, 60)
R = rng.random(60) * 1e6
mz, dp = S.PowerLawMZR(1.0, -2.0, 6.0), S.GaussianDispersion(0.2)
xt = S.calculate_coeffs(mz, dp, R, la, mh)
ds3 = S.DeviceStack.synthetic(60000, 2400, np.float64, 94823, 1e-5, xt)
d3 = ds3.download_data()
v = np.concatenate([R, [1.0, -2.0, 0.2]]); G = np.empty(63)
for _ in range(3): ds3.eval_fg(xt * 1.01)     # flat sfh_eval_fg: upload kernel -> fused -> finalize
for _ in range(5): S.fg_(True, G, mz, dp, v, ds3, d3, None, la, mh)
t0 = time.perf_counter()
for _ in range(200): S.fg_(True, G, mz, dp, v, ds3, d3, None, la, mh)
print("us per hier eval", (time.perf_counter() - t0) / 200 * 1e6)
