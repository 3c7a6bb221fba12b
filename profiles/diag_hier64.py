import sys, os, time, json
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sfh_b200 as S
from sfh_b200 import hierarchical as H
rng = np.random.Generator(np.random.Philox(94823))
uA = np.linspace(10.1, 6.6, 60); uM = np.linspace(-2.5, 0.0, 40)
la = np.repeat(uA, 40); mh = np.tile(uM, 60)
R = rng.random(60) * 1e6
mz, dp = S.PowerLawMZR(1.0, -2.0, 6.0), S.GaussianDispersion(0.2)
xt = S.calculate_coeffs(mz, dp, R, la, mh)
ds3 = S.DeviceStack.synthetic(60000, 2400, np.float64, 94823, 1e-5, xt)
d3 = ds3.download_data()
for Cn in (32, 48, 56, 64):
    V = np.concatenate([R, [1.0, -2.0, 0.2]])[:, None] * (1 + 0.01 * rng.standard_normal((63, Cn)))
    H.fg_batched_(mz, dp, V, ds3, d3, la, mh)
    t0 = time.perf_counter()
    for _ in range(5): H.fg_batched_(mz, dp, V, ds3, d3, la, mh)
    th = (time.perf_counter() - t0) / 5
    X = np.asfortranarray(np.stack([S.calculate_coeffs(mz, dp, V[:60, c], la, mh) for c in range(Cn)], axis=1))
    ds3.eval_fg_batched(X)
    t0 = time.perf_counter()
    for _ in range(5): ds3.eval_fg_batched(X)
    tf = (time.perf_counter() - t0) / 5
    print(json.dumps({"C": Cn, "hier_batched_ms": th * 1e3, "flat_batched_ms": tf * 1e3}), flush=True)
