import sys, time, numpy as np
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import sfh_b200 as S
from conftest import make_hier_problem
p = make_hier_problem(nj=14, nk=11, nb=3000)
mh, dp = S.PowerLawMZR(1.0, -2.0, 6.0), S.GaussianDispersion(0.2)
R = p["R"]; xt = S.calculate_coeffs(mh, dp, R, p["logAge"], p["MH"])
data = np.random.default_rng(5).poisson(p["M"] @ xt).astype(np.float64)
ds = S.DeviceStack(p["M"], data)
rng = np.random.default_rng(9)
for Cn in (16, 32, 33, 64, 70):
    V = np.concatenate([R, [1.0,-2.0,0.2]])[:, None] * (1 + 0.05 * rng.standard_normal((17, Cn)))
    nl, G = S.hierarchical.fg_batched_(mh, dp, V, ds, data, p["logAge"], p["MH"])
    bad = 0
    for c in range(Cn):
        g1 = np.empty(17); n1 = S.fg_(True, g1, mh, dp, V[:, c], ds, data, None, p["logAge"], p["MH"])
        if not (abs(nl[c]-n1) <= 1e-12*abs(n1) and np.allclose(G[:, c], g1, rtol=1e-8, atol=1e-9*np.abs(g1).max())): bad += 1
    t0=time.perf_counter()
    for _ in range(20): S.hierarchical.fg_batched_(mh, dp, V, ds, data, p["logAge"], p["MH"])
    print(Cn, 'bad', bad, 'ms/call', (time.perf_counter()-t0)/20*1e3)
