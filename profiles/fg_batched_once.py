"""One sfh_eval_fg_batched call per C (after one warm-up) — run under `ncu --metrics gpu__time_duration.sum` for the launch list."""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sfh_b200 as S

nb, nt = 60000, 2400
rng = np.random.default_rng(3)
x = 100 * rng.random(nt)
ds = S.DeviceStack.synthetic(nb, nt, np.float64, seed=3, scale=1.0, x_true=x)
for C in (int(a) for a in (sys.argv[1:] or ["8", "64"])):
    X = np.asfortranarray(x[:, None] * (1 + 0.05 * rng.standard_normal((nt, C))))
    ds.eval_fg_batched(X)
    ds.eval_fg_batched(X)
