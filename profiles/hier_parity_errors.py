"""Measured parity errors of the hierarchical fg! on the device, in units of the backward-error scale the tests assert against
(tests/conftest.py: hier_grad_scale): for every metallicity model x (regular / shuffled / ragged grid) the maximum over the
Nj + 3 gradient components of |G_device - G_float128| / scale, and the relative error of -logL.  Output: one JSON line per case."""
import json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle as O
import sfh_b200 as S
from conftest import hier_grad_scale, make_hier_problem

MODELS = {O.POWERLAW_MZR: (S.PowerLawMZR(1.0, -2.0, 6.0), (6.0,), "PowerLawMZR"),
          O.LINEAR_AMR: (S.LinearAMR(0.05, -1.6, 12.0), (12.0,), "LinearAMR"),
          O.LOG_AMR: (S.LogarithmicAMR(1e-4, 5e-5, 12.0), (12.0, 0.01524, 0.2485, 1.78), "LogarithmicAMR")}
worst = 0.0
for kind, (model, fixed, name) in MODELS.items():
    for shuffle, ragged, nj, nk, nb in [(False, False, 21, 26, 1500), (True, False, 21, 26, 1500), (True, True, 21, 26, 1500), (False, False, 60, 40, 3000)]:
        p = make_hier_problem(nj=nj, nk=nk, nb=nb, shuffle=shuffle, ragged=ragged, **({"la_hi": 10.1, "la_lo": 6.6} if nj == 60 and kind == O.POWERLAW_MZR else {}))
        x = O.calculate_coeffs(kind, model.alpha, model.beta, fixed, 0.2, p["R"], p["logAge"], p["MH"])
        data = p["rng"].poisson(p["M"] @ x).astype(np.float64)
        disp = S.GaussianDispersion(0.2)
        ds = S.DeviceStack(p["M"], data)
        for label, v in (("at truth", np.concatenate([p["R"], [model.alpha, model.beta, 0.2]])),
                         ("perturbed", np.concatenate([p["R"] * 1.5, [model.alpha * 1.2, model.beta, 0.2 * 1.3]]))):
            G = np.empty(v.shape[0])
            nl = S.fg_(True, G, model, disp, v, ds, data, None, p["logAge"], p["MH"])
            nlq, Gq, _ = O.fg_hier(kind, fixed, (1, 1, 1), v, p["M"], data, p["logAge"], p["MH"], quad=True)
            sc = hier_grad_scale(kind, fixed, v, p["M"], data, p["logAge"], p["MH"])
            e = float(np.max(np.abs(G - Gq) / sc)); worst = max(worst, e)
            print(json.dumps({"model": name, "grid": f"{nj}x{nk}" + (" shuffled" if shuffle else "") + (" ragged" if ragged else ""), "point": label,
                              "logL_rel_err": float(abs(nl - nlq) / abs(nlq)), "max_grad_err_over_scale": e,
                              "max_grad_err_over_abs_value": float(np.max(np.abs(G - Gq) / np.abs(Gq)))}), flush=True)
print(json.dumps({"worst_max_grad_err_over_scale": worst, "bar": 1e-10}))
