"""Regenerates the committed golden fixtures under tests/golden/.

  fitting_core_kats.json   hand-transcribed known-answer tests of the REFERENCE's own test-suite
                           (/root/reference/test/fitting/fitting_core_test.jl and the model doctests), each with the
                           file:line it comes from.  These are literals of the reference, not outputs of this repo.
  seeded_vectors.npz       __float128-arbiter outputs of the CPU oracle for seeded inputs (numpy Philox; the inputs are
                           regenerated from their seeds by tests/conftest.py, only an input checksum is stored):
                           frozen regression vectors so that the GPU parity tests also have a comparison that does not
                           depend on the oracle library being rebuilt identically on the GPU box.

Run from the repository root:  python tests/golden/make_golden.py
(The reference is Julia and cannot be executed in this image, so no fixture here is produced by running it.)
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

KATS = {
    "_source": "cgarling/StarFormationHistories.jl v1.3.1, test/fitting/fitting_core_test.jl + docstring doctests",
    "rtol": {"float32": 1e-3, "float64": 1e-7, "_cite": "fitting_core_test.jl:4-6"},
    "composite": {"_cite": "fitting_core_test.jl:14-28",
                  "A": [[0, 0, 0], [1, 1, 1], [0, 0, 0]], "B": [[0, 0, 0], [0, 0, 0], [1, 1, 1]], "coeffs": [1, 2],
                  "C": [[0, 0, 0], [1, 1, 1], [2, 2, 2]], "A2": [0, 1, 0, 0, 1, 0, 0, 1, 0], "B2": [0, 0, 1, 0, 0, 1, 0, 0, 1],
                  "C2": [0, 1, 2, 0, 1, 2, 0, 1, 2]},
    "loglikelihood": {"_cite": "fitting_core_test.jl:37-67",
                      "C": [[1, 1, 1], [2, 2, 2], [3, 3, 3]], "data": [[1, 1, 1], [2, 2, 2], [2, 2, 2]], "value": -0.5672093513510137,
                      "A": [[1, 1, 1], [0, 0, 0], [0, 0, 0]], "B": [[0, 0, 0], [1, 1, 1], [1.5, 1.5, 1.5]], "coeffs": [1, 2],
                      "C_zero": [[1.5, 1.5, 1.5], [3, 3, 3], [3, 3, 3]], "data_zero": [[0, 0, 0], [2, 2, 2], [2, 2, 2]],
                      "value_zero": -5.6344187027020260},
    "grad": {"_cite": "fitting_core_test.jl:77-159",
             "model": [[0, 0, 0], [0, 0, 0], [1, 1, 1]], "single": -1, "model_zero": [[1, 1, 1], [0, 0, 0], [0, 0, 0]], "single_zero": -3,
             "models": [[[1, 1, 1], [0, 0, 0], [0, 0, 0]], [[0, 0, 0], [1, 1, 1], [0, 0, 0]], [[0, 0, 0], [0, 0, 0], [1, 1, 1]]],
             "coeffs": [1.5, 3, 3], "G": [-1, -1, -1], "G_zero_data": [-3, -1, -1]},
    "fg": {"_cite": "fitting_core_test.jl:168-192", "neg_logL": 1.4180233783775342, "G": [1, 1, 1]},
    "doctests": {"gaussian_dispersion": {"_cite": "dispersion_models.jl:63-68", "sigma": 0.2, "x": 1.0, "mu": 1.2,
                                         "value_is_exp_minus_half": True, "grad": [3.0326532985631656, -3.0326532985631656]},
                 "powerlaw_mzr": {"_cite": "mzr.jl:245-250", "alpha": 1.0, "MH0": -1, "logMstar0": 6, "at_1e7": 0.0,
                                  "grad_at_1e8": [2.0, 1.0, "1/1e8/ln(10)"]}},
    "metallicity_utilities": {"_cite": "test/utilities/utilities_test.jl:32-47 (rtol 1e-3 Float32 / 1e-7 Float64; the literals were generated "
                                       "from the Float32 value of 1e-3 and are reproduced to every digit with that input)",
                              "Z": 1e-3, "Z_float32_literal": 0.0010000000474974513, "solZ": 0.01524, "Y_p": 0.2485, "gamma": 1.78,
                              "Y_from_Z": 0.2502800000845455, "X_from_Z": 0.748719999867957, "X_from_Z_Yp_0.25": 0.74722,
                              "MH_from_Z": -1.206576807011171, "Z_from_MH_at_-2": 0.00016140871730361718, "dMH_dZ": 435.9070188458886,
                              "Martin2016_complete(20,1,25,1)": 0.9933071490757151444406380196186748,
                              "exp_photerr(20,1.05,10,32,0.01)": 0.01286605230281143891186877135084309},
    "unreproducible_here": {"_why": "inputs come from StableRNG + Distributions.Poisson streams (Julia only)",
                            "mzr_test.jl:74-76": {"nlogL": 4917.491550052553}, "amr_test.jl:45-47": {"nlogL": 4903.0966770848445},
                            "amr_test.jl:264-265": {"nlogL": 5006.412301383171}},
}


def main():
    json.dump(KATS, open(os.path.join(HERE, "fitting_core_kats.json"), "w"), indent=1)
    import oracle as O
    from conftest import make_flat_problem, make_hier_problem
    out = {}
    for tag, (nb, nt, seed) in {"a": (257, 19, 101), "b": (1000, 64, 102), "c": (3001, 130, 103)}.items():
        M, x, data = make_flat_problem(nb, nt, seed=seed)
        nl, G, gs, comp = O.fg_quad(x * 1.2, M, data)
        out.update({f"flat_{tag}_shape": np.array([nb, nt, seed]), f"flat_{tag}_insum": np.array([M.sum(), x.sum(), data.sum()]),
                    f"flat_{tag}_nl": nl, f"flat_{tag}_G": G, f"flat_{tag}_gscale": gs, f"flat_{tag}_composite": comp})
    for kind, (a, b, fixed) in {0: (1.0, -2.0, (6.0,)), 1: (0.05, -1.6, (12.0,)), 2: (1e-4, 5e-5, (12.0, 0.01524, 0.2485, 1.78))}.items():
        p = make_hier_problem(nj=9, nk=7, nb=300, seed=200 + kind, shuffle=True)
        xt = O.calculate_coeffs(kind, a, b, fixed, 0.2, p["R"], p["logAge"], p["MH"])
        data = np.random.Generator(np.random.Philox(300 + kind)).poisson(p["M"] @ xt).astype(np.float64)
        v = np.concatenate([p["R"], [a, b, 0.2]]) * 1.1
        nl, G, _ = O.fg_hier(kind, fixed, (1, 1, 1), v, p["M"], data, p["logAge"], p["MH"], quad=True)
        out.update({f"hier_{kind}_insum": np.array([p["M"].sum(), p["logAge"].sum(), data.sum()]),
                    f"hier_{kind}_v": v, f"hier_{kind}_fixed": np.array(fixed), f"hier_{kind}_nl": nl, f"hier_{kind}_G": G,
                    f"hier_{kind}_coeffs": O.calculate_coeffs(kind, v[-3], v[-2], fixed, v[-1], v[:-3], p["logAge"], p["MH"])})
    np.savez_compressed(os.path.join(HERE, "seeded_vectors.npz"), **out)
    print("wrote", os.listdir(HERE))


if __name__ == "__main__":
    main()
