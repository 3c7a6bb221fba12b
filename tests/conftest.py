import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu under gpurun)")


def _have_gpu() -> bool:
    try:                                   # the library's own probe: no torch import (a minute on a fresh box)
        import sfh_b200
        return sfh_b200.device_count() > 0
    except Exception:
        pass
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _have_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


# ---------------------------------------------------------------------------------------------
# Synthetic inputs shared by oracle / GPU tests.  They mirror the reference's own generators
# (test/fitting/basic_linear_combinations.jl:93, test/fitting/mzr_test.jl:52-66) with numpy's
# Philox counter-based generator seeded by the reference's seed VALUES (its StableRNG streams are
# not reproducible without Julia -- SURVEY.md section 8d).
# ---------------------------------------------------------------------------------------------
def make_flat_problem(nb, nt, seed=58392, dtype=np.float64, scale=100.0):
    rng = np.random.Generator(np.random.Philox(seed))
    M = np.asfortranarray(rng.random((nb, nt)).astype(dtype))
    x = scale * rng.random(nt)
    lam = M.astype(np.float64) @ x
    data = rng.poisson(lam).astype(dtype)
    return M, x, data


def make_hier_problem(nj=21, nk=26, nb=400, seed=94823, la_hi=10.0, la_lo=8.0, shuffle=False, ragged=False):
    """mzr_test.jl:52-66 shaped problem: ages x metallicities grid, templates U(0,1)/1e5, R = 1e6 U(0,1)."""
    rng = np.random.Generator(np.random.Philox(seed))
    uA = np.linspace(la_hi, la_lo, nj)
    uM = np.linspace(-2.5, 0.0, nk)
    if shuffle:
        uA = uA[rng.permutation(nj)]
    logAge = np.repeat(uA, nk)
    MH = np.tile(uM, nj)
    if ragged:  # drop a few templates so Nk differs between ages
        keep = np.ones(nj * nk, dtype=bool)
        drop = rng.choice(nj * nk, size=max(1, nj * nk // 15), replace=False)
        keep[drop] = False
        for j in range(nj):  # never drop a whole age
            if not keep[j * nk:(j + 1) * nk].any():
                keep[j * nk] = True
        logAge, MH = logAge[keep], MH[keep]
    nt = logAge.shape[0]
    R = rng.random(nj) * 1e6
    M = np.asfortranarray(rng.random((nb, nt)) / 1e5)
    return dict(nj=nj, nt=nt, nb=nb, logAge=logAge, MH=MH, R=R, M=M, rng=rng)


def hier_grad_scale(kind, fixed, v, M, data, logAge, MH, eps=np.finfo(np.float64).eps):
    """Backward-error scale of the hierarchical gradient, the analogue of test_gpu_core.assert_grad_close's `gscale`.

    G_p = sum_t fullG_t * J_tp with fullG_t = -(M' r)_t and J = d r / d variables.  Any correct implementation carries a
    rounding error of eps-level times  sum_t |J_tp| * sum_i |M_it r_i|  per component (re-ordering the bin sums and the
    member sums; the parameter components are near-total cancellations at the optimum), so that is what 1e-10 is taken
    relative to -- not |G_p|.  J comes from central differences of the oracle's calculate_coeffs: it only sets a SCALE, so
    ~1e-6 relative accuracy is plenty."""
    import oracle as O
    v = np.asarray(v, dtype=np.float64)
    nj = v.shape[0] - 3

    def coeffs(u):
        return O.calculate_coeffs(kind, u[nj], u[nj + 1], fixed, u[nj + 2], u[:nj], logAge, MH)

    c = coeffs(v)
    m = np.asarray(M, dtype=np.float64) @ c
    r = 1.0 - np.asarray(data, dtype=np.float64) / np.maximum(m, eps)
    gs = np.abs(M).T @ np.abs(r)                       # flat backward-error scale per template
    scale = np.empty(nj + 3)
    for p in range(nj + 3):
        h = 1e-6 * max(abs(v[p]), 1e-12)
        up, dn = v.copy(), v.copy()
        up[p] += h; dn[p] -= h
        J = (coeffs(up) - coeffs(dn)) / (2 * h)
        scale[p] = np.abs(J) @ gs
    return scale


def assert_hier_grad_close(G, Gq, scale, rtol=1e-10, what=""):
    err = np.abs(np.asarray(G) - np.asarray(Gq)) / (scale + 1e-300)
    assert np.all(err <= rtol), f"{what}: max |dG| / scale = {float(err.max()):.3e} at component {int(err.argmax())} (bar {rtol:g})"
