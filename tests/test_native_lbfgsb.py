"""The library's native L-BFGS-B loop (csrc/sfh_drivers.h, include/sfhcuda.h: sfh_minimize_lbfgsb) on the CPU through its
objective callback.  fit_templates_lbfgsb hands fg! to LBFGSB.jl (the Fortran L-BFGS-B) with lb = 0, m = 10, factr = 1e-12,
pgtol = 1e-5 (solvers.jl:82-90); scipy wraps the same Fortran lineage, so it is the checker: same converged answers, same set of
active bounds, the same order of work.  (The engine is third-party in the reference: iterates are not a parity claim.)
"""
import ctypes as C

import numpy as np
import pytest
from scipy import optimize

import oracle as O
import sfh_b200
from conftest import make_flat_problem
from sfh_b200.solvers import native_lbfgsb

L = sfh_b200._lib


def rosen(x):
    return optimize.rosen(x), optimize.rosen_der(x)


def test_unconstrained_and_bounded_rosenbrock():
    x0 = np.array([-1.2, 1.0, -1.2, 1.0, 0.5, 0.5])
    x, f, info = native_lbfgsb(rosen, x0, pgtol=1e-8)
    xs, fs, si = optimize.fmin_l_bfgs_b(rosen, x0, factr=1e-12, pgtol=1e-8)
    np.testing.assert_allclose(x, np.ones(6), atol=1e-7)
    assert info["status"] == 0 and info["pg_norm"] <= 1e-8 and info["funcalls"] < 2 * si["funcalls"]
    lb, ub = np.full(6, -2.0), np.full(6, 0.8)                      # the upper bound is active at the solution
    x, f, info = native_lbfgsb(rosen, x0, lb, ub, pgtol=1e-8)
    xs, fs, si = optimize.fmin_l_bfgs_b(rosen, x0, bounds=list(zip(lb, ub)), factr=1e-12, pgtol=1e-8)
    np.testing.assert_allclose(x, xs, atol=1e-6)
    assert x[0] == 0.8 and abs(f - fs) <= 1e-10 * abs(fs) and np.all(x >= lb) and np.all(x <= ub)
    assert info["funcalls"] < 2 * si["funcalls"]
    # a start outside the box is projected onto it first
    x, f, _ = native_lbfgsb(rosen, np.full(6, 5.0), lb, ub, pgtol=1e-8)
    np.testing.assert_allclose(x, xs, atol=1e-6)


def test_bound_constrained_quadratic_kkt():
    rng = np.random.default_rng(2)
    n = 40
    A = rng.standard_normal((n, n)); A = A @ A.T / n + 0.1 * np.eye(n)
    b = rng.standard_normal(n) * 3
    fun = lambda x: (0.5 * x @ A @ x - b @ x, A @ x - b)
    lb, ub = np.full(n, -0.5), np.full(n, 0.7)
    x, f, info = native_lbfgsb(fun, np.zeros(n), lb, ub, pgtol=1e-9)
    g = A @ x - b
    lo, hi, free = x == lb, x == ub, (x > lb) & (x < ub)
    assert lo.sum() > 0 and hi.sum() > 0 and free.sum() > 0         # a mix of active and free variables
    assert np.all(g[lo] >= -1e-8) and np.all(g[hi] <= 1e-8) and np.abs(g[free]).max() <= 1e-8   # KKT conditions
    xs, fs, _ = optimize.fmin_l_bfgs_b(fun, np.zeros(n), bounds=list(zip(lb, ub)), factr=1e-12, pgtol=1e-9)
    np.testing.assert_allclose(x, xs, atol=1e-7)


@pytest.mark.parametrize("nb,nt,zero_every", [(2000, 30, 3), (10000, 100, 0)])
def test_poisson_fit_with_active_lower_bounds(nb, nt, zero_every):
    """fit_templates_lbfgsb's problem (solvers.jl:82-90) on the oracle's fg!: the second case is BASELINE config 1's shape."""
    M, xt, data = make_flat_problem(nb, nt)
    if zero_every:
        xt = xt.copy(); xt[::zero_every] = 0
        data = np.random.default_rng(1).poisson(M @ xt).astype(np.float64)

    def fg(z):
        f, G, _ = O.fg(z, M, data)
        return float(f), G
    x0 = np.ones(nt) * data.sum() / (M @ np.ones(nt)).sum()         # renormalize_x0
    x, f, info = native_lbfgsb(fg, x0, lb=0.0)
    xs, fs, si = optimize.fmin_l_bfgs_b(fg, x0, bounds=[(0, None)] * nt, factr=1e-12, pgtol=1e-5, m=10, maxfun=100000, maxiter=100000)
    assert info["status"] == 0 and info["pg_norm"] <= 1e-5
    assert np.linalg.norm(x - xs) <= 1e-5 * np.linalg.norm(xs) and abs(f - fs) <= 1e-10 * abs(fs)
    assert np.array_equal(x == 0, xs == 0) and np.all(x >= 0)      # the same coefficients sit on the bound
    if zero_every:
        assert (x == 0).sum() > 0
    assert info["funcalls"] <= 1.5 * si["funcalls"] + 5


def test_failure_modes_and_raw_abi():
    def boom(x):
        raise ArithmeticError("objective failed")
    with pytest.raises(ArithmeticError):
        native_lbfgsb(boom, np.zeros(3))
    x, f, info = native_lbfgsb(rosen, np.full(4, -1.0), maxiter=2)
    assert info["status"] == 2 and info["nit"] == 2
    x, f, info = native_lbfgsb(lambda x: (np.inf, np.zeros_like(x)), np.zeros(2))
    assert info["status"] == 4
    x, f, info = native_lbfgsb(lambda x: (float(x @ x), 2 * x), np.zeros(3))          # already optimal
    assert info["status"] == 0 and info["nit"] == 0 and info["funcalls"] == 1
    with pytest.raises(ValueError):
        native_lbfgsb(rosen, np.zeros(2), lb=1.0, ub=0.0)
    dp = C.POINTER(C.c_double)
    xx = np.zeros(2)
    o = L.sfh_lbfgsb_opts(); o.struct_size = 1
    cb = L.sfh_objective_fn(lambda *a: 0)
    assert L.lib.sfh_minimize_lbfgsb(cb, None, 2, xx.ctypes.data_as(dp), None, None, C.byref(o), None) == L.SFH_ERR_INVALID_ARG
    assert L.lib.sfh_minimize_lbfgsb(L.sfh_objective_fn(), None, 2, xx.ctypes.data_as(dp), None, None, None, None) == L.SFH_ERR_INVALID_ARG
    assert L.lib.sfh_fit_templates_lbfgsb(None, None, None, None) == L.SFH_ERR_INVALID_ARG
    assert C.sizeof(L.sfh_lbfgsb_opts) == 40 and C.sizeof(L.sfh_lbfgsb_report) == 40
