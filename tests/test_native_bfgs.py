"""The library's native BFGS loop (csrc/sfh_drivers.h, include/sfhcuda.h: sfh_minimize_bfgs) on the CPU, through its generic
objective callback.  The engine is third-party in the reference (Optim.jl BFGS + HagerZhang), so what is pinned is what the
reference's own tests pin for it: CONVERGED answers (basic_linear_combinations.jl:16-118) -- here against scipy's BFGS and
closed forms -- plus the quality of the returned inverse Hessian, which the drivers turn into parameter uncertainties
(solvers.jl:209-219, generic_fitting.jl:352-407).  The device-bound objectives (sfh_fit_*_bfgs) need a GPU
(tests/test_zz_gpu_native.py); the Poisson objective used here is the CPU oracle's fg!.
"""
import ctypes as C

import numpy as np
import pytest
from scipy import optimize

import oracle as O
import sfh_b200
from conftest import make_flat_problem
from sfh_b200.solvers import native_bfgs

L = sfh_b200._lib


def rosen(x):
    return optimize.rosen(x), optimize.rosen_der(x)


@pytest.mark.parametrize("alphaguess", [1, 2])
@pytest.mark.parametrize("n", [2, 10])
def test_rosenbrock(n, alphaguess):
    x0 = np.array([-1.2, 1.0]) if n == 2 else np.linspace(0.3, 1.6, n)   # (from -1.2, 1, ... a 10-d run may end in the local minimum)
    r = native_bfgs(rosen, x0, gtol=1e-9, alphaguess=alphaguess)
    assert r.success and r.status == 0 and r.g_norm <= 1e-9
    np.testing.assert_allclose(r.x, np.ones(n), rtol=0, atol=1e-7)
    assert r.fun < 1e-15 and r.nfev >= r.nit + 1
    ref = optimize.minimize(rosen, x0, jac=True, method="BFGS", options={"gtol": 1e-9})
    np.testing.assert_allclose(r.x, ref.x, atol=1e-6)
    assert r.nfev < 5 * ref.nfev + 50            # same order of work as scipy's engine


def test_quadratic_and_inverse_hessian():
    rng = np.random.default_rng(11)
    n = 30
    Q, _ = np.linalg.qr(rng.standard_normal((n, n)))
    A = Q @ np.diag(np.logspace(0, 4, n)) @ Q.T            # condition number 1e4
    A = (A + A.T) / 2
    b = rng.standard_normal(n)
    fun = lambda x: (0.5 * x @ A @ x - b @ x, A @ x - b)
    r = native_bfgs(fun, np.zeros(n), gtol=1e-10)
    assert r.success
    np.testing.assert_allclose(r.x, np.linalg.solve(A, b), rtol=1e-8, atol=1e-10)
    H = r.hess_inv
    np.testing.assert_allclose(H, H.T, rtol=0, atol=1e-12 * np.abs(H).max())          # symmetric
    assert np.linalg.eigvalsh((H + H.T) / 2).min() > 0                                # positive definite
    # on a quadratic BFGS reproduces the inverse Hessian on the subspace it has explored: after >= n steps, all of it
    assert r.nit >= n
    Ai = np.linalg.inv(A)
    assert np.abs(H - Ai).max() <= 2e-2 * np.abs(Ai).max()                            # (scipy's BFGS: 8e-3 on this problem)
    np.testing.assert_allclose(np.sqrt(np.diag(H)), np.sqrt(np.diag(Ai)), rtol=2e-2)  # what the drivers report as sigma


def test_poisson_objective_matches_scipy_and_exact_covariance():
    """fit_templates' MAP then MLE objectives (solvers.jl:178-195) on the oracle's fg!, native loop vs scipy's BFGS."""
    M, xtrue, data = make_flat_problem(900, 8, scale=100.0)

    def fg_map(th):
        x = np.exp(th)
        nl, G, _ = O.fg(x, M, data)
        return nl - th.sum(), G * x - 1

    def fg_mle(th):
        x = np.exp(th)
        nl, G, _ = O.fg(x, M, data)
        return nl, G * x
    x0 = np.log(np.full(8, data.sum() / M.sum()))
    for fun in (fg_map, fg_mle):
        r = native_bfgs(fun, x0, gtol=1e-8)
        ref = optimize.minimize(fun, x0, jac=True, method="BFGS", options={"gtol": 1e-8})
        assert r.success, r
        np.testing.assert_allclose(np.exp(r.x), np.exp(ref.x), rtol=1e-6)
        assert abs(r.fun - ref.fun) <= 1e-9 * abs(ref.fun)
        # exact Hessian of the objective in log space: diag(x) M' diag(n / m^2) M diag(x) + diag(G x)
        x = np.exp(r.x)
        m = M @ x
        Hex = (M * x).T @ ((data / m ** 2)[:, None] * (M * x)) + np.diag(fun(r.x)[1] + (1 if fun is fg_map else 0))
        sig_exact = np.sqrt(np.diag(np.linalg.inv(Hex)))
        sig = np.sqrt(np.diag(r.hess_inv))
        sig_scipy = np.sqrt(np.diag(ref.hess_inv))
        # the BFGS estimate is an approximation in either engine; ours must be at least as usable as scipy's
        err, err_scipy = np.abs(sig / sig_exact - 1).max(), np.abs(sig_scipy / sig_exact - 1).max()
        assert err < max(0.5, 2 * err_scipy), (err, err_scipy)
        x0 = r.x


@pytest.mark.parametrize("free3", [(True, True, True), (True, False, True), (False, True, False)])
@pytest.mark.parametrize("kind", [O.POWERLAW_MZR, O.LINEAR_AMR])
def test_fit_sfh_objective_and_loop_against_the_oracle_adapter(kind, free3):
    """sfh_fit_sfh_bfgs_generic (the transformed objective of generic_fitting.jl:90-199 + the loop) around the ORACLE's
    hierarchical fg!, against scipy's BFGS on the oracle's own restatement of logdensity_and_gradient."""
    from conftest import make_hier_problem
    from sfh_b200.solvers import native_fit_sfh_generic
    P = make_hier_problem(nj=6, nk=7, nb=300)
    nj = P["nj"]
    fixed = [6.0] if kind == O.POWERLAW_MZR else [13.7]
    true_par = np.array([1.0, -2.0, 0.2]) if kind == O.POWERLAW_MZR else np.array([0.1, -1.8, 0.2])
    coeffs = O.calculate_coeffs(kind, true_par[0], true_par[1], fixed, true_par[2], P["R"], P["logAge"], P["MH"])
    data = P["rng"].poisson(P["M"] @ coeffs).astype(np.float64)
    free = np.array(free3, dtype=bool)
    tf = np.array(O.TRANSFORMS[kind])
    inner = lambda v: O.fg_hier(kind, fixed, free3, v, P["M"], data, P["logAge"], P["MH"])[:2]
    par_start = true_par * np.where(free, 1.2, 1.0)
    xstart = np.concatenate([np.log(P["R"] * 1.3), np.array([np.log(p) if t == 1 else p for p, t in zip(par_start, tf)])[free]])
    for jac in (True, False):
        def neg(xv):
            lp, g = O.hier_logdensity_and_gradient(kind, fixed, free3, par_start, xv, P["M"], data, P["logAge"], P["MH"], jac)
            return -lp, -g
        # (1) the objective itself: with an unreachable tolerance the loop returns after ONE evaluation, at the start point
        r0 = native_fit_sfh_generic(inner, nj, par_start, tf, free, xstart, jac, gtol=1e300)
        f0, g0 = neg(xstart)
        assert r0.nfev == 1 and r0.nit == 0
        assert abs(r0.fun - f0) <= 1e-13 * abs(f0) and abs(r0.g_norm - np.abs(g0).max()) <= 1e-12 * np.abs(g0).max()
        # (2) the converged fit
        r = native_fit_sfh_generic(inner, nj, par_start, tf, free, xstart, jac, gtol=1e-6)
        assert r.success, r
        fr, gr = neg(r.x)                                      # stationarity checked independently with the oracle's adapter
        assert np.abs(gr).max() <= 1e-6 and abs(fr - r.fun) <= 1e-12 * abs(fr)
        with np.errstate(all="ignore"):
            ref = optimize.minimize(neg, xstart, jac=True, method="BFGS", options={"gtol": 1e-6})
        if not np.isfinite(ref.fun) or np.abs(ref.jac).max() > 1e-4:
            continue                                           # scipy's engine diverged from this start (overflow in exp); ours did not
        assert r.fun <= ref.fun + 1e-9 * abs(ref.fun)
        # (an age whose best-fit mass is zero runs to log R -> -inf in both engines: compare on the scale of the masses)
        np.testing.assert_allclose(np.exp(r.x[:nj]), np.exp(ref.x[:nj]), rtol=2e-3, atol=1e-6 * P["R"].max())
        np.testing.assert_allclose(r.x[nj:], ref.x[nj:], rtol=2e-3, atol=2e-4)
    # refused: a free parameter with the reference's unvalidated negative-log transform
    with pytest.raises(sfh_b200.SFHError) as ei:
        native_fit_sfh_generic(inner, nj, par_start, np.array([-1, 0, 1]), np.array([True, True, True]), np.zeros(nj + 3))
    assert ei.value.status == L.SFH_ERR_UNSUPPORTED


def test_infinite_region_and_failure_modes():
    # a barrier: f = +inf outside (0, 2); the line search must back off into the domain
    def barrier(x):
        if np.any(x <= 0) or np.any(x >= 2):
            return np.inf, np.full_like(x, np.nan)
        return float(np.sum(-np.log(x) - np.log(2 - x) + 3 * x)), -1 / x + 1 / (2 - x) + 3
    r = native_bfgs(barrier, np.array([1.9, 0.05, 1.0]), gtol=1e-10)
    assert r.success
    np.testing.assert_allclose(r.x, np.full(3, (8 - np.sqrt(40)) / 6), rtol=1e-8)   # the root of 3x^2 - 8x + 2 inside (0, 2)
    # iteration limit
    x0 = np.full(10, -1.2)
    r = native_bfgs(rosen, x0, gtol=1e-12, maxiter=3)
    assert not r.success and r.status == 1 and r.nit == 3
    # start where the objective is not finite
    r = native_bfgs(barrier, np.array([3.0, 1.0, 1.0]))
    assert not r.success and r.status == 3 and r.nit == 0
    # an exception in the objective surfaces as that exception, not as a crash
    def boom(x):
        raise ZeroDivisionError("objective failed")
    with pytest.raises(ZeroDivisionError):
        native_bfgs(boom, np.zeros(2))
    # already converged at the start: zero iterations, identity inverse Hessian (Optim's initial invH)
    r = native_bfgs(lambda x: (float(x @ x), 2 * x), np.zeros(4))
    assert r.success and r.nit == 0 and r.nfev == 1 and np.array_equal(r.hess_inv, np.eye(4))


def test_large_problem_takes_the_threaded_update():
    # n = 600: the n x n update is split over host threads (for_columns); answer must not depend on that
    rng = np.random.default_rng(3)
    n = 600
    d = np.linspace(1.0, 50.0, n)
    c = rng.standard_normal(n)
    fun = lambda x: (float(0.5 * d @ (x - c) ** 2 + 0.25 * np.sum((x - c) ** 4)), d * (x - c) + (x - c) ** 3)
    r = native_bfgs(fun, np.zeros(n), gtol=1e-9)
    assert r.success
    np.testing.assert_allclose(r.x, c, atol=1e-8)
    np.testing.assert_allclose(r.hess_inv, r.hess_inv.T, atol=1e-12)


def test_raw_abi_argument_checks():
    rep = L.sfh_bfgs_report()
    x = np.zeros(2)
    dp = C.POINTER(C.c_double)
    cb = L.sfh_objective_fn(lambda u, xp, n, fp, gp: 0)
    assert L.lib.sfh_minimize_bfgs(L.sfh_objective_fn(), None, 2, x.ctypes.data_as(dp), None, C.byref(rep), None) == L.SFH_ERR_INVALID_ARG
    assert L.lib.sfh_minimize_bfgs(cb, None, 0, x.ctypes.data_as(dp), None, C.byref(rep), None) == L.SFH_ERR_INVALID_ARG
    o = L.sfh_bfgs_opts(); o.struct_size = 3
    assert L.lib.sfh_minimize_bfgs(cb, None, 2, x.ctypes.data_as(dp), C.byref(o), C.byref(rep), None) == L.SFH_ERR_INVALID_ARG
    assert b"struct_size" in L.lib.sfh_last_error()
    assert C.sizeof(L.sfh_bfgs_opts) == 32 and C.sizeof(L.sfh_bfgs_report) == 40
    o = L.sfh_bfgs_opts(); o.struct_size = C.sizeof(L.sfh_bfgs_opts); o.device_hessian = 1      # needs an entry point with a context
    assert L.lib.sfh_minimize_bfgs(cb, None, 2, x.ctypes.data_as(dp), C.byref(o), C.byref(rep), None) == L.SFH_ERR_UNSUPPORTED
    # device-bound variants validate their arguments before touching a device
    assert L.lib.sfh_fit_templates_bfgs(None, 0, x.ctypes.data_as(dp), None, None, None) == L.SFH_ERR_INVALID_ARG
    assert L.lib.sfh_fit_sfh_bfgs(None, 0, None, 0, None, None, None, 1, None, None, None, None) == L.SFH_ERR_INVALID_ARG
    assert L.lib.sfh_fit_fixed_amr_bfgs(None, None, None, 1, 0, None, None, None, None) == L.SFH_ERR_INVALID_ARG
