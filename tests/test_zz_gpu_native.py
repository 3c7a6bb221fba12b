"""GPU tests of the native driver loops (include/sfhcuda.h: sfh_fit_templates_bfgs, sfh_fit_fixed_amr_bfgs, sfh_fit_sfh_bfgs):
the same assertions the reference makes of its Optim-driven solvers (basic_linear_combinations.jl:16-118, mzr_test.jl:178-216,
fixed_amr_test.jl:56-110), now with the whole BFGS optimisation running inside the library around the device evaluations --
plus agreement with the scipy-driven host loop on the same device objective.  (The engine itself is pinned on the CPU by
tests/test_native_bfgs.py.)"""
import numpy as np
import pytest

from conftest import make_flat_problem, make_hier_problem

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def S():
    import sfh_b200
    return sfh_b200


def isapprox(a, b, rtol):
    a, b = np.asarray(a, float), np.asarray(b, float)
    return np.linalg.norm(a - b) <= rtol * max(np.linalg.norm(a), np.linalg.norm(b))


def test_fit_templates_native_recovers_truth(S):                   # basic_linear_combinations.jl:16-89
    models = [np.array([[0, 0, 0], [0, 0, 0], [1, 1, 1]], dtype=np.float64)]
    data = np.array([[0, 0, 0], [0, 0, 0], [3, 3, 3]], dtype=np.int64)
    assert S.fit_templates(models, data, x0=np.array([1.0]), engine="native")["mle"].mu[0] == pytest.approx(3, rel=1e-7)
    assert S.fit_templates_fast(models, data, x0=np.array([1.0]), engine="native")[0][0] == pytest.approx(3, rel=1e-7)
    rng = np.random.Generator(np.random.Philox(58392))
    N = 10
    x, x0 = rng.random(N), rng.random(N)
    models = [rng.random((100, 100)) for _ in range(N)]
    sm = S.stack_models(models)
    sd = sum(c * m for c, m in zip(x, models)).reshape(-1, order="F")
    r = S.fit_templates(sm, sd, x0=x0, engine="native")
    assert isapprox(r["mle"].mu, x, 1e-7) and r["mle"].result.nit > 0 and r["map"].invH.shape == (N, N)
    assert isapprox(S.fit_templates_fast(sm, sd, x0=x0, engine="native")[0], x, 1e-7)


def test_fit_templates_native_agrees_with_host_loop_on_poisson_data(S):   # basic_linear_combinations.jl:92-118
    M, x, data = make_flat_problem(10000, 20)
    x0 = np.ones(20)
    rn = S.fit_templates(M, data, x0=x0, engine="native")
    rs = S.fit_templates(M, data, x0=x0, engine="scipy")
    for k in ("map", "mle"):
        assert isapprox(rn[k].mu, rs[k].mu, 1e-5)
        assert isapprox(rn[k].mu, x, 5e-2)                     # Poisson noise: ~2 % on these coefficients
        assert np.all(rn[k].sigma > 0) and np.allclose(rn[k].sigma, rs[k].sigma, rtol=0.5)
    assert rn["map"].result.success
    assert isapprox(S.fit_templates_fast(M, data, x0=x0, engine="native")[0], rs["mle"].mu, 1e-4)


def test_fit_templates_lbfgsb_native(S):                          # basic_linear_combinations.jl:16-118
    models = [np.array([[0, 0, 0], [0, 0, 0], [1, 1, 1]], dtype=np.float64)]
    data = np.array([[0, 0, 0], [0, 0, 0], [3, 3, 3]], dtype=np.int64)
    assert S.fit_templates_lbfgsb(models, data, x0=np.array([1.0]), engine="native")[1][0] == pytest.approx(3, rel=1e-7)
    rng = np.random.Generator(np.random.Philox(58392))
    N = 10
    x, x0 = rng.random(N), rng.random(N)
    models = [rng.random((100, 100)) for _ in range(N)]
    sm = S.stack_models(models)
    sd = sum(c * m for c, m in zip(x, models)).reshape(-1, order="F")
    assert isapprox(S.fit_templates_lbfgsb(sm, sd, x0=x0, engine="native")[1], x, 1e-6)    # (stops at pgtol = 1e-5 like the reference)
    x2 = x.copy(); x2[0] = 0; x2[-1] = 0                               # :66-89 coefficients on the bound
    d2 = sum(c * m for c, m in zip(x2, models)).reshape(-1, order="F")
    f2, r2 = S.fit_templates_lbfgsb(sm, d2, x0=x0, engine="native")
    assert isapprox(r2, x2, 1e-6) and np.all(r2 >= 0)
    M, xt, data = make_flat_problem(10000, 100)                        # BASELINE config 1
    fn, xn = S.fit_templates_lbfgsb(M, data, x0=np.ones(100), engine="native")
    fs, xs = S.fit_templates_lbfgsb(M, data, x0=np.ones(100), engine="scipy")
    assert isapprox(xn, xs, 1e-5) and abs(fn - fs) <= 1e-10 * abs(fs) and isapprox(xn, xt, 5e-2)


def test_fit_sfh_native(S):                                         # mzr_test.jl:178-216
    p = make_hier_problem(nj=21, nk=26, nb=10000)
    mz, dp = S.PowerLawMZR(1.0, -2.0, 6.0), S.GaussianDispersion(0.2)
    xt = S.calculate_coeffs(mz, dp, p["R"], p["logAge"], p["MH"])
    truth = np.concatenate([p["R"], [1.0, -2.0, 0.2]])
    data2 = p["M"] @ xt
    x0 = p["R"] * 1.5
    start_mz, start_dp = S.PowerLawMZR(1.2, -2.2, 6.0), S.GaussianDispersion(0.25)
    res = S.fit_sfh(start_mz, start_dp, p["M"], data2, p["logAge"], p["MH"], x0=x0, engine="native")
    assert np.allclose(res["mle"].mu, truth, rtol=1e-4), np.max(np.abs(res["mle"].mu / truth - 1))
    data = p["rng"].poisson(p["M"] @ xt).astype(np.float64)
    rn = S.fit_sfh(start_mz, start_dp, p["M"], data, p["logAge"], p["MH"], x0=x0, engine="native")
    rs = S.fit_sfh(start_mz, start_dp, p["M"], data, p["logAge"], p["MH"], x0=x0, engine="scipy")
    z = np.abs(rn["map"].mu - truth) / rn["map"].sigma
    assert np.mean(z < 3) > 0.9
    assert np.allclose(rn["map"].mu, rs["map"].mu, rtol=1e-4)
    # the native loop's objective at its minimiser equals the host adapter's objective there (same device fg!)
    opt = S.HierarchicalOptimizer(start_mz, start_dp, p["M"], data, p["logAge"], p["MH"], True, True, True)
    lp, g = opt.logdensity_and_gradient(rn["map"].result.x)
    assert abs(-lp - rn["map"].result.fun) <= 1e-12 * abs(lp) and np.abs(g).max() <= 1e-6
    # fixed sigma stays fixed, gets zero uncertainty, and is absent from the fitting space (:202-216)
    rf = S.fit_sfh(start_mz, S.GaussianDispersion(0.2, (False,)), p["M"], data, p["logAge"], p["MH"], x0=x0, engine="native")
    assert rf["mle"].mu[-1] == 0.2 and rf["mle"].sigma[-1] == 0.0 and rf["mle"].invH.shape == (23, 23)
    ra = S.fit_sfh(S.LinearAMR(0.12, -2.2, 13.7), start_dp, p["M"], data, p["logAge"], p["MH"], x0=x0, engine="native")
    assert np.all(np.isfinite(ra["map"].mu)) and ra["map"].result.nit > 0


def test_fixed_amr_native(S):                                       # fixed_amr_test.jl:56-75
    rng = np.random.Generator(np.random.Philox(58392))
    uA, uM = np.linspace(10.0, 8.0, 12), np.linspace(-2.5, 0.0, 15)
    la, mh = np.repeat(uA, 15), np.tile(uM, 12)
    mz, dp = S.PowerLawMZR(1.0, -2.0, 6.0), S.GaussianDispersion(0.2)
    SFRs = rng.random(12)
    x = S.calculate_coeffs(mz, dp, SFRs, la, mh)
    models = [rng.random((30, 25)) * 100 for _ in range(la.shape[0])]
    data = sum(c * m for c, m in zip(x, models))
    x0 = S.construct_x0_mdf(la, 13.7, normalize_value=1)
    relw = S.calculate_coeffs(mz, dp, np.ones(12), la, mh)
    rn = S.fixed_amr(models, data, la, mh, relw, x0=x0, engine="native")
    assert np.allclose(rn["mle"]["mu"], SFRs, rtol=1e-5) and rn["mle"]["invH"].shape == (12, 12)
    rs = S.fixed_amr(models, data, la, mh, relw, x0=x0, engine="scipy")
    assert np.allclose(rn["map"]["mu"], rs["map"]["mu"], rtol=1e-5)


def test_native_loop_reports_failures(S):
    M, x, data = make_flat_problem(400, 4)
    ds = S.DeviceStack(M, data)
    L = S._lib
    import ctypes as C
    th = np.zeros(4)
    rep = L.sfh_bfgs_report()
    dp = C.POINTER(C.c_double)
    assert L.lib.sfh_fit_templates_bfgs(ds.ctx().handle, 7, th.ctypes.data_as(dp), None, C.byref(rep), None) == L.SFH_ERR_INVALID_ARG
    par = np.array([1.0, -2.0, 0.2]); tf = np.array([1, 0, 1], dtype=np.int32); fr = np.array([1, 1, 1], dtype=np.uint8)
    xv = np.zeros(5)
    st = L.lib.sfh_fit_sfh_bfgs(ds.new_ctx().handle, 0, np.array([6.0, 0, 0, 0]).ctypes.data_as(dp), 0, par.ctypes.data_as(dp),
                                tf.ctypes.data_as(C.POINTER(C.c_int32)), fr.ctypes.data_as(C.POINTER(C.c_uint8)), 1, xv.ctypes.data_as(dp),
                                None, C.byref(rep), None)
    assert st == L.SFH_ERR_NOT_BOUND
    o = L.sfh_bfgs_opts(); o.struct_size = C.sizeof(L.sfh_bfgs_opts); o.maxiter = 2
    th = np.log(np.full(4, 1.0))
    assert L.lib.sfh_fit_templates_bfgs(ds.ctx().handle, L.SFH_FIT_LOG_MLE, th.ctypes.data_as(dp), C.byref(o), C.byref(rep), None) == L.SFH_OK
    assert rep.status == 1 and rep.iterations == 2 and not rep.converged


# ---- the BFGS inverse Hessian resident on the device (sfh_bfgs_opts.device_hessian): three kernels in csrc/sfh_small.cuh ----
@pytest.mark.parametrize("nb,nt", [(10000, 20), (20000, 600)])
def test_device_hessian_bfgs_matches_host_hessian(S, nb, nt):
    M, x, data = make_flat_problem(nb, nt)
    a = S.fit_templates(M, data, x0=np.ones(nt), engine="native")
    b = S.fit_templates(M, data, x0=np.ones(nt), engine="native", device_hessian=True)
    for k in ("map", "mle"):
        assert b[k].result.success == a[k].result.success
        assert np.linalg.norm(a[k].mu - b[k].mu) <= 1e-6 * np.linalg.norm(a[k].mu)
        H = b[k].invH
        assert np.allclose(H, H.T, atol=1e-10 * np.abs(H).max()) and np.all(np.diag(H) > 0)
        assert np.allclose(np.sqrt(np.diag(H)), np.sqrt(np.diag(a[k].invH)), rtol=0.3)
