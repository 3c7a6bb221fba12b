"""Pins the CPU oracle against every golden value in the reference's own tests for the hot path
that is reproducible without Julia:

  test/fitting/fitting_core_test.jl:9-31    composite! / stack_models
  test/fitting/fitting_core_test.jl:32-70   loglikelihood  (-0.5672093513510137, -5.6344187027020260)
  test/fitting/fitting_core_test.jl:71-126  grad-loglikelihood (single / multi / coeffs forms)
  test/fitting/fitting_core_test.jl:127-162 grad-loglikelihood! incl. zero-data bins
  test/fitting/fitting_core_test.jl:163-195 fg!  (+1.4180233783775342, G = [1,1,1])
  doctests  dispersion_models.jl:63-68, mzr.jl:245-250
  test/fitting/mzr_test.jl:9-47             calculate_coeffs: sum_k r_jk = R_j, ordering

Float32 and Float64, tolerances exactly the reference's (rtol 1e-3 / 1e-7, fitting_core_test.jl:6).
"""
import numpy as np
import pytest

import oracle as O
from conftest import make_hier_problem

TYPES = [(np.float32, 1e-3), (np.float64, 1e-7)]  # fitting_core_test.jl:4-6


def jl(rows, dt):
    """Julia matrix literal `T[a b c; d e f]` -> numpy array (row-major literal, like Julia's)."""
    return np.array(rows, dtype=dt)


@pytest.mark.parametrize("T,rtol", TYPES)
def test_composite_and_stack_models(T, rtol):          # fitting_core_test.jl:9-31
    A = jl([[0, 0, 0], [1, 1, 1], [0, 0, 0]], T)
    B = jl([[0, 0, 0], [0, 0, 0], [1, 1, 1]], T)
    coeffs = np.array([1, 2], dtype=T)
    models2 = O.stack_models([A, B])
    A2 = np.array([0, 1, 0, 0, 1, 0, 0, 1, 0], dtype=T)
    B2 = np.array([0, 0, 1, 0, 0, 1, 0, 0, 1], dtype=T)
    assert np.array_equal(models2, np.stack([A2, B2], axis=1))            # :22-24
    C2 = O.composite(coeffs, models2, dtype=T)
    assert C2.dtype == T
    assert np.array_equal(C2, np.array([0, 1, 2, 0, 1, 2, 0, 1, 2], dtype=T))  # :27
    # matrix form: C == T[0 0 0; 1 1 1; 2 2 2]  (:18)
    assert np.array_equal(C2.reshape(3, 3, order="F"), jl([[0, 0, 0], [1, 1, 1], [2, 2, 2]], T))


@pytest.mark.parametrize("T,rtol", TYPES)
def test_loglikelihood(T, rtol):                        # fitting_core_test.jl:32-70
    Cm = jl([[1, 1, 1], [2, 2, 2], [3, 3, 3]], T)
    data = np.array([[1, 1, 1], [2, 2, 2], [2, 2, 2]], dtype=np.int64)
    r = O.loglikelihood(Cm, data, dtype=T)
    assert isinstance(r, T)                                                # :40 `isa T`
    assert r == pytest.approx(-0.5672093513510137, rel=rtol)              # :39
    A = jl([[1, 1, 1], [0, 0, 0], [0, 0, 0]], T)
    B = jl([[0, 0, 0], [1, 1, 1], [1.5, 1.5, 1.5]], T)
    comp = O.composite(np.array([1, 2], dtype=T), O.stack_models([A, B]), dtype=T)
    assert O.loglikelihood(comp, data, dtype=T) == pytest.approx(-0.5672093513510137, rel=rtol)  # :46,58
    C2 = np.array([1, 2, 3, 1, 2, 3, 1, 2, 3], dtype=T)
    d2 = np.array([1, 2, 2, 1, 2, 2, 1, 2, 2], dtype=np.int64)
    assert O.loglikelihood(C2, d2, dtype=T) == pytest.approx(-0.5672093513510137, rel=rtol)      # :51
    Cz = jl([[1.5, 1.5, 1.5], [3, 3, 3], [3, 3, 3]], T)
    dz = np.array([[0, 0, 0], [2, 2, 2], [2, 2, 2]], dtype=np.int64)
    assert O.loglikelihood(Cz, dz, dtype=T) == pytest.approx(-5.6344187027020260, rel=rtol)      # :66


@pytest.mark.parametrize("T,rtol", TYPES)
def test_grad_loglikelihood_forms(T, rtol):             # fitting_core_test.jl:71-126
    model = jl([[0, 0, 0], [0, 0, 0], [1, 1, 1]], T)
    Cm = jl([[1, 1, 1], [2, 2, 2], [3, 3, 3]], T)
    data = np.array([[1, 1, 1], [2, 2, 2], [2, 2, 2]], dtype=np.int64)
    r = O.grad_single(model, Cm, data, dtype=T)
    assert isinstance(r, T) and r == pytest.approx(-1, rel=rtol)          # :77-79
    mz = jl([[1, 1, 1], [0, 0, 0], [0, 0, 0]], T)
    Cz = jl([[1.5, 1.5, 1.5], [3, 3, 3], [3, 3, 3]], T)
    dz = np.array([[0, 0, 0], [2, 2, 2], [2, 2, 2]], dtype=np.int64)
    assert O.grad_single(mz, Cz, dz, dtype=T) == pytest.approx(-3, rel=rtol)   # :92-97
    # multi-model form == comprehension over the single form (:171-182)
    G, _ = O.grad_inplace(Cm, O.stack_models([model, model]), data, dtype=T)
    assert np.allclose(G, [-1, -1], rtol=rtol)                            # :99-108
    # coeffs form (:110-124)
    models = [jl([[1, 1, 1], [0, 0, 0], [0, 0, 0]], T), jl([[0, 0, 0], [1, 1, 1], [0, 0, 0]], T),
              jl([[0, 0, 0], [0, 0, 0], [1, 1, 1]], T)]
    coeffs = np.array([1.5, 3, 3], dtype=T)
    S = O.stack_models(models)
    G, _ = O.grad_inplace(O.composite(coeffs, S, dtype=T), S, data, dtype=T)
    assert G.dtype == T and G.shape == (3,) and np.allclose(G, [-1, -1, -1], rtol=rtol)


@pytest.mark.parametrize("T,rtol", TYPES)
def test_grad_inplace_zero_bins(T, rtol):               # fitting_core_test.jl:127-162
    models = [jl([[1, 1, 1], [0, 0, 0], [0, 0, 0]], T), jl([[0, 0, 0], [1, 1, 1], [0, 0, 0]], T),
              jl([[0, 0, 0], [0, 0, 0], [1, 1, 1]], T)]
    coeffs = np.array([1.5, 3, 3], dtype=T)
    S = O.stack_models(models)
    data = np.array([[1, 1, 1], [2, 2, 2], [2, 2, 2]], dtype=np.int64)
    G, resid = O.grad_inplace(S @ coeffs, S, data, dtype=T)
    assert np.allclose(G, [-1, -1, -1], rtol=rtol)                        # :137,144
    # documented side effect: composite now holds 1 - n/m (fitting_base.jl:219)
    assert np.allclose(resid, 1 - data.reshape(-1, order="F") / (S @ coeffs), rtol=rtol)
    data3 = np.array([[0, 0, 0], [2, 2, 2], [2, 2, 2]], dtype=np.int64)
    G3, _ = O.grad_inplace(S @ coeffs, S, data3, dtype=T)
    assert np.allclose(G3, [-3, -1, -1], rtol=rtol)                       # :150,157


@pytest.mark.parametrize("T,rtol", TYPES)
def test_fg(T, rtol):                                   # fitting_core_test.jl:163-195
    models = [jl([[1, 1, 1], [0, 0, 0], [0, 0, 0]], T), jl([[0, 0, 0], [1, 1, 1], [0, 0, 0]], T),
              jl([[0, 0, 0], [0, 0, 0], [1, 1, 1]], T)]
    coeffs = np.array([1.5, 3, 3], dtype=T)
    data = np.array([[1, 1, 1], [2, 2, 2], [2, 2, 2]], dtype=np.int64)
    S = O.stack_models(models)
    r, G, _ = O.fg(coeffs, S, data, dtype=T)
    assert isinstance(r, T)
    assert -r == pytest.approx(-1.4180233783775342, rel=rtol)             # :175,185
    assert np.allclose(-G, [-1, -1, -1], rtol=rtol)                       # :174,184
    r2, G2, _ = O.fg(coeffs, S, data, want_G=False, dtype=T)              # G = nothing :188-191
    assert G2 is None and -r2 == pytest.approx(-1.4180233783775342, rel=rtol)
    # G only (solvers.jl:32-34): returns nothing, G filled
    r3, G3, _ = O.fg(coeffs, S, data, want_F=False, dtype=T)
    assert r3 is None and np.allclose(G3, G)


def test_zero_sum_guard_and_nan():                      # fitting_base.jl:95 ; SURVEY 8a items 3, 11
    assert O.loglikelihood(np.zeros(4), np.zeros(4)) != 0  # all-zero -> sum of -eps, not exactly 0
    # exact zero: m == n == 1 everywhere -> every term 1-1-1*log(1) = 0 -> -Inf
    assert O.loglikelihood(np.ones(5), np.ones(5)) == -np.inf
    assert np.isnan(O.loglikelihood(np.array([np.nan, 1.0]), np.array([1.0, 1.0])))
    # clamp: m < eps -> eps
    eps = np.finfo(np.float64).eps
    assert O.loglikelihood(np.array([0.0]), np.array([0.0])) == -eps
    assert O.loglikelihood(np.array([-5.0]), np.array([2.0])) == pytest.approx(2 - eps - 2 * np.log(2 / eps))


def test_doctests_models():
    # dispersion_models.jl:63-68
    assert O.disp_gauss(1.0, 1.2, 0.2) == pytest.approx(np.exp(-0.5))
    ds, dm = O.disp_gauss_grad(1.0, 1.2, 0.2)
    assert ds == pytest.approx(3.0326532985631656) and dm == pytest.approx(-3.0326532985631656)
    # mzr.jl:245-250
    assert O.mh_mean(O.POWERLAW_MZR, 1.0, -1.0, (6.0,), 1e7) == pytest.approx(0.0, abs=1e-15)
    ga, gb, gm = O.mh_grad(O.POWERLAW_MZR, 1.0, -1.0, (6.0,), 1e8)
    assert (ga, gb) == (pytest.approx(2.0), pytest.approx(1.0)) and gm == pytest.approx(1 / 1e8 / np.log(10))
    # LinearAMR amr.jl:206-209: mu = beta + alpha (T_max - 10^(logAge-9))
    assert O.mh_mean(O.LINEAR_AMR, 0.05, -1.6, (12.0,), 9.0) == pytest.approx(-1.6 + 0.05 * 11.0)
    # LogarithmicAMR amr.jl:284: MH_from_Z of a linear Z(t); src/utilities.jl:145
    Z = 5e-5 + 1e-4 * (12.0 - 1.0)
    X = 1 - ((0.2485 + 1.78 * Z) + Z); Xs = 1 - ((0.2485 + 1.78 * 0.01524) + 0.01524)
    assert O.mh_mean(O.LOG_AMR, 1e-4, 5e-5, (12.0,), 9.0) == pytest.approx(np.log10(Z / (X * 0.01524) * Xs))


@pytest.mark.parametrize("T", [np.float32, np.float64])
def test_calculate_coeffs_properties(T):                # mzr_test.jl:9-47
    p = make_hier_problem(nj=21, nk=26, nb=8)
    R = p["R"].astype(T)
    x = O.calculate_coeffs(O.POWERLAW_MZR, 1.0, -1.0, (6.0,), 0.2, R, p["logAge"], p["MH"], dtype=T)
    assert x.dtype == T and x.shape == (p["nt"],)                         # :35-36
    rt = 1e-4 if T == np.float32 else 1e-12
    for j in range(21):
        assert x[j * 26:(j + 1) * 26].sum() == pytest.approx(R[j], rel=rt)     # :32-34
    rng = np.random.default_rng(7)
    perm = rng.permutation(21)
    uA = np.linspace(10, 8, 21)[perm]
    la = np.repeat(uA, 26); mh = np.tile(np.linspace(-2.5, 0, 26), 21)
    y = O.calculate_coeffs(O.POWERLAW_MZR, 1.0, -1.0, (6.0,), 0.2, R[perm], la, mh, dtype=T)
    order = np.argsort(-la, kind="stable")
    assert np.allclose(x, y[order], rtol=rt * 10)                         # :44
    with pytest.raises(ValueError):                                       # mzr.jl:55 argcheck
        O.calculate_coeffs(O.POWERLAW_MZR, 1.0, -1.0, (6.0,), 0.2, R[:-1], p["logAge"], p["MH"], dtype=T)


# ---- metallicity conversions on the LogarithmicAMR path: the reference's own known answers (test/utilities/utilities_test.jl:32-43) ----
def test_metallicity_conversion_kats_pin_oracle_and_host_mirror():
    import sfh_b200 as S
    # The reference asserts these at rtol 1e-7 for Float64; its literals were evidently generated from the Float32 value of 1e-3
    # (0.0010000000474974513): with that input every digit is reproduced, with the Float64 1e-3 they hold at the reference's 1e-7.
    z32 = float(np.float32(1e-3))
    assert S.Y_from_Z(z32, 0.2485) == pytest.approx(0.2502800000845455, rel=1e-15)          # :32
    assert S.X_from_Z(z32) == pytest.approx(0.748719999867957, rel=1e-15)                   # :33
    assert S.X_from_Z(1e-3, 0.25) == pytest.approx(0.74722, rel=1e-7)                       # :34
    assert S.MH_from_Z(z32, 0.01524) == pytest.approx(-1.206576807011171, rel=1e-15)        # :36
    assert S.MH_from_Z(1e-3, 0.01524) == pytest.approx(-1.206576807011171, rel=1e-7)
    assert S.Z_from_MH(-2.0, 0.01524, Y_p=0.2485) == pytest.approx(0.00016140871730361718, rel=1e-12)   # :37
    assert S.MH_from_Z(S.Z_from_MH(-2.0, 0.01524), 0.01524) == pytest.approx(-2.0, rel=1e-12)           # :39 inverses
    assert S.MH_from_Z(S.Z_from_MH(1.0, 0.01524), 0.01524) == pytest.approx(1.0, rel=1e-12)             # :41 (positive [M/H])
    assert S.dMH_dZ(1e-3, 0.01524, Y_p=0.2485, gamma=1.78) == pytest.approx(435.9070188458886, rel=1e-12)   # :43
    # dZ_dMH is the derivative of Z_from_MH (central difference) and the reciprocal of dMH_dZ at the same point
    h = 1e-6
    assert S.dZ_dMH(-1.0) == pytest.approx((S.Z_from_MH(-1.0 + h) - S.Z_from_MH(-1.0 - h)) / (2 * h), rel=1e-8)
    assert S.dZ_dMH(-1.0) * S.dMH_dZ(S.Z_from_MH(-1.0)) == pytest.approx(1.0, rel=1e-12)
    # the ORACLE's LogarithmicAMR (amr.jl:284-297) at alpha = 0, beta = Z: mean = MH_from_Z(Z), d mean / d beta = dMH_dZ(Z)
    fixed = (13.7, 0.01524, 0.2485, 1.78)
    assert O.mh_mean(O.LOG_AMR, 0.0, z32, fixed, 9.5) == pytest.approx(-1.206576807011171, rel=1e-14)
    gA, gB, _ = O.mh_grad(O.LOG_AMR, 0.0, 1e-3, fixed, 9.5)
    assert gB == pytest.approx(435.9070188458886, rel=1e-12)
    assert gA == pytest.approx(435.9070188458886 * (13.7 - 10.0 ** (9.5 - 9)), rel=1e-12)   # amr.jl:290-291
    assert np.isnan(O.mh_mean(O.LOG_AMR, 0.0, 0.5, fixed, 9.5))                             # X <= 0 -> NaN (utilities.jl:145)


def test_amr_constraint_constructors_doctests():
    import sfh_b200 as S
    a = S.LinearAMR.from_constraints((-2.5, 13.7), (-1.0, 0.0), 13.7)                       # amr.jl:221-227
    b = S.LinearAMR.from_constraints((-1.0, 0.0), (-2.5, 13.7), 13.7)
    assert a == b and a.alpha == pytest.approx(1.5 / 13.7) and a.beta == pytest.approx(-2.5)
    assert a(9 + np.log10(13.7)) == pytest.approx(-2.5) and a(-np.inf) == pytest.approx(-1.0)   # passes through both constraints
    c = S.LogarithmicAMR.from_constraints((-2.5, 13.7), (-1.0, 0.0), 13.7)                  # amr.jl:308-315
    d = S.LogarithmicAMR.from_constraints((-1.0, 0.0), (-2.5, 13.7), 13.7)
    assert c == d and c(9 + np.log10(13.7)) == pytest.approx(-2.5, rel=1e-10) and c(-np.inf) == pytest.approx(-1.0, rel=1e-10)
    with pytest.raises(ValueError):
        S.LinearAMR.from_constraints((-2.5, 5.0), (-1.0, 5.0))                              # identical times
    with pytest.raises(ValueError):
        S.LinearAMR.from_constraints((-1.0, 13.7), (-2.5, 0.0))                             # metallicity decreasing towards the present
    with pytest.raises(ValueError):
        S.LogarithmicAMR.from_constraints((-2.5, 1.0), (-1.0, 0.0), 13.7)                   # Z < 0 at T_max
