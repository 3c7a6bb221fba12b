"""Pins the oracle's MZR/AMR chain rules (mzr.jl:84-215, amr.jl:78-173).

The reference pins these with ForwardDiff goldens on StableRNG inputs (mzr_test.jl:74-80,
amr_test.jl:45-47,264-265) which cannot be regenerated without Julia.  We assert the SAME thing on
our inputs: analytic gradient == automatic derivative of the forward model.  The derivative is
taken by complex-step differentiation of an INDEPENDENT numpy forward model (no subtractive
cancellation => exact to round-off, like ForwardDiff), plus the structural tests of
mzr_test.jl:82-112 (stacked == unstacked is trivial here; permutation invariance) and :116-176
(Jacobian / fixed-parameter identities of logdensity_and_gradient).
"""
import numpy as np
import pytest

import oracle as O
from conftest import make_hier_problem

EPS = np.finfo(np.float64).eps
DEFAULT_FIXED = {O.POWERLAW_MZR: (6.0,), O.LINEAR_AMR: (12.0,), O.LOG_AMR: (12.0, 0.01524, 0.2485, 1.78)}
TRUE_PARAMS = {O.POWERLAW_MZR: (1.0, -2.0, 0.2),      # mzr_test.jl:53-54
               O.LINEAR_AMR: (0.05, -1.6, 0.2),       # amr_test.jl LinearAMR(0.05, -1.6, 12)
               O.LOG_AMR: (1e-4, 5e-5, 0.2)}          # amr_test.jl LogarithmicAMR(1e-4, 5e-5, 12)


def forward_nlogL(kind, fixed, v, M, data, logAge, MH):
    """Independent numpy forward model of -logL(variables); complex-safe."""
    v = np.asarray(v)
    nj = v.shape[0] - 3
    R, alpha, beta, sigma = v[:nj], v[nj], v[nj + 1], v[nj + 2]
    _, first = np.unique(logAge, return_index=True)
    uniq = logAge[np.sort(first)]                                  # first-appearance order
    jidx = np.array([np.nonzero(uniq == a)[0][0] for a in logAge])
    if kind == O.POWERLAW_MZR:
        s = np.argsort(-uniq, kind="stable")
        cum = np.empty(nj, dtype=v.dtype)
        cum[s] = np.cumsum(R[s])
        mu = beta + alpha * (np.log10(cum) - fixed[0])
    else:
        age = 10.0 ** (uniq - 9)
        if kind == O.LINEAR_AMR:
            mu = beta + alpha * (fixed[0] - age)
        else:
            Z = beta + alpha * (fixed[0] - age)
            solZ, Yp, gam = fixed[1:4]
            X = 1 - ((Yp + gam * Z) + Z)
            Xs = 1 - ((Yp + gam * solZ) + solZ)
            mu = np.log10(Z / (X * solZ) * Xs)
    A = np.exp(-(((MH - mu[jidx]) / sigma) ** 2) / 2)
    Asum = np.zeros(nj, dtype=v.dtype)
    np.add.at(Asum, jidx, A)
    r = R[jidx] * A / Asum[jidx]
    m = M @ r
    m = np.where(m.real < EPS, EPS + 0 * m, m)
    safe = np.where(data > 0, data, 1.0)
    term = np.where(data > 0, data - m - data * np.log(safe / m), -m)
    return -term.sum()


def complex_step_grad(kind, fixed, v, M, data, logAge, MH):
    g = np.empty(v.shape[0])
    for i in range(v.shape[0]):
        vc = v.astype(np.complex128)
        h = 1e-30 * max(1.0, abs(v[i]))
        vc[i] += 1j * h
        g[i] = forward_nlogL(kind, fixed, vc, M, data, logAge, MH).imag / h
    return g


def build(kind, shuffle=False, ragged=False, noisy=True, nb=300):
    la_hi, la_lo = (10.0, 8.0)
    p = make_hier_problem(nj=21, nk=26, nb=nb, shuffle=shuffle, ragged=ragged, la_hi=la_hi, la_lo=la_lo)
    a, b, s = TRUE_PARAMS[kind]
    fixed = DEFAULT_FIXED[kind]
    x = O.calculate_coeffs(kind, a, b, fixed, s, p["R"], p["logAge"], p["MH"])
    lam = p["M"] @ x
    data = p["rng"].poisson(lam).astype(np.float64) if noisy else lam
    v = np.concatenate([p["R"], [a, b, s]])
    return p, fixed, v, data


@pytest.mark.parametrize("kind", [O.POWERLAW_MZR, O.LINEAR_AMR, O.LOG_AMR])
@pytest.mark.parametrize("shuffle,ragged", [(False, False), (True, False), (True, True)])
def test_chain_rule_vs_complex_step(kind, shuffle, ragged):
    p, fixed, v, data = build(kind, shuffle, ragged)
    nl, G, fullG = O.fg_hier(kind, fixed, (1, 1, 1), v, p["M"], data, p["logAge"], p["MH"])
    assert nl == pytest.approx(forward_nlogL(kind, fixed, v, p["M"], data, p["logAge"], p["MH"]).real, rel=1e-13)
    ref = complex_step_grad(kind, fixed, v, p["M"], data, p["logAge"], p["MH"])
    scale = np.abs(ref) + 1e-14 * np.abs(ref).max()
    assert np.max(np.abs(G - ref) / scale) < 1e-9, np.max(np.abs(G - ref) / scale)
    # the double oracle against its own __float128 instantiation (arbiter)
    nlq, Gq, fullGq = O.fg_hier(kind, fixed, (1, 1, 1), v, p["M"], data, p["logAge"], p["MH"], quad=True)
    assert nl == pytest.approx(nlq, rel=1e-13)
    assert np.max(np.abs(G - Gq) / scale) < 1e-9
    # fullG is the un-flipped d logL / d r_jk (mzr.jl:129) == -G of flat fg!
    _, Gflat, _ = O.fg(O.calculate_coeffs(kind, v[-3], v[-2], fixed, v[-1], v[:-3], p["logAge"], p["MH"]), p["M"], data)
    assert np.allclose(fullG, -Gflat, rtol=1e-12, atol=0)


@pytest.mark.parametrize("kind", [O.POWERLAW_MZR, O.LINEAR_AMR])
def test_age_permutation_invariance(kind):                # mzr_test.jl:87-112
    p, fixed, v, data = build(kind)
    nj, nk = 21, 26
    nl, G, _ = O.fg_hier(kind, fixed, (1, 1, 1), v, p["M"], data, p["logAge"], p["MH"])
    perm = np.random.default_rng(3).permutation(nj)
    cols = np.concatenate([np.arange(j * nk, (j + 1) * nk) for j in perm])
    v2 = np.concatenate([v[:nj][perm], v[nj:]])
    nl2, G2, _ = O.fg_hier(kind, fixed, (1, 1, 1), v2, p["M"][:, cols], data, p["logAge"][cols], p["MH"][cols])
    assert nl2 == pytest.approx(nl, rel=1e-13)
    assert np.allclose(G2, np.concatenate([G[:nj][perm], G[nj:]]), rtol=1e-9)


@pytest.mark.parametrize("kind", [O.POWERLAW_MZR, O.LINEAR_AMR, O.LOG_AMR])
def test_fixed_parameters_get_zero_gradient(kind):        # mzr.jl:196,201 ; mzr_test.jl:133-176
    p, fixed, v, data = build(kind)
    _, Gall, _ = O.fg_hier(kind, fixed, (1, 1, 1), v, p["M"], data, p["logAge"], p["MH"])
    for free in [(1, 1, 0), (1, 0, 1), (0, 1, 1), (0, 0, 0)]:
        _, G, _ = O.fg_hier(kind, fixed, free, v, p["M"], data, p["logAge"], p["MH"])
        assert np.array_equal(G[:-3], Gall[:-3])
        for i in range(3):
            assert G[-3 + i] == (Gall[-3 + i] if free[i] else 0.0)


def test_logdensity_and_gradient_identities():            # mzr_test.jl:115-176
    kind = O.POWERLAW_MZR
    p, fixed, v, data = build(kind)
    nj = 21
    a, b, s = TRUE_PARAMS[kind]
    nl, G, _ = O.fg_hier(kind, fixed, (1, 1, 1), v, p["M"], data, p["logAge"], p["MH"])
    tv = np.concatenate([np.log(p["R"]), [np.log(a), b, np.log(s)]])               # :116
    Gt = np.concatenate([G[:nj] * p["R"], [G[nj] * a, G[nj + 1], G[nj + 2] * s]])    # :118-119
    lp, gr = O.hier_logdensity_and_gradient(kind, fixed, (1, 1, 1), (a, b, s), tv, p["M"], data,
                                            p["logAge"], p["MH"], jacobian_corrections=False)
    assert lp == pytest.approx(-nl, rel=1e-14) and np.allclose(gr, -Gt, rtol=1e-12)  # :124-125
    lpj, grj = O.hier_logdensity_and_gradient(kind, fixed, (1, 1, 1), (a, b, s), tv, p["M"], data,
                                              p["logAge"], p["MH"], jacobian_corrections=True)
    assert lpj == pytest.approx(-nl + np.log(p["R"]).sum() + np.log(a) + np.log(s), rel=1e-14)  # :129-131
    # sigma fixed (:133-152): xvec loses its last entry, gradient has nj+2 entries
    tvf = tv[:-1]
    lpf, grf = O.hier_logdensity_and_gradient(kind, fixed, (1, 1, 0), (a, b, s), tvf, p["M"], data,
                                              p["logAge"], p["MH"], jacobian_corrections=False)
    assert lpf == pytest.approx(-nl, rel=1e-14) and grf.shape == (nj + 2,)
    assert np.allclose(grf, -Gt[:-1], rtol=1e-12)
    # beta (MH0) fixed (:154-176)
    tvb = np.concatenate([tv[:nj], [tv[nj], tv[nj + 2]]])
    lpb, grb = O.hier_logdensity_and_gradient(kind, fixed, (1, 0, 1), (a, b, s), tvb, p["M"], data,
                                              p["logAge"], p["MH"], jacobian_corrections=False)
    assert np.allclose(grb, -np.concatenate([Gt[:nj], [Gt[nj], Gt[nj + 2]]]), rtol=1e-12)


def test_hmc_and_mcmc_adapters():                         # hmc_sample.jl:24-37 ; mcmc_sample.jl:12-23
    from conftest import make_flat_problem
    M, x, data = make_flat_problem(500, 12)
    logx = np.log(x)
    lp, g = O.hmc_logdensity_and_gradient(logx, M, data)
    nl, G, _ = O.fg(x, M, data)
    assert lp == pytest.approx(-nl + logx.sum()) and np.allclose(g, -G * x + 1)
    # finite-difference sanity of the transformed gradient
    h = 1e-6
    for i in (0, 5, 11):
        e = np.zeros(12); e[i] = h
        fd = (O.hmc_logdensity_and_gradient(logx + e, M, data)[0] - O.hmc_logdensity_and_gradient(logx - e, M, data)[0]) / (2 * h)
        assert fd == pytest.approx(g[i], rel=1e-5, abs=1e-4)
    X = np.stack([x, x * 1.1, np.where(np.arange(12) == 3, -1e-300, x), np.zeros(12)], axis=1)
    out = O.mcmc_logl(X, M, data)
    assert out[0] == pytest.approx(-nl) and out[2] == -np.inf and np.isfinite(out[1])
    assert out[3] == pytest.approx(O.loglikelihood(np.zeros(500), data))   # exact zeros are legal (:15-19)
