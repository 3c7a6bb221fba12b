"""Property-based checks of the oracle's flat path (composite -> Poisson logL -> gradient; fitting_base.jl:55-65, 84-96, 265-285;
solvers.jl:20-38) over randomly drawn ragged and degenerate shapes: against an independent numpy restatement, against its own
__float128 instantiation, and through the size-independent identities the GPU tests use at BASELINE sizes
(x.G = sum(m - n), bin / template permutations, the all-threads variant).  No device needed."""
import numpy as np
import pytest
from hypothesis import given, settings, strategies as st

import oracle as O

EPS = float(np.finfo(np.float64).eps)


def numpy_fg(x, M, data):
    m = M @ x                                                              # fitting_base.jl:60
    mc = np.maximum(m, EPS)                                                # :90
    with np.errstate(divide="ignore", invalid="ignore"):
        terms = np.where(data > 0, data - mc - data * np.log(data / mc), -mc)   # :92
    logl = terms.sum()
    r = 1.0 - data / mc                                                    # :279
    return (-logl if logl != 0 else np.inf), M.T @ r, m                     # :95 ; solvers.jl:28-31


@st.composite
def problems(draw):
    nb = draw(st.integers(1, 70))
    nt = draw(st.integers(1, 24))
    seed = draw(st.integers(0, 2 ** 31 - 1))
    zero_frac = draw(st.sampled_from([0.0, 0.3, 1.0]))      # share of zero-count bins (all of them: the -m branch only)
    zero_coeff = draw(st.booleans())
    fractional = draw(st.booleans())                        # noise-free "data" is not integer (mzr_test.jl:66)
    rng = np.random.default_rng(seed)
    M = np.asfortranarray(rng.random((nb, nt)) * draw(st.sampled_from([1e-5, 1.0, 1e3])))
    x = rng.random(nt) * 10
    if zero_coeff:
        x[rng.integers(0, nt)] = 0.0
    lam = M @ x
    data = lam * (1 + 0.1 * rng.standard_normal(nb)) if fractional else rng.poisson(np.minimum(lam, 1e8)).astype(np.float64)
    data = np.abs(data)
    data[rng.random(nb) < zero_frac] = 0.0
    return M, x, data


@settings(max_examples=200, deadline=None, derandomize=True)
@given(problems())
def test_oracle_fg_matches_numpy_and_quad(p):
    M, x, data = p
    f, G, left = O.fg(x, M, data)
    comp = O.composite(x, M)
    fn, Gn, mn = numpy_fg(x, M, data)
    fq, Gq, gscale, compq = O.fg_quad(x, M, data)
    resid = 1.0 - data / np.maximum(mn, EPS)
    # backward-error scale of each component: the sum over bins of |M_ij r_i|, plus what the rounding of m_i itself does to
    # r_i = 1 - n_i / m_i (a near-total cancellation where the model fits: delta r_i ~ eps n_i / m_i)
    scale = np.abs(M).T @ (np.abs(resid) + data / np.maximum(mn, EPS))
    assert np.allclose(comp, mn, rtol=1e-13, atol=0) and np.allclose(compq, mn, rtol=1e-13, atol=0)
    # the reference's documented side effect: after fg! with a gradient, `composite` holds the residual (fitting_base.jl:219);
    # after an F-only call it still holds the composite (solvers.jl:35-36)
    assert np.allclose(left, resid, rtol=1e-9, atol=1e-13)
    assert np.array_equal(O.fg(x, M, data, want_G=False)[2], comp)
    terms = np.abs(M @ x).sum() + np.abs(data).sum() + 1.0                 # logL is a sum of nb terms of this size
    assert abs(f - fn) <= 1e-13 * terms and abs(f - fq) <= 1e-13 * terms
    assert np.all(np.abs(G - Gn) <= 1e-12 * scale + 1e-300)
    assert np.all(np.abs(G - Gq) <= 1e-12 * scale + 1e-300)
    assert np.allclose(gscale, np.abs(M).T @ np.abs(resid), rtol=1e-6, atol=1e-300)   # the arbiter's own scale: sum_i |M_ij r_i|


@settings(max_examples=200, deadline=None, derandomize=True)
@given(problems())
def test_oracle_identities(p):
    M, x, data = p
    f, G, _ = O.fg(x, M, data)
    comp = O.composite(x, M)
    # checksum of checksums: x . G = sum_i m_i r_i = sum(m - n) wherever the clamp is inactive
    if np.all(comp > EPS):
        lhs, rhs = float(x @ G), float((comp - data).sum())
        assert abs(lhs - rhs) <= 1e-11 * (np.abs(comp).sum() + np.abs(data).sum())
    # F-only and G-only requests give the same numbers as the joint one (solvers.jl:32-36)
    f_only = O.fg(x, M, data, want_G=False)[0]
    assert f_only == f
    assert np.array_equal(O.fg(x, M, data, want_F=False)[1], G)
    # permuting the bins leaves logL and G unchanged up to re-association; permuting templates permutes G
    rng = np.random.default_rng(0)
    pb, pt = rng.permutation(M.shape[0]), rng.permutation(M.shape[1])
    f2, G2, _ = O.fg(x, np.asfortranarray(M[pb]), data[pb])
    tol = 1e-12 * (np.abs(comp).sum() + np.abs(data).sum() + 1.0)
    assert abs(f2 - f) <= tol
    f3, G3, _ = O.fg(x[pt], np.asfortranarray(M[:, pt]), data)
    assert abs(f3 - f) <= tol
    mc = np.maximum(comp, EPS)
    scale = np.abs(M).T @ (np.abs(1.0 - data / mc) + data / mc) + 1e-300   # see test_oracle_fg_matches_numpy_and_quad
    assert np.all(np.abs(G2 - G) <= 1e-12 * scale) and np.all(np.abs(G3 - G[pt]) <= 1e-12 * scale[pt])
    # the all-host-threads variant (the bench's cpu_baseline) agrees with the scalar one
    fo, Go = O.fg_omp(x, M, data)[:2]
    assert abs(fo - f) <= tol and np.all(np.abs(Go - G) <= 1e-12 * scale)
    # ... and so does the BLAS route (gemv 'N' / loops / gemv 'T' through OpenBLAS: what Julia's mul! calls), the other baseline
    fb, Gb = O.fg_blas(x, M, data)
    assert (abs(fb - f) <= tol or fb == f) and np.all(np.abs(Gb - G) <= 1e-12 * scale)


def test_mcmc_mask_and_hmc_adapter_properties():                  # mcmc_sample.jl:12-23 ; hmc_sample.jl:24-37
    rng = np.random.default_rng(7)
    M = np.asfortranarray(rng.random((40, 6)))
    data = rng.poisson(M @ (rng.random(6) * 20)).astype(np.float64)
    X = np.asfortranarray(rng.random((6, 9)) * 10)
    X[2, 3] = -1e-300          # any negative entry, however small, masks the walker; exact zeros do not
    X[4, 5] = 0.0
    ll = O.mcmc_logl(X, M, data)
    assert ll[3] == -np.inf and np.isfinite(np.delete(ll, 3)).all()
    for w in (0, 5, 8):
        assert ll[w] == pytest.approx(-O.fg(X[:, w], M, data, want_G=False)[0], rel=1e-14)
    th = np.log(X[:, 0])
    lp, g = O.hmc_logdensity_and_gradient(th, M, data)
    f, G, _ = O.fg(X[:, 0], M, data)
    assert lp == pytest.approx(-f + th.sum(), rel=1e-14) and np.allclose(g, -G * X[:, 0] + 1, rtol=1e-13)
