"""The library's native multi-chain NUTS (csrc/sfh_nuts.h; include/sfhcuda.h: sfh_nuts_run, sfh_sample_sfh_nuts_generic) on the
CPU through its batched log-density callback.

Sampler trajectories are not pinned by the reference (third-party engine; its tests check shapes), so the
checker here is the Python engine `nuts_chain` -- the same algorithm with the same order of random draws -- driven by the same
Philox streams (tests/nuts_ref.py): the native chains must reproduce it draw for draw.  Plus: posterior moments, batching
statistics, failure modes, and the HierarchicalOptimizer log-density (generic_fitting.jl:90-199) against the oracle's restatement.
"""
import ctypes as C

import numpy as np
import pytest

import oracle as O
import sfh_b200
from nuts_ref import PhiloxRng
from sfh_b200.solvers import native_nuts, native_sample_sfh_generic, nuts_sample

L = sfh_b200._lib


def gaussian(prec, mean):
    def single(th):
        d = th - mean
        return -0.5 * d @ prec @ d, -prec @ d

    def batch(Th):
        D = Th - mean[:, None]
        return -0.5 * np.einsum("ic,ij,jc->c", D, prec, D), -prec @ D
    return single, batch


def make_target(n=5, seed=0):
    rng = np.random.default_rng(seed)
    A = rng.standard_normal((n, n))
    cov = A @ A.T / n + 0.3 * np.eye(n)
    return np.linalg.inv(cov), rng.standard_normal(n), cov


@pytest.mark.parametrize("mass", ["identity", "diag", "dense"])
@pytest.mark.parametrize("warm,eps0", [(30, None), (0, 0.3)])
def test_native_chains_reproduce_the_python_engine(mass, warm, eps0):
    prec, mean, cov = make_target()
    single, batch = gaussian(prec, mean)
    n, seed = 5, 77
    inv_mass = {"identity": None, "diag": np.diag(cov).copy(), "dense": cov}[mass]
    starts = [mean + 0.5 * k * np.ones(n) for k in range(3)]
    lens = [12, 7, 9]
    res, stats = native_nuts(batch, starts, lens, nwarmup=warm, max_depth=6, eps0=eps0, seed=seed, inv_mass=inv_mass)
    nev = 0
    for c in range(3):
        cnt = [0]

        def counted(th):
            cnt[0] += 1
            return single(th)
        s, lp, step = nuts_sample(counted, starts[c], lens[c], warm, 6, 0.8, PhiloxRng(seed, c), inv_mass, eps0)
        nev += cnt[0]
        assert res[c][0].shape == (lens[c], n)
        np.testing.assert_allclose(res[c][0], s, rtol=1e-7, atol=1e-9)      # (rounding of dot products differs; decisions do not)
        np.testing.assert_allclose(res[c][1], lp, rtol=1e-7, atol=1e-9)
        assert res[c][2] == pytest.approx(step, rel=1e-7)
    assert stats.n_evals == nev                                   # every chain asked for exactly the evaluations it would alone
    assert stats.n_batches < nev and stats.n_batches >= nev / 3   # ... served in shared rounds


def test_posterior_moments():
    prec, mean, cov = make_target(n=4, seed=3)
    _, batch = gaussian(prec, mean)
    res, stats = native_nuts(batch, [mean.copy() for _ in range(8)], 400, nwarmup=150, max_depth=6, seed=5)
    S = np.concatenate([r[0] for r in res], axis=0)
    assert S.shape == (3200, 4)
    se = np.sqrt(np.diag(cov) / 400)                              # generous: treats each chain's draws as ~50 independent ones
    assert np.all(np.abs(S.mean(axis=0) - mean) < 4 * se)
    np.testing.assert_allclose(np.cov(S.T), cov, atol=0.25 * np.abs(cov).max())
    assert all(0.05 < r[2] < 5 for r in res)                      # adapted step sizes are sane
    assert stats.n_evals / stats.n_batches > 6                    # 8 chains share nearly every batched pass


def test_many_ragged_chains_do_not_deadlock():
    """64 chain threads with lengths 0..7 (some finish after their first evaluations, all at different times): every round still
    serves exactly the chains that are alive, and the run terminates."""
    prec, mean, cov = make_target(n=3, seed=8)
    _, batch = gaussian(prec, mean)
    sizes = []

    def counting(Th):
        sizes.append(Th.shape[1])
        return batch(Th)
    lens = [k % 8 for k in range(64)]
    res, stats = native_nuts(counting, [mean + 0.01 * k for k in range(64)], lens, nwarmup=3, max_depth=4, seed=11)
    assert [r[0].shape[0] for r in res] == lens and stats.n_batches == len(sizes) and stats.n_evals == sum(sizes)
    assert sizes[0] == 64 and sizes[-1] < 64 and all(np.all(np.isfinite(r[0])) for r in res)


def test_failure_modes():
    prec, mean, cov = make_target()
    _, batch = gaussian(prec, mean)

    def boom(Th):
        raise FloatingPointError("log-density failed")
    with pytest.raises(FloatingPointError):
        native_nuts(boom, [mean, mean], 5, nwarmup=2)

    calls = [0]

    def fail_late(Th):                                             # fails while chains are parked mid-trajectory: must not hang
        calls[0] += 1
        if calls[0] > 20:
            raise FloatingPointError("late failure")
        return batch(Th)
    with pytest.raises(FloatingPointError):
        native_nuts(fail_late, [mean, mean + 1, mean - 1], 50, nwarmup=10)
    bad = -np.eye(5)
    with pytest.raises(ValueError):                                # inv_mass not positive definite
        native_nuts(batch, [mean], 3, inv_mass=bad)
    # -inf log-density regions are rejected, not propagated
    def walled(Th):
        lp, g = batch(Th)
        out = np.where(Th[0] > mean[0] + 0.5, -np.inf, lp)
        return out, g
    res, _ = native_nuts(walled, [mean - 0.2], 100, nwarmup=30, seed=2)
    assert np.all(res[0][0][:, 0] <= mean[0] + 0.5) and np.all(np.isfinite(res[0][1]))
    # raw ABI
    o = L.sfh_nuts_opts(); o.struct_size = 5
    x = np.zeros(2); n1 = np.array([1], dtype=np.int64)
    dp, ip = C.POINTER(C.c_double), C.POINTER(C.c_int64)
    cb = L.sfh_batch_logdensity_fn(lambda *a: 0)
    assert L.lib.sfh_nuts_run(cb, None, 2, 1, x.ctypes.data_as(dp), n1.ctypes.data_as(ip), None, C.byref(o), x.ctypes.data_as(dp),
                              x.ctypes.data_as(dp), None, None, None) == L.SFH_ERR_INVALID_ARG
    assert L.lib.sfh_nuts_run(L.sfh_batch_logdensity_fn(), None, 2, 1, x.ctypes.data_as(dp), n1.ctypes.data_as(ip), None, None,
                              x.ctypes.data_as(dp), x.ctypes.data_as(dp), None, None, None) == L.SFH_ERR_INVALID_ARG
    assert L.lib.sfh_hmc_sample_nuts(None, 1, None, None, None, None, None, None, None, None, None) == L.SFH_ERR_INVALID_ARG
    assert C.sizeof(L.sfh_nuts_opts) == 48


@pytest.mark.parametrize("free3", [(True, True, True), (True, False, True)])
def test_hierarchical_logdensity_against_the_oracle_adapter(free3):
    """sfh_sample_sfh_nuts_generic: the transformed, Jacobian-corrected log-density of generic_fitting.jl:90-199 evaluated natively
    around the ORACLE's hierarchical fg!, checked (through the chains it produces) against the Python engine running on the oracle's
    own restatement of logdensity_and_gradient."""
    from conftest import make_hier_problem
    kind, fixed = O.POWERLAW_MZR, [6.0]
    P = make_hier_problem(nj=5, nk=6, nb=250)
    nj = P["nj"]
    par = np.array([1.0, -2.0, 0.2])
    coeffs = O.calculate_coeffs(kind, par[0], par[1], fixed, par[2], P["R"], P["logAge"], P["MH"])
    data = P["rng"].poisson(P["M"] @ coeffs).astype(np.float64)
    free = np.array(free3, dtype=bool)
    tf = np.array(O.TRANSFORMS[kind])

    def inner_batched(V):
        nl, G = np.empty(V.shape[1]), np.empty(V.shape, order="F")
        for c in range(V.shape[1]):
            f, g, _ = O.fg_hier(kind, fixed, free3, V[:, c], P["M"], data, P["logAge"], P["MH"])
            nl[c], G[:, c] = f, g
        return nl, G

    def single(xv):
        return O.hier_logdensity_and_gradient(kind, fixed, free3, par, xv, P["M"], data, P["logAge"], P["MH"], True)
    x0 = np.concatenate([np.log(P["R"]), np.array([np.log(par[0]), par[1], np.log(par[2])])[free]])
    starts = [x0 + 0.01 * k for k in range(3)]
    nx = x0.shape[0]
    inv_mass = np.full(nx, 1e-3)
    res, stats = native_sample_sfh_generic(inner_batched, nj, par, tf, free, starts, [6, 4, 5], nwarmup=4, max_depth=4, eps0=0.05, seed=9,
                                           inv_mass=inv_mass)
    for c in range(3):
        s, lp, _ = nuts_sample(single, starts[c], [6, 4, 5][c], 4, 4, 0.8, PhiloxRng(9, c), inv_mass, 0.05)
        np.testing.assert_allclose(res[c][0], s, rtol=1e-8, atol=1e-10)
        np.testing.assert_allclose(res[c][1], lp, rtol=1e-9)
