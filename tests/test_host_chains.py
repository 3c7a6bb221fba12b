"""Host logic of the multi-chain driver: sampler coroutines sharing one batched evaluation per round of requests
(the reference runs chains on threads, each with its own fg!, hmc_sample.jl:123-141).  CPU only: the batch function is a closed-form Gaussian."""
import numpy as np
import pytest

import sfh_b200 as S
from sfh_b200 import solvers as V


def _gauss(prec):
    def single(th):
        return float(-0.5 * th @ (prec * th)), -(prec * th)

    def batch(TH):                       # column by column with the very same arithmetic, so trajectories match bit for bit
        cols = [single(TH[:, c]) for c in range(TH.shape[1])]
        return np.array([lp for lp, _ in cols]), np.stack([g for _, g in cols], axis=1)
    return single, batch


def test_batched_chains_follow_the_sequential_trajectories():
    d, nchains = 5, 4
    prec = np.array([1.0, 4.0, 0.25, 9.0, 2.0])
    single, batch = _gauss(prec)
    th0 = [np.full(d, 0.3 * (c + 1)) for c in range(nchains)]
    seq = [V.nuts_sample(single, th0[c], 60, 40, 6, rng=np.random.default_rng(100 + c)) for c in range(nchains)]
    res, b = V.run_chains_batched(batch, th0, 60, 40, 6, [np.random.default_rng(100 + c) for c in range(nchains)])
    for c in range(nchains):
        assert np.array_equal(res[c][0], seq[c][0])
        assert res[c][2] == seq[c][2]
    # the requests really were grouped: far fewer batches than evaluations
    assert b.n_evals > 2 * b.n_batches and b.n_batches > 0
    # sanity of the target: pooled variance ~ 1/prec
    pooled = np.concatenate([r[0] for r in res])
    assert np.all(np.abs(pooled.var(axis=0) * prec - 1) < 0.8)


def test_chains_of_different_length_and_errors():
    prec = np.ones(3)
    single, batch = _gauss(prec)
    # chains finish at different times (a chain that is done must not block the others)
    lens = [5, 40, 17]

    def chain(th0, nsteps, nwarmup, max_depth, rng=None):
        return V.nuts_chain(th0 * 0 + 0.1, lens[int(th0[0])], 5, 4, rng=rng)
    res, _ = V.run_chains_batched(batch, [np.full(3, float(c)) for c in range(3)], 0, chain=chain,
                                  rngs=[np.random.default_rng(c) for c in range(3)])
    assert [r[0].shape[0] for r in res] == lens

    def bad(TH):
        raise RuntimeError("device failure")
    with pytest.raises(RuntimeError, match="device failure"):
        V.run_chains_batched(bad, [np.zeros(3)] * 2, 5, 5, 3, [np.random.default_rng(c) for c in range(2)])


def test_philox_restatement_known_answers():
    """The host Philox4x32-10 used to restate the device sampler reproduces the Random123 known-answer vectors."""
    from ensemble_ref import philox4x32_10, philox_u01
    kats = [((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
            ((0xffffffff,) * 4, (0xffffffff,) * 2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
            ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0), (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1))]
    for ctr, key, want in kats:
        got = philox4x32_10(*[np.array([v]) for v in ctr], *key)
        assert tuple(int(v[0]) for v in got) == want
    u = philox_u01(np.arange(200000), 987654321, 16)
    assert 0 <= u.min() and u.max() < 1 and abs(u.mean() - 0.5) < 5e-3 and abs(u.var() - 1 / 12) < 2e-3


def test_host_restatement_of_the_ensemble_sampler_samples_a_gaussian():
    from ensemble_ref import stretch_move_reference
    prec = np.array([1.0, 4.0, 0.25])
    logl = lambda X: -0.5 * np.einsum("iw,i,iw->w", X, prec, X)
    X0 = np.random.default_rng(0).standard_normal((3, 40))
    chain, lps, Xf, lp, acc = stretch_move_reference(logl, X0, 600, 2, 2.0, seed=77)
    assert chain.shape == (300, 3, 40) and 0.2 < acc < 0.9
    v = chain[100:].transpose(1, 0, 2).reshape(3, -1).var(axis=1)
    assert np.all(np.abs(v * prec - 1) < 0.35)


def test_dense_mass_nuts_samples_a_correlated_gaussian():
    """nuts_chain with a dense M^-1 (what sample_sfh passes: MAP.invH, generic_fitting.jl:479-482) and a fixed step size."""
    rng = np.random.default_rng(5)
    A = rng.standard_normal((4, 4))
    cov = A @ A.T + 0.5 * np.eye(4)
    prec = np.linalg.inv(cov)
    lg = lambda th: (float(-0.5 * th @ prec @ th), -(prec @ th))
    s, lps, eps = V.nuts_sample(lg, np.zeros(4), 1500, 0, 6, rng=np.random.default_rng(6), inv_mass=cov, eps0=0.6)
    assert eps == 0.6 and s.shape == (1500, 4)                     # no warm-up: the step size stays the one given
    emp = np.cov(s[200:].T)
    assert np.all(np.abs(emp - cov) < 0.35 * np.sqrt(np.outer(np.diag(cov), np.diag(cov))))
    # a diagonal inverse mass given as a vector or as a diagonal matrix walks the same trajectory
    d = np.array([1.0, 4.0, 0.25, 2.0])
    a = V.nuts_sample(lg, np.zeros(4), 40, 10, 5, rng=np.random.default_rng(7), inv_mass=d)
    b = V.nuts_sample(lg, np.zeros(4), 40, 10, 5, rng=np.random.default_rng(7), inv_mass=np.diag(d))
    assert np.allclose(a[0], b[0], rtol=1e-9, atol=1e-12)


def test_expand_posterior_transforms_and_fixed_rows():
    """Transformed samples -> natural units over all variables with fixed parameters written in (generic_fitting.jl:640-658)."""
    import types
    mz, dp = S.PowerLawMZR(1.0, -1.5, 6.0, (True, False)), S.GaussianDispersion(0.2, (True,))
    best = V.BFGSResult(np.array([10.0, 20.0, 1.0, -1.5, 0.2]), np.zeros(5), np.eye(4), types.SimpleNamespace(x=np.zeros(4)), mz, dp)
    Z = np.array([[0.0, 1.0], [np.log(3.0), 0.0], [np.log(2.0), np.log(0.5)], [np.log(0.3), np.log(0.1)]])   # R1, R2, alpha, sigma
    out = V._expand_posterior(best, Z)
    assert out.shape == (5, 2)
    assert np.allclose(out[0], [1.0, np.e]) and np.allclose(out[1], [3.0, 1.0]) and np.allclose(out[2], [2.0, 0.5])
    assert np.all(out[3] == -1.5) and np.allclose(out[4], [0.3, 0.1])
