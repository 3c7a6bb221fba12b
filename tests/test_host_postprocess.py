"""Host-side helpers either side of the hot path, pinned by the reference's own golden values
(test/fitting/fitting_core_test.jl:196-247).  No device needed."""
import types

import numpy as np
import pytest

import sfh_b200 as S
from sfh_b200 import solvers as V


@pytest.mark.parametrize("T", [np.float32, np.float64])
def test_construct_x0_golden(T):                                   # fitting_core_test.jl:196-212
    rtol = 1e-3 if T == np.float32 else 1e-7
    want = np.tile([0.015015015015015015, 0.15015015015015015, 1.5015015015015016], 3)
    la = np.tile(np.array([1, 2, 3], dtype=T), 3)
    got = S.construct_x0(la, 1e-5, normalize_value=5)
    assert np.allclose(got, want, rtol=rtol) and got.sum() == pytest.approx(5, rel=rtol)
    got = S.construct_x0(la[::-1], 1e-5, normalize_value=5)       # no sorting assumed
    assert np.allclose(got, want[::-1], rtol=rtol) and got.sum() == pytest.approx(5, rel=rtol)
    with pytest.raises(ValueError):
        S.construct_x0(la, 1e-9)


@pytest.mark.parametrize("T", [np.float32, np.float64])
def test_calculate_cum_sfr_golden(T):                              # fitting_core_test.jl:213-247
    coeffs, la, mh = np.array([1, 2, 2, 4], T), np.array([1, 2, 1, 2], T), np.array([-2, -2, -1, -1], T)
    r = S.calculate_cum_sfr(coeffs, la, mh, 1e-6, normalize_value=1, sorted=False)
    assert np.array_equal(r[0], [1, 2]) and np.allclose(r[1], [1, 2 / 3]) and np.allclose(r[2], [1 / 30, 2 / 300])
    assert np.allclose(r[3], [-4 / 3, -4 / 3])
    r = S.calculate_cum_sfr(coeffs, la, mh, 1e-6, normalize_value=5)
    assert np.allclose(r[1], [1, 2 / 3]) and np.allclose(r[2], [5 / 30, 10 / 300]) and np.allclose(r[3], [-4 / 3, -4 / 3])
    r = S.calculate_cum_sfr(coeffs, np.array([1, 1, 2, 2], T), np.array([-2, -1, -2, -1], T), 1e-6, sorted=True)
    assert np.array_equal(r[0], [1, 2]) and np.allclose(r[1], [1, 2 / 3]) and np.allclose(r[2], [1 / 30, 2 / 300])
    assert np.allclose(r[3], [-4 / 3, -4 / 3])
    # an age bin with zero mass inherits the previous bin's mean metallicity (utilities.jl:178-184)
    r = S.calculate_cum_sfr(np.array([1.0, 3.0, 0.0, 0.0]), np.array([1.0, 1.0, 2.0, 2.0]), np.array([-2.0, -1.0, -2.0, -1.0]), 1e-6)
    assert np.allclose(r[3], [-1.25, -1.25]) and np.allclose(r[1], [1.0, 0.0])


def _fake_result(nj, rng, free=(True, True), dfree=(True,)):
    mz, dp = S.PowerLawMZR(1.0, -1.5, 6.0, free), S.GaussianDispersion(0.2, dfree)
    R = rng.random(nj) * 1e5 + 1e4
    nfree = sum(free) + sum(dfree)
    x = np.concatenate([np.log(R), S.logtransform(np.array([1.0, -1.5, 0.2]), np.array([1, 0, 1]))[np.array(free + dfree)]])
    A = rng.standard_normal((nj + nfree, nj + nfree)) * 0.02
    invH = A @ A.T + 1e-4 * np.eye(nj + nfree)
    mu = np.concatenate([R, [1.0, -1.5, 0.2]])
    return V.BFGSResult(mu, np.sqrt(np.diag(invH))[:nj + 3] if nfree == 3 else np.zeros(nj + 3), invH,
                        types.SimpleNamespace(x=x), mz, dp)


def test_rand_and_cum_sfr_quantiles_shapes():                      # utilities.jl:239-302 (doctest shapes :226-236)
    rng = np.random.default_rng(3)
    nj, nk = 12, 9
    la = np.repeat(np.linspace(10.0, 8.9, nj), nk)
    mh = np.tile(np.linspace(-2.0, 0.0, nk), nj)
    res = _fake_result(nj, rng)
    smp = res.rand(rng, 500)
    assert smp.shape == (nj + 3, 500) and np.all(smp[:nj] > 0) and np.all(smp[nj] > 0) and np.all(smp[-1] > 0)
    assert np.allclose(np.median(smp, axis=1), res.mu, rtol=0.1)
    q = (0.16, 0.5, 0.84)
    out = S.cum_sfr_quantiles({"map": res, "mle": res}, la, mh, 13.7, 400, q, rng=rng)
    for k in ("cum_sfh", "sfrs", "mean_mh"):
        assert out[k].shape == (nj, 3) and np.all(np.diff(out[k], axis=1) >= 0)
    assert out["samples"].shape == (nj + 3, 400) and out["n_good"] == 400
    assert np.allclose(out["cum_sfh"][0], 1.0) and np.all(np.diff(out["cum_sfh"][:, 1]) <= 1e-12)
    # a fixed parameter is written in at its value for every draw (bfgs_result.jl:68-74)
    resf = _fake_result(nj, rng, free=(True, False), dfree=(False,))
    smp = resf.rand(rng, 50)
    assert np.all(smp[nj + 1] == -1.5) and np.all(smp[nj + 2] == 0.2) and np.all(smp[nj] > 0)


def test_construct_x0_mdf_doctests():                             # hierarchical/construct_x0_mdf.jl:19-55
    want = [4.504504504504504, 0.4504504504504504, 0.04504504504504504]
    assert np.allclose(S.construct_x0_mdf([9.0, 8.0, 7.0], 10.0, normalize_value=5.0), want)
    assert np.allclose(S.construct_x0_mdf(np.repeat([9.0, 8.0, 7.0, 8.0], 3), 10.0, normalize_value=5.0), want)
    assert np.allclose(S.construct_x0_mdf(np.tile([9.0, 8.0, 7.0, 8.0], 3), 10.0, normalize_value=5.0),
                       S.construct_x0([9.0, 8.0, 7.0], 10.0, normalize_value=5.0))
    assert np.allclose(S.construct_x0_mdf([9.0, 8.0, 7.0], [0.9009, 0.99099, 1.0], 10.0, normalize_value=5.0), [4.5045, 0.4504, 0.0450], atol=1e-3)
    assert np.allclose(S.construct_x0_mdf([9.0, 8.0, 7.0], [0.1, 0.5, 1.0], 10.0, normalize_value=5.0), [0.5, 2.0, 2.5])
    assert np.allclose(S.construct_x0_mdf([7.0, 8.0, 9.0], [1.0, 0.5, 0.1], 10.0, normalize_value=5.0), [2.5, 2.0, 0.5])
    a = S.construct_x0_mdf([9.0, 8.0, 7.0], [0.9009, 0.99099, 1.0], 10.0, normalize_value=5.0)
    assert np.allclose(S.construct_x0_mdf([9.0, 8.0, 7.0], [[9.0, 8.0, 7.0], [0.9009, 0.99099, 1.0]], 10.0, normalize_value=5.0), a)
    assert np.allclose(S.construct_x0_mdf([9.0, 8.0, 7.0], [[9.0, 8.5, 8.25, 7.0], [0.9009, 0.945945, 0.9887375, 1.0]], 10.0, normalize_value=5.0), a)
    la = np.repeat(np.linspace(6.6, 10.1, 36), 5)                  # fixed_amr_test.jl:35-53
    assert S.construct_x0_mdf(la, 13.7).sum() == pytest.approx(1) and S.construct_x0_mdf(la, 13.7, normalize_value=1e5).sum() == pytest.approx(1e5)
    with pytest.raises(ValueError):
        S.construct_x0_mdf([9.0, 8.0, 7.0], [0.5, 0.1, 1.0], 10.0)                                  # not monotonic
    with pytest.raises(ValueError):
        S.construct_x0_mdf([9.0, 10.2], 10.0)


def test_truncate_relweights():                                   # fixed_amr.jl:229-243, fixed_amr_test.jl:76-104
    la = np.array([10.0, 10.0, 10.0, 9.0, 9.0, 9.0, 9.0])
    rw = np.array([0.02, 0.5, 0.48, 0.7, 0.2, 0.05, 0.05])
    assert np.array_equal(S.truncate_relweights(0, rw, la), np.arange(7))
    assert np.array_equal(S.truncate_relweights(0.1, rw, la), [1, 2, 3, 4])
    assert np.array_equal(S.truncate_relweights(0.05, rw, la), [1, 2, 3, 4, 5, 6])
    # the doc example (:199-224): 11 metallicities around a Gaussian MDF keep 3 templates at relweightsmin = 0.1
    mhs = np.arange(-2.5, 0.01, 0.25)
    w = np.exp(-0.5 * ((mhs + 2.0) / 0.2) ** 2); w /= w.sum()
    assert np.allclose(w[:5], [0.021919934465195145, 0.2284109622221623, 0.4988954088848224, 0.2284109622221623, 0.021919934465195145])
    assert np.array_equal(S.truncate_relweights(0.1, w, np.full(11, 10.0)), [1, 2, 3])


def test_tau_doctests():                                          # fitting/utilities.jl:358-411 (the `tau` doctest block)
    ula, max_la = [8.0, 8.5, 9.0, 9.5, 10.0], 10.13
    assert S.tau(0.5, ula, max_la, [1.0, 0.8, 0.5, 0.2, 0.1]) == 1.0                       # :365-366 exact knot
    cum = [1.0, 0.8, 0.4, 0.2, 0.1]
    assert S.tau(0.5, ula, max_la, cum) == pytest.approx(0.8290569415042095, rel=1e-14)    # :374
    assert np.allclose(S.tau([0.5, 0.75], ula, max_la, cum), [0.8290569415042095, 0.40169929526473325], rtol=1e-14)   # :381
    assert S.tau(0.5, ula[::-1], max_la, cum[::-1]) == pytest.approx(0.8290569415042095, rel=1e-14)                    # :388
    upper, lower = [1.0, 0.85, 0.6, 0.3, 0.15], [1.0, 0.75, 0.3, 0.1, 0.05]
    got = S.tau(0.5, ula, max_la, cum, lower, upper)                                                                   # :400
    assert got.shape == (1, 3) and np.allclose(got, [[0.6961012293408169, 0.8290569415042095, 1.7207592200561264]], rtol=1e-14)
    got = S.tau([0.5, 0.75], ula, max_la, cum, lower, upper)                                                           # :407
    assert np.allclose(got, [[0.6961012293408169, 0.8290569415042095, 1.7207592200561264],
                             [0.31622776601683794, 0.40169929526473325, 0.5897366596101027]], rtol=1e-14)


def test_tau_interp_edges():                                      # fitting/utilities.jl:311-336
    ula, max_la = np.array([8.0, 8.5, 9.0, 9.5, 10.0]), 10.13
    itp = S.tau_interp(ula, max_la, [1.0, 0.8, 0.4, 0.2, 0.1])
    assert itp(0.0) == pytest.approx(10 ** 10.13 / 1e9) and itp(1.0) == pytest.approx(0.1)    # no mass at max_logAge; all of it at the youngest bin
    assert np.all(np.diff(itp.knots) > 0) and itp.knots[-1] == 1.0                            # :330 normalised to its maximum
    assert np.allclose(S.tau_interp(ula, max_la, 7.0 * np.array([1.0, 0.8, 0.4, 0.2, 0.1]))([0.3, 0.9]), itp([0.3, 0.9]), rtol=1e-14)
    for bad in (-0.1, 1.0001, np.nan):                                                        # Gridded(Linear()) throws outside its knots
        with pytest.raises(ValueError):
            itp(bad)
    # equal cumulative values (bins that formed no mass) are separated by one ulp each (:332), so the interpolant stays defined
    flat = S.tau_interp(ula, max_la, [1.0, 0.5, 0.5, 0.5, 0.1])
    assert np.all(np.diff(flat.knots) > 0) and flat.knots[3] == np.nextafter(0.5, 1) and flat.knots[4] == np.nextafter(np.nextafter(0.5, 1), 1)
    assert flat(0.3) == pytest.approx(np.interp(0.3, [0.1, 0.5], [10.0, 10 ** 0.5]))
    with pytest.raises(ValueError):                                                           # :312
        S.tau_interp(ula, max_la, [1.0, 0.5])
    with pytest.raises(ValueError):                                                           # :313
        S.tau_interp(ula, 10.0, [1.0, 0.8, 0.4, 0.2, 0.1])
    with pytest.raises(ValueError):                                                           # :319
        S.tau_interp(ula, max_la, [1.0, 0.3, 0.4, 0.2, 0.1])
    with pytest.raises(ValueError):                                                           # :320 ages not in the order of cum_sfh
        S.tau_interp(ula[::-1], max_la, [1.0, 0.8, 0.4, 0.2, 0.1])


def test_tau_from_result():                                       # fitting/utilities.jl:456-468
    rng = np.random.default_rng(5)
    nj, nk = 12, 9
    la = np.repeat(np.linspace(10.0, 8.9, nj), nk)
    mh = np.tile(np.linspace(-2.0, 0.0, nk), nj)
    res = _fake_result(nj, rng)
    t = S.tau({"map": res, "mle": res}, [0.5, 0.9], la, mh, 10.13, Nsamples=300, rng=rng)
    assert t.shape == (2, 3)
    # a LOWER cumulative-mass quantile reaches the fraction later (smaller lookback time): columns are ordered old <- young
    # exactly as the reference's (lower, best, upper) = (q16, q50, q84) of the cumulative SFH
    assert np.all(t[:, 0] <= t[:, 1]) and np.all(t[:, 1] <= t[:, 2])
    assert np.all(t[0] >= t[1]) and np.all(t > 10 ** 8.9 / 1e9 - 1e-12) and np.all(t < 10 ** 10.13 / 1e9)
    with pytest.raises(ValueError):
        S.tau({"map": res, "mle": res}, 0.5, la, mh, 10.13, Nsamples=10, q=(0.5,), rng=rng)


def test_result_accessors_and_calculate_coeffs_from_result():     # bfgs_result.jl:38-41, 81-90, 108-111, 151-154
    rng = np.random.default_rng(11)
    nj, nk = 7, 5
    la = np.repeat(np.linspace(10.0, 8.8, nj), nk)
    mh = np.tile(np.linspace(-2.0, 0.0, nk), nj)
    mle, mp = _fake_result(nj, rng), _fake_result(nj, rng)
    assert len(mle) == nj + 3 and mle.mode() is mle.mu and mle.median() is mle.mu and mle.std() is mle.sigma
    pair = {"map": mp, "mle": mle}
    assert S.result_mode(pair) is mle.mu and S.result_median(pair) is mle.mu and S.result_std(pair) is mp.sigma
    assert S.result_mode(mle) is mle.mu and S.result_std(mle) is mle.sigma
    want = S.calculate_coeffs(mle.MH_model, mle.disp_model, mle.mu[:nj], la, mh)
    assert np.array_equal(S.calculate_coeffs(mle, la, mh), want)
    assert np.array_equal(S.calculate_coeffs(pair, la, mh), want)          # the MLE of a CompositeBFGSResult
    got = S.calculate_coeffs(mle, la, mh).reshape(nj, nk).sum(axis=1)      # mzr_test.jl:32-34: the masses are conserved per age
    assert np.allclose(got, mle.mu[:nj], rtol=1e-13)


def test_transformations_and_nparams_doctests():                 # transformations.jl:13-16, :39-42 ; hierarchical_models.jl:12-15
    import math
    assert np.allclose(S.logtransform((0.5, -1.0, 1.0), (1, 0, 1)), [math.log(0.5), -1.0, 0.0], rtol=0, atol=0)
    assert np.allclose(S.exptransform((math.log(0.5), -1.0, 0.0), (1, 0, 1)), [0.5, -1.0, 1.0], rtol=1e-16)
    # the -1 branch (x' = log(-x), inverse -exp(x'); transformations.jl:24-25, :50-51) and the round trip
    p, tf = np.array([0.3, -2.5, -0.7, 4.0]), np.array([1, 0, -1, 1])
    assert S.logtransform(p, tf)[2] == math.log(0.7)
    assert np.allclose(S.exptransform(S.logtransform(p, tf), tf), p, rtol=1e-15)
    assert S.nparams(S.LinearAMR(1.0, 1.0), S.GaussianDispersion(0.2)) == 3


def test_calculate_coeffs_doctest():                              # generic_fitting.jl:13-26
    n_logage, n_mh = 10, 20
    R = np.random.default_rng(0).random(n_logage)
    coeffs = S.calculate_coeffs(S.PowerLawMZR(1.0, -1.0), S.GaussianDispersion(0.2), R,
                                np.repeat(np.linspace(7.0, 10.0, n_logage), n_mh), np.tile(np.linspace(-2.0, 0.0, n_mh), n_logage))
    assert isinstance(coeffs, np.ndarray) and coeffs.dtype == np.float64 and coeffs.shape == (n_logage * n_mh,)
    assert np.allclose(coeffs.reshape(n_logage, n_mh).sum(axis=1), R, rtol=1e-13)      # mzr_test.jl:32-34
