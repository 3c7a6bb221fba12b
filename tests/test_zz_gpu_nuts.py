"""GPU tests of the native multi-chain NUTS (include/sfhcuda.h: sfh_hmc_sample_nuts, sfh_sample_sfh_nuts): the reference's own
assertions of hmc_sample / sample_sfh / tsample_sfh (shapes, positivity, fixed rows: basic_linear_combinations.jl:156-186,
mzr_test.jl:218-251) with the chains and their batching running inside the library, plus agreement with the Python engine: the
native chains follow `nuts_chain` driven by the same Philox stream on the same device log-density (first draws, before rounding
differences between the batched and single-vector kernels can grow), and their posterior moments match the fit."""
import numpy as np
import pytest

from conftest import make_flat_problem, make_hier_problem
from nuts_ref import PhiloxRng

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def S():
    import sfh_b200
    return sfh_b200


def test_hmc_sample_native(S):                                    # basic_linear_combinations.jl:156-186
    rng = np.random.Generator(np.random.Philox(9))
    N = 6
    x = rng.random(N) * 100 + 20
    models = [rng.random((40, 40)) for _ in range(N)]
    data = rng.poisson(sum(c * m for c, m in zip(x, models))).astype(np.int64)
    out = S.hmc_sample(models, data, 150, nchains=4, nwarmup=100, rng=np.random.default_rng(3), engine="native")
    assert out.shape == (150, N, 4) and np.all(out > 0) and np.all(np.isfinite(out))
    ft = S.fit_templates(models, data, x0=np.ones(N))
    z = np.abs(out.mean(axis=(0, 2)) - ft["map"].mu) / ft["map"].sigma
    assert np.all(z < 1.0), z                                      # posterior mean within 1 sigma of the MAP


def test_native_chain_follows_python_engine_on_the_device_logdensity(S):
    from sfh_b200.solvers import _nuts_call, nuts_sample, renormalize_x0
    M, x, data = make_flat_problem(3000, 8, seed=77)
    ds = S.DeviceStack(M, data)
    model = S.HMCModel(ds, None, data)
    th0 = np.log(renormalize_x0(data, ds, x))
    st, res, stats = _nuts_call(S._lib.lib.sfh_hmc_sample_nuts, (ds.ctx().handle,), [th0, th0], 4, 6, 5, 0.8, None, 1234, None)
    assert st == 0 and stats.n_evals > stats.n_batches > 0
    for c in range(2):
        s, lp, step = nuts_sample(model.logdensity_and_gradient, th0, 4, 6, 5, 0.8, PhiloxRng(1234, c))
        # same stream, same algorithm; the two sides evaluate fg! with different kernels (batched DMMA vs fused), so allow the
        # 1e-13-level differences a few leapfrogs of amplification
        np.testing.assert_allclose(res[c][0][0], s[0], rtol=1e-4, atol=1e-6)
        assert res[c][2] == pytest.approx(step, rel=1e-4)


def test_sample_sfh_and_tsample_sfh_native(S):                     # mzr_test.jl:218-251 (shapes, fixed rows)
    p = make_hier_problem(nj=10, nk=12, nb=4000)
    mz, dp = S.PowerLawMZR(1.0, -2.0, 6.0), S.GaussianDispersion(0.2, (False,))
    xt = S.calculate_coeffs(mz, dp, p["R"], p["logAge"], p["MH"])
    data = np.random.default_rng(3).poisson(p["M"] @ xt).astype(np.float64)
    ds = S.DeviceStack(p["M"], data)
    res = S.fit_sfh(S.PowerLawMZR(1.1, -2.1, 6.0), dp, ds, data, p["logAge"], p["MH"], x0=p["R"] * 1.2)
    one = S.sample_sfh(res, ds, data, p["logAge"], p["MH"], 60, eps=0.2, rng=np.random.default_rng(1), engine="native")
    assert one["posterior_matrix"].shape == (13, 60) and np.all(one["posterior_matrix"][-1] == 0.2)   # fixed sigma row
    assert np.all(np.isfinite(one["logp"])) and one["step_size"] > 0
    many = S.tsample_sfh(res, ds, data, p["logAge"], p["MH"], 130, eps=0.2, rng=np.random.default_rng(2), chain_length=20, engine="native")
    pm = many["posterior_matrix"]
    assert pm.shape == (13, 130) and np.all(pm[-1] == 0.2) and np.all(pm[:10] > 0) and many["logp"].shape == (130,)
    z = np.abs(pm.mean(axis=1)[:12] - res["mle"].mu[:12]) / np.maximum(res["map"].sigma[:12], 1e-12)
    assert np.all(z < 3), z
    # the native log-density equals the host adapter's at the sampled points (same device fg!, transforms done natively)
    opt = S.HierarchicalOptimizer(res["mle"].MH_model, res["mle"].disp_model, ds, data, p["logAge"], p["MH"], True, True, True)
    Z = np.log(pm[:10, :5])
    par = np.stack([np.log(pm[10, :5]), pm[11, :5]])              # alpha (log-transformed), beta; sigma is fixed
    for k in range(5):
        lp, _ = opt.logdensity_and_gradient(np.concatenate([Z[:, k], par[:, k]]))
        assert lp == pytest.approx(many["logp"][k], rel=1e-9)
