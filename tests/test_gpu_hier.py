"""GPU parity tests of the hierarchical path (calculate_coeffs, the MZR/AMR fg! chain rules, the
HierarchicalOptimizer / HMCModel / MCMCModel adapters) and of the batched-walker kernel, through the
C-ABI, against the CPU oracle (itself pinned by complex-step differentiation: test_oracle_chainrule.py).

Structure follows test/fitting/mzr_test.jl and amr_test.jl: fg! value + gradient (:78-80), stacked layout
(:82-84), age-permutation invariance (:87-112), logdensity_and_gradient with / without Jacobian and with
sigma or MH0 fixed (:115-176).

Tolerances (BASELINE.json north_star): 1e-12 relative on logL; every gradient component within 1e-10 of its BACKWARD-ERROR
SCALE  sum_t |d r_t / d p| * sum_i |M_it r_i|  (conftest.hier_grad_scale), the chain-rule image of test_gpu_core's per-template
scale -- the parameter components (alpha, beta, sigma) are near-total cancellations at the optimum, so |G_p| itself is no
yardstick.  Measured on B200 (profiles/r2_hier_parity_errors.txt): max |dG| / scale between 1e-16 and 4e-14 over all cases.
"""
import numpy as np
import pytest

import oracle as O
from conftest import assert_hier_grad_close, hier_grad_scale, make_flat_problem, make_hier_problem

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def S():
    import sfh_b200
    assert sfh_b200.device_count() >= 1
    return sfh_b200


def models_for(S, kind):
    return {O.POWERLAW_MZR: (S.PowerLawMZR(1.0, -2.0, 6.0), (6.0,)),
            O.LINEAR_AMR: (S.LinearAMR(0.05, -1.6, 12.0), (12.0,)),
            O.LOG_AMR: (S.LogarithmicAMR(1e-4, 5e-5, 12.0), (12.0, 0.01524, 0.2485, 1.78))}[kind]


def build(S, kind, shuffle=False, ragged=False, nb=1500, noisy=True, nj=21, nk=26):
    p = make_hier_problem(nj=nj, nk=nk, nb=nb, shuffle=shuffle, ragged=ragged)
    model, fixed = models_for(S, kind)
    x = O.calculate_coeffs(kind, model.alpha, model.beta, fixed, 0.2, p["R"], p["logAge"], p["MH"])
    lam = p["M"] @ x
    data = p["rng"].poisson(lam).astype(np.float64) if noisy else lam
    v = np.concatenate([p["R"], [model.alpha, model.beta, 0.2]])
    return p, model, fixed, v, data


KINDS = [O.POWERLAW_MZR, O.LINEAR_AMR, O.LOG_AMR]


@pytest.mark.parametrize("kind", KINDS)
@pytest.mark.parametrize("shuffle,ragged", [(False, False), (True, False), (True, True)])
def test_calculate_coeffs_device(S, kind, shuffle, ragged):      # mzr.jl:50-79 / amr.jl:50-73
    p, model, fixed, v, data = build(S, kind, shuffle, ragged)
    ds = S.DeviceStack(p["M"], data)
    got = S.calculate_coeffs(model, S.GaussianDispersion(0.2), p["R"], p["logAge"], p["MH"], models=ds)
    want = O.calculate_coeffs(kind, model.alpha, model.beta, fixed, 0.2, p["R"], p["logAge"], p["MH"])
    assert np.allclose(got, want, rtol=1e-12, atol=0)
    uniq = p["logAge"][np.sort(np.unique(p["logAge"], return_index=True)[1])]
    for j, a in enumerate(uniq):                                  # sum_k r_jk == R_j  (mzr_test.jl:32-34)
        assert got[p["logAge"] == a].sum() == pytest.approx(p["R"][j], rel=1e-12)


@pytest.mark.parametrize("kind", KINDS)
@pytest.mark.parametrize("shuffle,ragged", [(False, False), (True, False), (True, True)])
def test_fg_hier_parity(S, kind, shuffle, ragged):                # mzr_test.jl:78-84 / amr_test.jl:69-106
    p, model, fixed, v, data = build(S, kind, shuffle, ragged)
    disp = S.GaussianDispersion(0.2)
    ds = S.DeviceStack(p["M"], data)
    nlq, Gq, _ = O.fg_hier(kind, fixed, (1, 1, 1), v, p["M"], data, p["logAge"], p["MH"], quad=True)
    nlo, Go, _ = O.fg_hier(kind, fixed, (1, 1, 1), v, p["M"], data, p["logAge"], p["MH"])
    G = np.empty(v.shape[0])
    nl = S.fg_(True, G, model, disp, v, ds, data, None, p["logAge"], p["MH"])
    assert nl == pytest.approx(nlq, rel=1e-12)
    scale = hier_grad_scale(kind, fixed, v, p["M"], data, p["logAge"], p["MH"])
    assert_hier_grad_close(G, Gq, scale, what="device vs __float128 arbiter")
    assert_hier_grad_close(Go, Gq, scale, what="double oracle vs __float128 arbiter")      # the restatement obeys the same bar
    # F only (G === nothing)  mzr_test.jl:78
    assert S.fg_(True, None, model, disp, v, ds, data, None, p["logAge"], p["MH"]) == pytest.approx(nlq, rel=1e-12)
    # perturbed start (mzr_test.jl:180-182 uses 1.5x) -- away from the optimum plain relative error holds
    v2 = v.copy(); v2[:-3] *= 1.5; v2[-3] *= 1.2; v2[-1] *= 1.3
    nlq2, Gq2, _ = O.fg_hier(kind, fixed, (1, 1, 1), v2, p["M"], data, p["logAge"], p["MH"], quad=True)
    G2 = np.empty(v.shape[0])
    assert S.fg_(True, G2, model, disp, v2, ds, data, None, p["logAge"], p["MH"]) == pytest.approx(nlq2, rel=1e-12)
    assert_hier_grad_close(G2, Gq2, hier_grad_scale(kind, fixed, v2, p["M"], data, p["logAge"], p["MH"]), what="perturbed start")


@pytest.mark.parametrize("kind", [O.POWERLAW_MZR, O.LINEAR_AMR])
def test_age_permutation_invariance(S, kind):                     # mzr_test.jl:87-112
    p, model, fixed, v, data = build(S, kind)
    disp = S.GaussianDispersion(0.2)
    nj, nk = 21, 26
    G = np.empty(nj + 3)
    nl = S.fg_(True, G, model, disp, v, S.DeviceStack(p["M"], data), data, None, p["logAge"], p["MH"])
    perm = np.random.default_rng(3).permutation(nj)
    cols = np.concatenate([np.arange(j * nk, (j + 1) * nk) for j in perm])
    v2 = np.concatenate([v[:nj][perm], v[nj:]])
    G2 = np.empty(nj + 3)
    nl2 = S.fg_(True, G2, model, disp, v2, S.DeviceStack(p["M"][:, cols], data), data, None, p["logAge"][cols], p["MH"][cols])
    assert nl2 == pytest.approx(nl, rel=1e-12)
    scale = hier_grad_scale(kind, fixed, v, p["M"], data, p["logAge"], p["MH"])
    assert_hier_grad_close(G2, np.concatenate([G[:nj][perm], G[nj:]]), np.concatenate([scale[:nj][perm], scale[nj:]]), what="age permutation")


@pytest.mark.parametrize("kind", KINDS)
def test_free_masks_and_logdensity(S, kind):                      # mzr_test.jl:115-176
    p, model0, fixed, v, data = build(S, kind)
    nj = 21
    ds = S.DeviceStack(p["M"], data)
    a, b, s = model0.alpha, model0.beta, 0.2
    tf = list(model0.transforms()) + [1]
    for free in [(True, True, True), (True, True, False), (True, False, True), (False, True, True)]:
        model = type(model0)(a, b, fixed[0], free[:2]) if kind != O.LOG_AMR else S.LogarithmicAMR(a, b, fixed[0], free[:2])
        disp = S.GaussianDispersion(s, (free[2],))
        G = np.empty(nj + 3)
        nl = S.fg_(True, G, model, disp, v, ds, data, None, p["logAge"], p["MH"])
        nlo, Go, _ = O.fg_hier(kind, fixed, [int(f) for f in free], v, p["M"], data, p["logAge"], p["MH"])
        assert nl == pytest.approx(nlo, rel=1e-12)
        for i in range(3):
            if not free[i]:
                assert G[nj + i] == 0.0                           # mzr.jl:175,196,201
        scale = hier_grad_scale(kind, fixed, v, p["M"], data, p["logAge"], p["MH"])
        assert_hier_grad_close(G, Go, scale, what=f"free={free}")
        pars = np.array([a, b, s]); fm = np.array(free)
        tpar = np.array([np.log(pv) if t == 1 else pv for pv, t in zip(pars, tf)])
        xvec = np.concatenate([np.log(p["R"]), tpar[fm]])
        for jac in (False, True):
            opt = S.HierarchicalOptimizer(model, disp, ds, data, p["logAge"], p["MH"], True, True, jac)
            assert opt.dimension() == nj + fm.sum()
            lp, gr = opt.logdensity_and_gradient(xvec)
            lpo, gro = O.hier_logdensity_and_gradient(kind, fixed, [int(f) for f in free], (a, b, s), xvec, p["M"], data,
                                                      p["logAge"], p["MH"], jacobian_corrections=jac)
            assert lp == pytest.approx(lpo, rel=1e-12)
            # transformed variables: d/d(log R_j) = R_j d/dR_j, d/d(log p) = p d/dp (generic_fitting.jl:150-165): the scale follows
            tscale = np.concatenate([scale[:nj] * p["R"], (scale[nj:] * np.array([pv if t == 1 else 1.0 for pv, t in zip(pars, tf)]))[fm]])
            assert gr.shape == (nj + fm.sum(),)
            assert_hier_grad_close(gr, gro, tscale, what=f"HierarchicalOptimizer free={free} jacobian={jac}")
        # F-only and G-only protocol (generic_fitting.jl:172-178,195-197)
        assert S.HierarchicalOptimizer(model, disp, ds, data, p["logAge"], p["MH"], True, None, True).logdensity_and_gradient(xvec) == pytest.approx(lp, rel=1e-13)


def test_hier_errors(S):
    p, model, fixed, v, data = build(S, O.POWERLAW_MZR, nb=200)
    ds = S.DeviceStack(p["M"], data)
    disp = S.GaussianDispersion(0.2)
    with pytest.raises(ValueError):                               # mzr.jl:55
        S.fg_(True, None, model, disp, v[:-1], ds, data, None, p["logAge"], p["MH"])
    with pytest.raises(ValueError):                               # mzr.jl:57
        S.fg_(True, None, model, disp, v, ds, data, None, p["logAge"], p["MH"][:-1])
    with pytest.raises(ValueError):                               # mzr.jl:126
        S.fg_(True, np.empty(3), model, disp, v, ds, data, None, p["logAge"], p["MH"])


def test_hier_full_config3_shape(S):
    """BASELINE config 3: 60 ages x 40 [M/H] = 2400 templates; bins reduced so the oracle finishes in seconds."""
    p = make_hier_problem(nj=60, nk=40, nb=3000, la_hi=10.1, la_lo=6.6)
    model, disp = S.PowerLawMZR(1.0, -2.0, 6.0), S.GaussianDispersion(0.2)
    x = O.calculate_coeffs(O.POWERLAW_MZR, 1.0, -2.0, (6.0,), 0.2, p["R"], p["logAge"], p["MH"])
    data = p["rng"].poisson(p["M"] @ x).astype(np.float64)
    v = np.concatenate([p["R"], [1.0, -2.0, 0.2]]) * 1.1
    ds = S.DeviceStack(p["M"], data)
    G = np.empty(63)
    nl = S.fg_(True, G, model, disp, v, ds, data, None, p["logAge"], p["MH"])
    nlq, Gq, _ = O.fg_hier(O.POWERLAW_MZR, (6.0,), (1, 1, 1), v, p["M"], data, p["logAge"], p["MH"], quad=True)
    assert nl == pytest.approx(nlq, rel=1e-12)
    assert_hier_grad_close(G, Gq, hier_grad_scale(O.POWERLAW_MZR, (6.0,), v, p["M"], data, p["logAge"], p["MH"]), what="config-3 grid")


# ------------------------------------------------------------------ sampler adapters
def test_hmc_model(S):                                            # hmc_sample.jl:24-37
    M, x, data = make_flat_problem(4000, 64, seed=21)
    hm = S.HMCModel(M, None, data)
    assert hm.dimension() == 64
    logx = np.log(x) + 0.01
    lp, g = hm.logdensity_and_gradient(logx)
    lpo, go = O.hmc_logdensity_and_gradient(logx, M, data)
    assert lp == pytest.approx(lpo, rel=1e-12)
    # d/d(log x_j) = x_j d/dx_j (hmc_sample.jl:30-35): 1e-10 of the flat backward-error scale times x_j
    xe = np.exp(logx)
    gscale = xe * (np.abs(M).T @ np.abs(1.0 - data / np.maximum(M @ xe, np.finfo(np.float64).eps)))
    assert np.all(np.abs(g - go) <= 1e-10 * gscale), float(np.max(np.abs(g - go) / gscale))
    assert hm.logdensity(logx) == pytest.approx(lpo, rel=1e-12)


@pytest.mark.parametrize("nb,nt,W", [(4000, 64, 1), (5000, 100, 37), (9801, 142, 256), (40000, 500, 128)])
def test_batched_walkers(S, nb, nt, W):                           # mcmc_sample.jl:12-23
    M, x, data = make_flat_problem(nb, nt, seed=31)
    rng = np.random.default_rng(5)
    X = np.maximum(0.0, x[:, None] + rng.standard_normal((nt, W)))   # examples/fitting1.ipynb cell 143
    if W > 2:
        X[3, 1] = -1e-9                                           # one negative entry -> -Inf (:15-19)
        X[:, 2] = 0.0                                             # exact zeros are legal
    mm = S.MCMCModel(M, data)
    got = mm.batch(X)
    want = O.mcmc_logl(X, M, data)
    fin = np.isfinite(want)
    assert np.array_equal(np.isfinite(got), fin)
    assert np.array_equal(got[~fin], want[~fin])                  # -Inf exactly
    assert np.allclose(got[fin], want[fin], rtol=1e-12, atol=0)
    assert mm(X[:, 0]) == pytest.approx(want[0], rel=1e-12)
    # agrees with the fused single-vector path
    nl, _, _ = mm.models.eval_fg(X[:, 0], want_G=False)
    assert got[0] == pytest.approx(-nl, rel=1e-13)


def test_batched_walkers_f32_stack(S):
    M, x, data = make_flat_problem(6000, 120, seed=33, dtype=np.float32)
    X = np.maximum(0.0, x[:, None] + np.random.default_rng(6).standard_normal((120, 40)))
    got = S.MCMCModel(M, data).batch(X)
    for w in (0, 7, 39):
        nlq, _, _ = O.fg_quad_f32(X[:, w], M, data, want_G=False)
        assert got[w] == pytest.approx(-nlq, rel=1e-6)


def test_batched_epilogue_fast_and_exact_paths(S):
    """The DMMA kernel's branch-free Poisson epilogue (table log, Newton reciprocal) and the exact libdevice path it falls
    back to outside its domain give the single-vector kernel's answer: zero counts, exact m == n (logL = 0 -> -Inf,
    fitting_base.jl:95), tiny / huge / non-integer "counts", NaN coefficients."""
    rng = np.random.default_rng(12)
    nb, nt, W = 1500, 24, 40
    M = np.asfortranarray(rng.random((nb, nt)))
    x = 50 * rng.random(nt) + 1
    X = np.maximum(0.0, x[:, None] * (1 + 0.3 * rng.standard_normal((nt, W))))
    cases = {}
    d = rng.poisson(M @ x).astype(np.float64); d[::7] = 0.0
    cases["poisson with zero bins"] = d
    cases["fractional, ratios from 1e-6 to 1e6"] = (M @ x) * 10.0 ** rng.uniform(-6, 6, nb)
    d = rng.poisson(M @ x).astype(np.float64); d[5] = 1e-200; d[9] = 1e200; d[11] = 3e-310
    cases["outside the fast domain (exact path)"] = d
    for name, data in cases.items():
        ds = S.DeviceStack(M, data)
        got = ds.eval_logl_batched(X)
        for w in (0, 17, W - 1):
            nl, _, _ = ds.eval_fg(X[:, w], want_G=False)
            assert got[w] == pytest.approx(-nl, rel=1e-12), name
            nlq, _, _, _ = O.fg_quad(X[:, w], M, data, want_G=False)
            assert got[w] == pytest.approx(-nlq, rel=1e-12), name
        nlb, G = ds.eval_fg_batched(X[:, :9])
        nl1, G1, _ = ds.eval_fg(X[:, 3])
        assert nlb[3] == pytest.approx(nl1, rel=1e-12) and np.all(np.abs(G[:, 3] - G1) <= 1e-10 * np.abs(G1).max()), name
    # m == n in every bin: each term is exactly 0, the sum is 0 and the guard turns it into -Inf
    Mi = np.asfortranarray(np.eye(40)); di = np.arange(1.0, 41.0)
    dsi = S.DeviceStack(Mi, di)
    Xi = np.stack([di, di * 1.5, di], axis=1)
    out = dsi.eval_logl_batched(Xi)
    assert out[0] == -np.inf and out[2] == -np.inf and np.isfinite(out[1])
    # NaN coefficient: that walker (only) is NaN
    Xn = X[:, :6].copy(); Xn[4, 2] = np.nan
    got = S.DeviceStack(M, cases["poisson with zero bins"]).eval_logl_batched(Xn)
    assert np.isnan(got[2]) and np.all(np.isfinite(np.delete(got, 2)))


@pytest.mark.parametrize("nb,nt,C,dtype", [(4000, 64, 1, np.float64), (5003, 100, 5, np.float64), (9801, 142, 16, np.float64),
                                           (20000, 600, 33, np.float64), (3000, 200, 70, np.float64), (6000, 120, 12, np.float32),
                                           (7000, 90, 24, np.float64), (4100, 64, 9, np.float32), (2500, 33, 64, np.float32)])
def test_fg_batched_matches_single_vector_path(S, nb, nt, C, dtype):
    """sfh_eval_fg_batched: C coefficient vectors in one pass == C calls of fg! (solvers.jl:20-38), and == the quad oracle."""
    M, x, data = make_flat_problem(nb, nt, seed=41, dtype=dtype)
    X = np.maximum(1e-3, x[:, None] * (1 + 0.2 * np.random.default_rng(7).standard_normal((nt, C))))
    ds = S.DeviceStack(M, data)
    nl, G = ds.eval_fg_batched(X)
    assert nl.shape == (C,) and G.shape == (nt, C)
    for c in sorted(set([0, C // 2, C - 1])):
        nl1, G1, _ = ds.eval_fg(X[:, c])
        assert nl[c] == pytest.approx(nl1, rel=1e-12)
        if dtype == np.float64:
            nlq, Gq, gs, _ = O.fg_quad(X[:, c], M, data)
            assert nl[c] == pytest.approx(nlq, rel=1e-12)
            assert np.all(np.abs(G[:, c] - Gq) <= 1e-10 * gs)
        else:
            nlq, Gq, gs = O.fg_quad_f32(X[:, c], M, data)
            assert nl[c] == pytest.approx(nlq, rel=1e-6) and np.all(np.abs(G[:, c] - Gq) <= 1e-6 * gs)
        assert np.all(np.abs(G[:, c] - G1) <= 1e-10 * np.maximum(np.abs(G1), 1e-3 * np.abs(G1).max()))
    nl_only, none = ds.eval_fg_batched(X, want_G=False)
    assert none is None and np.array_equal(nl_only, nl)
    # HMC adapter, all chains at once (hmc_sample.jl:24-37 per chain)
    hm = S.HMCModel(ds, None, data)
    LP, GR = hm.logdensity_and_gradient_batched(np.log(X))
    lp0, g0 = hm.logdensity_and_gradient(np.log(X[:, 0]))
    Mf, df = M.astype(np.float64), data.astype(np.float64)
    gscale0 = X[:, 0] * (np.abs(Mf).T @ np.abs(1.0 - df / np.maximum(Mf @ X[:, 0], np.finfo(np.float64).eps)))
    assert LP[0] == pytest.approx(lp0, rel=1e-12) and np.all(np.abs(GR[:, 0] - g0) <= 1e-10 * gscale0)
