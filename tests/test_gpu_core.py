"""GPU parity tests of the core path (composite! / loglikelihood / grad-loglikelihood! / fg!) through the
C-ABI, against (i) the reference's own known-answer tests (fitting_core_test.jl), (ii) the CPU oracle
and its __float128 arbiter on seeded inputs, (iii) size-independent properties at BASELINE sizes.

Tolerances (BASELINE.json north_star): Float64 stacks 1e-12 relative on logL, 1e-10 per gradient
component -- measured against the backward-error scale sum_i |M_ij (1 - n_i/m_i)| that any reordering of
the sum is subject to (SURVEY.md section 7 "Tolerance definition"), and as plain relative error where the
component is not a near-total cancellation.  Float32-stored stacks with FP64 accumulation: 1e-6.
"""
import numpy as np
import pytest

import oracle as O
from conftest import make_flat_problem

pytestmark = pytest.mark.gpu

RTOL_LOGL = 1e-12
RTOL_GRAD = 1e-10
RTOL_F32 = 1e-6


@pytest.fixture(scope="module")
def S():
    import sfh_b200
    assert sfh_b200.device_count() >= 1
    return sfh_b200


def jl(rows, dt):
    return np.array(rows, dtype=dt)


def assert_grad_close(G, Gq, gscale, rtol=RTOL_GRAD):
    err = np.abs(G - Gq)
    assert np.all(err <= rtol * gscale + 1e-300), float(np.max(err / (gscale + 1e-300)))
    big = np.abs(Gq) > 1e-3 * gscale          # not a near-total cancellation: plain relative error
    if big.any():
        assert np.max(err[big] / np.abs(Gq[big])) <= rtol * 1e3


# ------------------------------------------------------------------ reference KATs through the C-ABI
@pytest.mark.parametrize("T,rtol", [(np.float32, 1e-3), (np.float64, 1e-7)])
def test_reference_kats(S, T, rtol):
    # composite!  fitting_core_test.jl:9-31
    A = jl([[0, 0, 0], [1, 1, 1], [0, 0, 0]], T); B = jl([[0, 0, 0], [0, 0, 0], [1, 1, 1]], T)
    coeffs = np.array([1, 2], dtype=T)
    models2 = S.stack_models([A, B])
    assert np.array_equal(models2[:, 0], [0, 1, 0, 0, 1, 0, 0, 1, 0]) and np.array_equal(models2[:, 1], [0, 0, 1, 0, 0, 1, 0, 0, 1])
    Cm = np.zeros((3, 3), dtype=T)
    assert S.composite_(Cm, coeffs, [A, B], data=np.zeros((3, 3))) is None
    assert np.array_equal(Cm, jl([[0, 0, 0], [1, 1, 1], [2, 2, 2]], T))
    C2 = np.zeros(9, dtype=T)
    S.composite_(C2, coeffs, models2, data=np.zeros(9))
    assert np.array_equal(C2, np.array([0, 1, 2, 0, 1, 2, 0, 1, 2], dtype=T))

    # loglikelihood  :32-70
    data = np.array([[1, 1, 1], [2, 2, 2], [2, 2, 2]], dtype=np.int64)
    A = jl([[1, 1, 1], [0, 0, 0], [0, 0, 0]], T); B = jl([[0, 0, 0], [1, 1, 1], [1.5, 1.5, 1.5]], T)
    r = S.loglikelihood(coeffs, [A, B], data)
    assert isinstance(r, T) and r == pytest.approx(-0.5672093513510137, rel=rtol)
    ds = S.DeviceStack([A, B], data)
    Cmat = jl([[1, 1, 1], [2, 2, 2], [3, 3, 3]], T)
    r2 = S.loglikelihood(Cmat, ds)
    assert isinstance(r2, T) and r2 == pytest.approx(-0.5672093513510137, rel=rtol)
    dz = np.array([[0, 0, 0], [2, 2, 2], [2, 2, 2]], dtype=np.int64)
    dsz = S.DeviceStack([A, B], dz)
    assert S.loglikelihood(jl([[1.5, 1.5, 1.5], [3, 3, 3], [3, 3, 3]], T), dsz) == pytest.approx(-5.6344187027020260, rel=rtol)

    # grad-loglikelihood / grad-loglikelihood!  :71-162
    models = [jl([[1, 1, 1], [0, 0, 0], [0, 0, 0]], T), jl([[0, 0, 0], [1, 1, 1], [0, 0, 0]], T), jl([[0, 0, 0], [0, 0, 0], [1, 1, 1]], T)]
    c3 = np.array([1.5, 3, 3], dtype=T)
    # grad-loglikelihood(model, composite, data): ONE template, scalar result of the promoted type  fitting_core_test.jl:77-97
    model1 = jl([[0, 0, 0], [0, 0, 0], [1, 1, 1]], T)
    C1 = jl([[1, 1, 1], [2, 2, 2], [3, 3, 3]], T)
    d1 = np.asfortranarray(np.array([[1, 1, 1], [2, 2, 2], [2, 2, 2]], dtype=np.int64))
    r = S.grad_loglikelihood(model1, C1, d1)
    assert r == pytest.approx(-1, rel=rtol) and isinstance(r, T)                       # :79-80
    r = S.grad_loglikelihood(model1.reshape(-1, order="F"), C1.reshape(-1, order="F"), d1.reshape(-1, order="F"))
    assert r == pytest.approx(-1, rel=rtol) and isinstance(r, T)                       # :85-86 (flattened inputs)
    mz_ = jl([[1, 1, 1], [0, 0, 0], [0, 0, 0]], T)
    Cz_ = jl([[1.5, 1.5, 1.5], [3, 3, 3], [3, 3, 3]], T)
    dz_ = np.asfortranarray(np.array([[0, 0, 0], [2, 2, 2], [2, 2, 2]], dtype=np.int64))
    r = S.grad_loglikelihood(mz_, Cz_, dz_)
    assert r == pytest.approx(-3, rel=rtol) and isinstance(r, T)                       # :92-97 zero-data bins contribute -c_ij
    with pytest.raises(ValueError):
        S.grad_loglikelihood(model1, C1, d1[:2])                                       # fitting_base.jl:147
    g = S.grad_loglikelihood(c3, S.DeviceStack(models, data), data)
    assert g.dtype == T and g.shape == (3,) and np.allclose(g, [-1, -1, -1], rtol=rtol)
    grad = np.empty(3, dtype=T)
    Cc = sum(c * m for c, m in zip(c3, models)).astype(T)
    S.grad_loglikelihood_(grad, Cc, S.DeviceStack(models, data), data)
    assert np.allclose(grad, [-1, -1, -1], rtol=rtol)
    # side effect: composite now holds 1 - n/m  (fitting_base.jl:219)
    assert np.allclose(Cc, 1 - data / sum(c * m for c, m in zip(c3, models)), rtol=rtol)
    grad3 = np.empty(3, dtype=T)
    S.grad_loglikelihood_(grad3, sum(c * m for c, m in zip(c3, models)).astype(T), S.DeviceStack(models, dz), dz)
    assert np.allclose(grad3, [-3, -1, -1], rtol=rtol)

    # fg!  :163-195
    G = np.empty(3, dtype=T); Cs = np.empty((3, 3), dtype=T)
    res = S.fg_(True, G, c3, models, data, Cs)
    assert isinstance(res, T) and -res == pytest.approx(-1.4180233783775342, rel=rtol)
    assert np.allclose(-G, [-1, -1, -1], rtol=rtol)
    res2 = S.fg_(True, None, c3, S.stack_models(models), data.reshape(-1, order="F"), np.empty(9, dtype=T))
    assert -res2 == pytest.approx(-1.4180233783775342, rel=rtol)
    G2 = np.empty(3, dtype=T)
    assert S.fg_(None, G2, c3, models, data) is None and np.allclose(G2, G)


def test_argchecks(S):
    M, x, data = make_flat_problem(64, 5)
    ds = S.DeviceStack(M, data)
    with pytest.raises(ValueError):
        ds.eval_fg(np.ones(4))                                  # solvers.jl:10
    with pytest.raises(ValueError):
        S.DeviceStack(M, data[:-1])                             # solvers.jl:11
    with pytest.raises(ValueError):
        S.fg_(True, np.empty(4), x, ds, data)
    with pytest.raises(ValueError):
        S.composite_(np.empty(63), x, ds)                       # fitting_base.jl:58


# ------------------------------------------------------------------ seeded parity vs oracle + quad arbiter
SHAPES = [(1, 1), (7, 3), (64, 16), (100, 100), (999, 37), (9801, 142), (4096, 600), (10000, 100), (2500, 2400)]


@pytest.mark.parametrize("nb,nt", SHAPES)
def test_fg_parity_f64(S, nb, nt):
    M, x, data = make_flat_problem(nb, nt, seed=58392 + nb)
    ds = S.DeviceStack(M, data)
    nl, G, resid = ds.eval_fg(x * 1.3, want_composite=True)
    nlq, Gq, gs, compq = O.fg_quad(x * 1.3, M, data)
    nlo, Go, resid_o = O.fg(x * 1.3, M, data)
    assert nl == pytest.approx(nlq, rel=RTOL_LOGL)
    assert nlo == pytest.approx(nlq, rel=RTOL_LOGL)             # the double oracle obeys the same bar
    assert_grad_close(G, Gq, gs)
    assert_grad_close(Go, Gq, gs)
    assert np.allclose(resid, resid_o, rtol=1e-9, atol=1e-12)   # reference side effect on `composite`
    # logL-only call (G === nothing) and composite!
    nl2, G2, comp = ds.eval_fg(x * 1.3, want_G=False, want_composite=True)
    assert G2 is None and nl2 == nl
    assert np.allclose(comp, compq, rtol=1e-13, atol=0)
    info = ds.info()
    assert info.cc_major == 10 and info.fused == 1, "fused sm_100a kernel must be the path that runs"


@pytest.mark.parametrize("nw", [8, 16])
@pytest.mark.parametrize("tile,cluster", [(8, 2), (16, 1), (16, 2), (32, 2), (32, 4), (64, 4), (64, 8), (32, 8), (16, 8), (64, 16)])
def test_fused_configs_agree(S, tile, cluster, nw):
    """The cluster-tile kernel (sfh_opts.variant = 1) with 8 / 16 consumer warps x tile x cluster."""
    nb, nt = 3001, 517
    M, x, data = make_flat_problem(nb, nt, seed=11)
    nlq, Gq, gs, _ = O.fg_quad(x, M, data)
    ds = S.DeviceStack(M, data, tile_bins=tile, cluster=cluster, consumer_warps=nw, variant=1)
    i = ds.info()
    if not i.fused:
        pytest.skip("this (tile, cluster, warps) combination cannot hold T in its per-lane registers")
    assert i.tile_bins == tile and i.cluster == cluster and i.consumer_warps == nw and i.variant == 1
    nl, G, _ = ds.eval_fg(x)
    assert nl == pytest.approx(nlq, rel=RTOL_LOGL)
    assert_grad_close(G, Gq, gs)


@pytest.mark.parametrize("nw,tile,cluster", [(8, 8, 2), (8, 8, 4), (16, 8, 1), (8, 16, 4), (16, 32, 2), (8, 64, 8), (16, 128, 4)])
def test_fused_configs_f32(S, nw, tile, cluster):
    """Float32-stored stacks through every tile width (8 ... 128 bins) of the cluster-tile kernel."""
    nb, nt = 2777, 701
    M, x, data = make_flat_problem(nb, nt, seed=13, dtype=np.float32)
    nlq, Gq, gs = O.fg_quad_f32(x, M, data)
    ds = S.DeviceStack(M, data, tile_bins=tile, cluster=cluster, consumer_warps=nw, variant=1)
    i = ds.info()
    if not i.fused:
        pytest.skip("combination cannot hold T")
    assert i.tile_bins == tile and i.cluster == cluster and i.panel_layout == 1
    nl, G, _ = ds.eval_fg(x)
    assert nl == pytest.approx(nlq, rel=RTOL_F32)
    assert_grad_close(G, Gq, gs, rtol=RTOL_F32)
    Md, dd = ds.download()                       # the panel re-tiling is invisible to the host
    assert np.array_equal(Md, M) and np.array_equal(dd, data.astype(np.float64))


# the warp-specialised stream kernel (sfh_opts.variant = 4): every lanes-per-row width x cluster size, F64 and F32, with
# shapes whose last chunk / last stage / last tile are partial and whose slices run past the last template
@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("lpr,cluster", [(1, 1), (1, 2), (1, 4), (2, 1), (2, 2), (4, 1), (4, 8), (8, 1), (8, 2), (16, 1), (16, 4), (32, 1), (32, 2)])
@pytest.mark.parametrize("nb,nt", [(3001, 517), (777, 2400), (4099, 129), (50, 5000)])
def test_stream_kernel_configs(S, dtype, lpr, cluster, nb, nt):
    vec = 16 // np.dtype(dtype).itemsize
    M, x, data = make_flat_problem(nb, nt, seed=17 + lpr, dtype=dtype)
    ds = S.DeviceStack(M, data, tile_bins=vec * lpr, cluster=cluster, variant=4)
    i = ds.info()
    if not i.fused:
        pytest.skip("this (lanes per row, cluster) combination cannot hold T in its per-lane registers")
    assert i.variant == 4 and i.tile_bins == vec * lpr and i.cluster == cluster and i.panel_layout == 1
    nl, G, resid = ds.eval_fg(x * 1.1, want_composite=True)
    if dtype is np.float64:
        nlq, Gq, gs, compq = O.fg_quad(x * 1.1, M, data)
        rt_l, rt_g = RTOL_LOGL, 1e-10
    else:
        nlq, Gq, gs = O.fg_quad_f32(x * 1.1, M, data)
        rt_l, rt_g = RTOL_F32, RTOL_F32
    assert nl == pytest.approx(nlq, rel=rt_l)
    assert_grad_close(G, Gq, gs, rtol=rt_g)
    nl2, G2, comp = ds.eval_fg(x * 1.1, want_G=False, want_composite=True)      # logL-only instantiation: same logL, bit for bit
    assert G2 is None and nl2 == nl
    if dtype is np.float64:
        assert np.allclose(comp, compq, rtol=1e-13, atol=0)
        assert np.allclose(resid, 1.0 - data / np.maximum(comp, np.finfo(np.float64).eps), rtol=0, atol=0)
    nl3, G3, _ = ds.eval_fg(x * 1.1)                                            # bitwise run-to-run determinism
    assert nl3 == nl and np.array_equal(G3, G)
    Md, dd = ds.download()
    assert np.array_equal(Md, M) and np.array_equal(dd, data.astype(np.float64))


def test_stream_kernel_f32_unpack_paths(S):
    """Float32 stacks take the conversion-free unpack only when every element is finite and non-negative (checked at creation);
    otherwise the converting instantiation.  Both must agree with the oracle -- including float denormals (exact in both),
    a negative template value and NaN propagation (never clamp NaN away)."""
    nb, nt = 3000, 777
    M, x, data = make_flat_problem(nb, nt, seed=31, dtype=np.float32)
    M[5, 7] = np.float32(1e-42)                    # float denormal
    M[11, 0] = np.float32(0.0)
    x = x * 1.07
    want = O.fg_quad_f32(x, M, data)
    ds = S.DeviceStack(M, data, variant=4)
    assert ds.info().variant == 4
    nl, G, _ = ds.eval_fg(x)
    assert nl == pytest.approx(want[0], rel=RTOL_F32)
    assert_grad_close(G, want[1], want[2], rtol=RTOL_F32)
    # one negative entry switches the whole stack over to the converting instantiation
    M2 = M.copy(); M2[100, 3] = -M2[100, 3]
    w2 = O.fg_quad_f32(x, M2, data)
    ds2 = S.DeviceStack(M2, data, variant=4)
    nl2, G2, _ = ds2.eval_fg(x)
    assert nl2 == pytest.approx(w2[0], rel=RTOL_F32)
    assert_grad_close(G2, w2[1], w2[2], rtol=RTOL_F32)
    # NaN in a template: logL and every gradient component touching that bin are NaN, nothing is silently clamped
    M3 = M.copy(); M3[17, 5] = np.nan
    ds3 = S.DeviceStack(M3, data, variant=4)
    nl3, G3, _ = ds3.eval_fg(x)
    assert np.isnan(nl3) and np.all(np.isnan(G3))
    M4 = M.copy(); M4[17, 5] = np.inf
    ds4 = S.DeviceStack(M4, data, variant=4)
    nl4, G4, _ = ds4.eval_fg(x)
    assert not np.isfinite(nl4)


def test_unfused_two_pass_agrees(S):
    M, x, data = make_flat_problem(5000, 301, seed=5)
    nlq, Gq, gs, _ = O.fg_quad(x, M, data)
    ds = S.DeviceStack(M, data, force_unfused=True)
    assert ds.info().fused == 0
    nl, G, _ = ds.eval_fg(x)
    assert nl == pytest.approx(nlq, rel=RTOL_LOGL)
    assert_grad_close(G, Gq, gs)


def test_bitwise_determinism(S):
    M, x, data = make_flat_problem(20000, 500, seed=3)
    ds = S.DeviceStack(M, data)
    a = ds.eval_fg(x)
    for _ in range(5):
        b = ds.eval_fg(x)
        assert a[0] == b[0] and np.array_equal(a[1], b[1])
    ds2 = S.DeviceStack(M, data)                                # a second upload, another context
    c = ds2.eval_fg(x)
    assert a[0] == c[0] and np.array_equal(a[1], c[1])


@pytest.mark.parametrize("nb,nt", [(700, 33), (10000, 100), (20000, 1203)])
def test_packet_completion_never_returns_a_stale_result(S, nb, nt):
    """The host-synchronous call returns when the finalize kernel's self-validating packets {lo, epoch, hi, epoch} have
    arrived in pinned memory -- there is no stream synchronisation behind it (csrc/sfh_api.cu wait_packets).  Thousands of
    back-to-back calls with a DIFFERENT coefficient vector each time, with and without the gradient: every answer must be
    the one of its own inputs, bit for bit (a stale or torn packet would surface as the previous call's value)."""
    M, x, data = make_flat_problem(nb, nt, seed=11)
    ds = S.DeviceStack(M, data)
    rng = np.random.default_rng(5)
    xs = [x * (1 + 0.2 * rng.random(nt)) for _ in range(5)]
    want = [ds.eval_fg(xk) for xk in xs]
    for k, xk in enumerate(xs):                                   # the references themselves: against the oracle
        nlq, Gq, gs, _ = O.fg_quad(xk, M, data)
        assert want[k][0] == pytest.approx(nlq, rel=RTOL_LOGL)
        assert_grad_close(want[k][1], Gq, gs)
    for it in range(3000):
        k = int(rng.integers(5))
        if it % 7 == 3:
            nl = ds.eval_fg(xs[k], want_G=False)[0]
            assert nl == want[k][0], (it, k)
        else:
            nl, G = ds.eval_fg(xs[k])[:2]
            assert nl == want[k][0] and np.array_equal(G, want[k][1]), (it, k)


def test_l2_resident_head_changes_nothing_but_time(S, monkeypatch):
    """A stack larger than L2 is streamed evict_first except the head of every CTA's tile sequence (evict_last: it stays in L2
    between evaluations, sfh_info.l2_resident_mb).  Cache policies must be invisible in the results: the same stack evaluated
    with the budget at its default, switched off and oversized gives bit-identical answers (the switch is read per call)."""
    rng = np.random.default_rng(3)
    nb, nt = 21000, 1000                                         # 168 MB of Float64 > 126 MB of L2
    x = 100 * rng.random(nt)
    ds = S.DeviceStack.synthetic(nb, nt, np.float64, seed=77, scale=1.0, x_true=x)
    info = ds.info()
    assert info.fused == 1 and info.variant == 4
    assert 0 < info.l2_resident_mb <= 0.08 * info.stack_bytes / 2**20 + 1
    x1 = x * (1 + 0.05 * rng.standard_normal(nt))
    ref = ds.eval_fg(x1)
    for mb in ("0", "200", "16"):
        monkeypatch.setenv("SFH_L2_KEEP_MB", mb)
        for _ in range(3):
            got = ds.eval_fg(x1)
            assert got[0] == ref[0] and np.array_equal(got[1], ref[1]), mb
    monkeypatch.setenv("SFH_L2_KEEP_MB", "0")
    assert ds.info().l2_resident_mb == 0
    small = S.DeviceStack.synthetic(2000, 100, np.float64, seed=1, scale=1.0, x_true=x[:100])   # fits L2: nothing is streamed
    assert small.info().l2_resident_mb == 0


@pytest.mark.parametrize("nb,nt", [(100, 100), (9801, 142), (11250, 2000), (5000, 2400)])
def test_fg_parity_f32_storage(S, nb, nt):
    """Float32-stored templates, FP64 accumulation, vs exact arithmetic on the same stored values (1e-6)."""
    M, x, data = make_flat_problem(nb, nt, seed=77, dtype=np.float32)
    eps32 = float(np.finfo(np.float32).eps)
    ds = S.DeviceStack(M, data)
    assert ds.info().dtype == 0 and ds.info().clamp_eps == eps32
    nl, G, _ = ds.eval_fg(x)
    nlq, Gq, gs = O.fg_quad_f32(x, M, data)
    assert nl == pytest.approx(nlq, rel=RTOL_F32)
    assert_grad_close(G, Gq, gs, rtol=RTOL_F32)
    assert isinstance(S.fg_(True, None, x, ds, data), np.float32)   # returned scalar typed like the stack


def test_f32_special_values(S):
    """F32 storage widens exactly: zeros, negative values, subnormals and FLT_MAX/FLT_MIN come out bit for bit, the fused
    and two-pass paths agree, and Inf/NaN templates propagate like the reference's arithmetic."""
    nb, nt = 3000, 300
    M, x, data = make_flat_problem(nb, nt, seed=5, dtype=np.float32)
    rng = np.random.default_rng(8)
    M[rng.random(M.shape) < 0.3] = 0.0                                   # Hess templates are mostly empty
    M[rng.random(M.shape) < 0.05] *= -1.0                                # sign bit
    sub = rng.random(M.shape) < 0.02
    M[sub] = (rng.integers(1, 1 << 23, size=int(sub.sum())).astype(np.uint32)).view(np.float32)   # subnormals
    M[0, 0] = np.float32(np.finfo(np.float32).max); M[1, 1] = np.float32(np.finfo(np.float32).tiny)
    ds = S.DeviceStack(M, data)
    assert ds.info().fused == 1
    x = x * 1e-3
    x[0] = 1e-30                                                          # keeps FLT_MAX * x finite
    nl, G, comp = ds.eval_fg(x, want_composite=True)
    ds2 = S.DeviceStack(M, data, force_unfused=True)
    nl2, G2, _ = ds2.eval_fg(x)
    nlq, Gq, gs = O.fg_quad_f32(x, M, data)
    assert nl == pytest.approx(nlq, rel=RTOL_F32) and nl == pytest.approx(nl2, rel=1e-12)
    assert_grad_close(G, Gq, gs, rtol=RTOL_F32)
    assert np.all(np.abs(G - G2) <= 1e-11 * np.maximum(np.abs(G2), 1e-3 * np.abs(G2).max()))
    # exact F32 -> F64 widening: a one-hot coefficient vector returns the stored column bit for bit (composite = M[:, j] * 1)
    for j in (0, 1, 17):
        e = np.zeros(nt); e[j] = 1.0
        c = np.empty(nb)
        S.composite_(c, e, ds)
        assert np.array_equal(c, M[:, j].astype(np.float64))
    # Inf / NaN templates propagate exactly like the reference's arithmetic
    for bad in (np.inf, np.nan):
        Mb = M.copy(); Mb[7, 3] = bad
        db = S.DeviceStack(Mb, data)
        nlb, Gb, _ = db.eval_fg(np.abs(x) + 1.0)
        nlo, Go, _ = O.fg(np.abs(x) + 1.0, Mb.astype(np.float64), data.astype(np.float64))
        assert (np.isnan(nlb) and np.isnan(nlo)) or nlb == pytest.approx(nlo, rel=1e-5)
        assert np.array_equal(np.isnan(Gb), np.isnan(Go))


# ------------------------------------------------------------------ semantics the reference fixes (SURVEY 8a checklist)
def test_edge_semantics(S):
    eps = np.finfo(np.float64).eps
    # zero-sum guard: m == n everywhere -> every term 0 -> logL = -Inf -> fg! returns +Inf (fitting_base.jl:95)
    M = np.asfortranarray(np.eye(6)); data = np.ones(6)
    ds = S.DeviceStack(M, data)
    nl, G, _ = ds.eval_fg(np.ones(6))
    assert nl == np.inf and np.array_equal(G, np.zeros(6))
    # clamp: composite <= 0 -> eps (fitting_base.jl:90,277): zero and negative coefficients
    nl, G, _ = ds.eval_fg(np.zeros(6))
    nlo, Go, _ = O.fg(np.zeros(6), M, data)
    assert nl == pytest.approx(nlo, rel=1e-14) and np.allclose(G, Go, rtol=1e-14)
    assert G[0] == pytest.approx(1 - 1 / eps)
    nl, G, _ = ds.eval_fg(-np.ones(6))
    nlo, Go, _ = O.fg(-np.ones(6), M, data)
    assert nl == pytest.approx(nlo, rel=1e-14) and np.allclose(G, Go, rtol=1e-14)
    # zero-count bins contribute -m_i and -c_ij; fractional "data" is legal (mzr_test.jl:66)
    M2, x2, _ = make_flat_problem(300, 9, seed=2)
    d2 = M2 @ x2
    d2[::3] = 0.0
    ds2 = S.DeviceStack(M2, d2)
    nl, G, _ = ds2.eval_fg(x2 * 0.9)
    nlq, Gq, gs, _ = O.fg_quad(x2 * 0.9, M2, d2)
    assert nl == pytest.approx(nlq, rel=RTOL_LOGL)
    assert_grad_close(G, Gq, gs)
    # NaN propagates
    xn = x2.copy(); xn[4] = np.nan
    nl, G, _ = ds2.eval_fg(xn)
    assert np.isnan(nl) and np.isnan(G).all()
    # Int64 and Float32 data on a Float64 stack (fitting_core_test.jl:38)
    di = np.arange(300, dtype=np.int64) % 7
    a = S.DeviceStack(M2, di).eval_fg(x2); b = S.DeviceStack(M2, di.astype(np.float64)).eval_fg(x2)
    c = S.DeviceStack(M2, di.astype(np.float32)).eval_fg(x2)
    assert a[0] == b[0] == c[0] and np.array_equal(a[1], b[1]) and np.array_equal(a[1], c[1])
    # set_data rebinding (MCMCModelDistance-style callers)
    s3 = S.DeviceStack(M2, di)
    s3.set_data(d2)
    assert s3.eval_fg(x2 * 0.9)[0] == pytest.approx(nlq, rel=RTOL_LOGL)


def test_stacks_with_different_configs_coexist(S):
    """cudaFuncAttributeMaxDynamicSharedMemorySize is per kernel function: a later, smaller stack must not
    shrink it under an earlier, larger one that shares the instantiation."""
    M, x, data = make_flat_problem(5003, 120, seed=12)
    big = S.DeviceStack(M, data)
    a = big.eval_fg(x)
    smalls = [S.DeviceStack(M[:n], data[:n]) for n in (2502, 700, 64)]
    for s_ in smalls:
        s_.eval_fg(x)
    b = big.eval_fg(x)
    assert a[0] == b[0] and np.array_equal(a[1], b[1])


def test_row_shards_sum_to_whole(S):
    """Bin-row sharding: logL and G are sums over shards."""
    M, x, data = make_flat_problem(7777, 260, seed=8)
    whole = S.DeviceStack(M, data)
    nl, G, _ = whole.eval_fg(x)
    cuts = [0, 1000, 1001, 5000, 7777]
    parts = [S.DeviceStack(M, data, rows=(a, b)).eval_fg(x) for a, b in zip(cuts[:-1], cuts[1:])]
    assert sum(p[0] for p in parts) == pytest.approx(nl, rel=1e-13)
    assert np.allclose(sum(p[1] for p in parts), G, rtol=1e-11, atol=1e-9)
    Ms, ds_ = S.DeviceStack(M, data, rows=(1000, 1001)).download()
    assert np.array_equal(Ms, M[1000:1001]) and np.array_equal(ds_, data[1000:1001])


def test_concurrent_callers_share_a_stack(S):
    """Many host threads, one immutable stack, one context each (hmc_sample.jl:127,135)."""
    import threading
    M, x, data = make_flat_problem(6000, 200, seed=4)
    ds = S.DeviceStack(M, data)
    ref = [ds.eval_fg(x * (1 + 0.01 * k)) for k in range(8)]
    out = [None] * 8

    def work(k):
        for _ in range(10):
            out[k] = ds.eval_fg(x * (1 + 0.01 * k))

    th = [threading.Thread(target=work, args=(k,)) for k in range(8)]
    [t.start() for t in th]; [t.join() for t in th]
    for k in range(8):
        assert out[k][0] == ref[k][0] and np.array_equal(out[k][1], ref[k][1])


# ------------------------------------------------------------------ BASELINE sizes: full oracle comparison + size-independent properties
@pytest.mark.parametrize("nb,nt,dtype", [(60000, 2400, np.float64), (40000, 500, np.float64), (200000, 1000, np.float32)])
def test_full_size_properties(S, nb, nt, dtype):
    rng = np.random.default_rng(1)
    x = 100 * rng.random(nt)
    ds = S.DeviceStack.synthetic(nb, nt, dtype, seed=94823, scale=1.0, x_true=x)
    assert ds.info().fused == 1
    x1 = x * (1 + 0.05 * rng.standard_normal(nt))
    nl, G, resid = ds.eval_fg(x1, want_composite=True)
    _, _, comp = ds.eval_fg(x1, want_G=False, want_composite=True)
    Md, data = ds.download()
    # (1) checksum of checksums:  sum_j x_j G_j == sum_i m_i (1 - n_i/m_i) == sum_i (m_i - n_i)
    assert np.dot(x1, G) == pytest.approx(np.sum(comp - data), rel=1e-9)
    # (2) residual identity and logL recomputed from the returned composite in numpy float64
    assert np.allclose(resid, 1 - data / comp, rtol=1e-12)
    term = np.where(data > 0, data - comp - data * np.log(np.where(data > 0, data, 1) / comp), -comp)
    assert nl == pytest.approx(-term.sum(), rel=1e-11)
    # (3) linearity of composite!
    _, _, comp2 = ds.eval_fg(2.5 * x1, want_G=False, want_composite=True)
    assert np.allclose(comp2, 2.5 * comp, rtol=1e-13)
    # (4) the FULL result against the oracle on the downloaded stack: logL and every one of the nt gradient components from the
    # double-precision restatement of the reference's two-pass fg! (O.fg; Float32 stacks: the __float128 arbiter on the
    # Float32-stored values, since the product accumulates those in FP64), every composite row from a numpy gemv.
    if dtype is np.float64:
        nlo, Go, _ = O.fg(x1, Md, data)
        gscale = np.abs(Md).T @ np.abs(1.0 - data / np.maximum(Md @ x1, np.finfo(np.float64).eps))
        assert nl == pytest.approx(nlo, rel=RTOL_LOGL)
        assert_grad_close(G, Go, gscale)
        assert np.allclose(comp, Md @ x1, rtol=1e-13, atol=0)
    else:
        nlq, Gq, gs = O.fg_quad_f32(x1, Md, data.astype(np.float32))
        assert nl == pytest.approx(nlq, rel=RTOL_F32)
        assert_grad_close(G, Gq, gs, rtol=RTOL_F32)
        assert np.allclose(comp, Md.astype(np.float64) @ x1, rtol=1e-12, atol=0)
    # (5) shards of the SAME synthetic matrix reproduce the whole (counter-based generator)
    h = nb // 2 + 7
    a = S.DeviceStack.synthetic(nb, nt, dtype, seed=94823, scale=1.0, x_true=x, rows=(0, h)).eval_fg(x1)
    b = S.DeviceStack.synthetic(nb, nt, dtype, seed=94823, scale=1.0, x_true=x, rows=(h, nb)).eval_fg(x1)
    assert a[0] + b[0] == pytest.approx(nl, rel=1e-12)
    assert np.allclose(a[1] + b[1], G, rtol=1e-9, atol=1e-6)


def test_very_wide_stack_falls_back_to_two_pass(S):
    """More templates than the fused tiling can hold (T > 8 CTAs x 20 chunks x 256 rows): the library must switch to
    the two-pass kernels by itself and still match the oracle."""
    nb, nt = 48, 50000
    M, x, data = make_flat_problem(nb, nt, seed=21, scale=0.01)
    ds = S.DeviceStack(M, data)
    assert ds.info().fused == 0
    nl, G, _ = ds.eval_fg(x)
    nlq, Gq, gs, _ = O.fg_quad(x, M, data)
    assert nl == pytest.approx(nlq, rel=RTOL_LOGL)
    assert_grad_close(G, Gq, gs)
    # a stack near the widest the fused path still takes (8-CTA clusters of the stream kernel, 16 of the cluster-tile kernel)
    M2, x2, d2 = make_flat_problem(40, 30000, seed=22, scale=0.01)
    ds2 = S.DeviceStack(M2, d2)
    assert ds2.info().fused == 1 and ds2.info().cluster in (8, 16)
    nl2, G2, _ = ds2.eval_fg(x2)
    nlq2, Gq2, gs2, _ = O.fg_quad(x2, M2, d2)
    assert nl2 == pytest.approx(nlq2, rel=RTOL_LOGL)
    assert_grad_close(G2, Gq2, gs2)


def test_empty_stacks(S):
    """Empty inputs: zero bins / zero templates (SURVEY section 8c edge cases).  An empty sum is exactly 0, which the
    reference maps to -typemax (fitting_base.jl:95), i.e. fg! returns +Inf."""
    ds = S.DeviceStack(np.zeros((0, 3)), np.zeros(0))
    nl, G, _ = ds.eval_fg(np.ones(3))
    assert nl == np.inf and np.array_equal(G, np.zeros(3))
    ds = S.DeviceStack(np.zeros((5, 0)), np.arange(5.0))
    nl, G, _ = ds.eval_fg(np.zeros(0))
    assert G.shape == (0,)
    assert nl == np.inf        # no templates: nothing is evaluated, the empty sum rule applies
    assert S.MCMCModel(S.DeviceStack(np.ones((4, 2)), np.ones(4)), None).batch(np.zeros((2, 0))).shape == (0,)
