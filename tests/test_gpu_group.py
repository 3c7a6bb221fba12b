"""Single-process multi-GPU groups (include/sfhcuda.h: sfh_group_*): ONE process, bin rows sharded over the visible GPUs, the
fused fg! / hierarchical fg! all-reduced by the one-shot NVLink exchange inside the finalize kernel.  Every assertion compares a
group with the same stack held whole on one GPU (and that with the oracle); runs with however many GPUs the box has (ndev = 1
exercises the same code without peers), the N > 1 cases need gpurun --gpus N."""
import ctypes as C

import numpy as np
import pytest

import oracle as O
from conftest import make_flat_problem, make_hier_problem

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def S():
    import sfh_b200
    assert sfh_b200.device_count() >= 1
    return sfh_b200


def ndevs(S):
    n = S.device_count()
    return sorted({1, min(2, n), n})


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_group_fg_matches_whole_and_oracle(S, dtype):
    nb, nt = 40013, 700
    M, x, data = make_flat_problem(nb, nt, seed=41, dtype=dtype)
    whole = S.DeviceStack(M, data)
    rt = 1e-12 if dtype is np.float64 else 1e-6
    want = O.fg_quad(x * 1.1, M, data)[:3] if dtype is np.float64 else O.fg_quad_f32(x * 1.1, M, data)
    for n in ndevs(S):
        g = S.DeviceStackGroup(M, data, ndev=n)
        infos = g.infos()
        assert len(infos) == n and all(i.fused == 1 for i in infos)
        assert infos[0].row_begin == 0 and infos[-1].row_end == nb
        assert all(infos[k].row_end == infos[k + 1].row_begin for k in range(n - 1))
        rng = np.random.default_rng(5)
        for it in range(25):                         # epochs, inbox parities, replayed graphs; gradient and logL-only calls
            xi = x * (1.0 + 0.2 * rng.random(nt)) if it else x * 1.1
            wg = it % 3 != 2
            a = whole.eval_fg(xi, want_G=wg); b = g.eval_fg(xi, want_G=wg)
            assert abs(a[0] - b[0]) <= rt * abs(a[0]), (n, it)
            if wg:
                assert np.allclose(a[1], b[1], rtol=1e-10 if dtype is np.float64 else 1e-6, atol=1e-10 * np.abs(a[1]).max()), (n, it)
            if it == 0:
                assert b[0] == pytest.approx(want[0], rel=rt)
                assert np.all(np.abs(b[1] - want[1]) <= (1e-10 if dtype is np.float64 else 1e-6) * want[2])
        with pytest.raises(S.SFHError):
            g.eval_logl_batched(np.ones((nt, 4)))    # NCCL-reduced paths are refused on a group, not silently shard-local
        with pytest.raises(S.SFHError):
            g.eval_fg(x, want_composite=True)
        g.close()


def test_group_hierarchical_and_native_drivers(S):
    p = make_hier_problem(nj=12, nk=10, nb=20011)
    mz, dp = S.PowerLawMZR(1.0, -2.0, 6.0), S.GaussianDispersion(0.2)
    xt = S.calculate_coeffs(mz, dp, p["R"], p["logAge"], p["MH"])
    d = np.random.default_rng(2).poisson(p["M"] @ xt).astype(np.float64)
    v = np.concatenate([p["R"], [1.0, -2.0, 0.2]]) * 1.07
    whole = S.DeviceStack(p["M"], d)
    Ga = np.empty(15)
    na = S.fg_(True, Ga, mz, dp, v, whole, d, None, p["logAge"], p["MH"])
    for n in ndevs(S):
        g = S.DeviceStackGroup(p["M"], d, ndev=n)
        Gb = np.empty(15)
        for _ in range(3):
            nb_ = S.fg_(True, Gb, mz, dp, v, g, None, None, p["logAge"], p["MH"])
            assert abs(na - nb_) <= 1e-12 * abs(na)
            assert np.allclose(Ga, Gb, rtol=1e-9, atol=1e-10 * np.abs(Ga).max())
        nl_only = S.fg_(True, None, mz, dp, v, g, None, None, p["logAge"], p["MH"])
        assert abs(nl_only - na) <= 1e-12 * abs(na)
        # the native L-BFGS-B loop runs unchanged on the group's primary context
        M2, x2, d2 = make_flat_problem(20000, 40, seed=3)
        a = S.fit_templates_lbfgsb(M2, d2, x0=np.ones(40), engine="native")
        g2 = S.DeviceStackGroup(M2, d2, ndev=n)
        b = S.fit_templates_lbfgsb(g2, d2, x0=np.ones(40), engine="native")
        assert abs(a[0] - b[0]) <= 1e-9 * abs(a[0]) and np.allclose(a[1], b[1], rtol=1e-5, atol=1e-8 * np.abs(a[1]).max())
        assert np.allclose(g2.column_sums(), M2.sum(axis=0), rtol=1e-12)
        g.close(); g2.close()


def test_group_argument_errors(S):
    M, x, data = make_flat_problem(4096, 33, seed=1)
    with pytest.raises((S.SFHError, ValueError)):
        S.DeviceStackGroup(M, data, ndev=S.device_count() + 1)
    with pytest.raises((S.SFHError, ValueError)):
        S.DeviceStackGroup(M, data, devices=[0, 0])
    with pytest.raises((S.SFHError, ValueError)):
        S.DeviceStackGroup(M[:64], data[:64], ndev=1)   # < 128 bins per GPU
    g = S.DeviceStackGroup(M, data, ndev=1)
    h = C.c_void_p(g.ctx().handle.value)
    assert S._lib.lib.sfh_ctx_destroy(h) != 0          # the primary context belongs to the group
    g.close()
