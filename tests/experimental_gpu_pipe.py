"""GPU tests of the two-tiles-in-flight fused kernel (csrc/sfh_fused_pipe.cuh, sfh_opts.variant = 3).

NOT collected by the default `pytest tests` run (file name): the variant was written after round 1's GPU budget was spent and has
not run on a device yet.  First call of round 2:  python -m pytest tests/experimental_gpu_pipe.py -m gpu -q
Same bar as tests/test_gpu_core.py: logL 1e-12, gradient 1e-10 of its backward-error scale against the __float128 arbiter;
bitwise run-to-run determinism; F32 storage 1e-6; and agreement with the default variant.
"""
import numpy as np
import pytest

import oracle as O
from conftest import make_flat_problem

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def S():
    import sfh_b200
    assert sfh_b200.device_count() >= 1
    return sfh_b200


SHAPES = [(64, 16), (999, 37), (9801, 142), (4096, 600), (2500, 2400), (20000, 1000)]


@pytest.mark.parametrize("nb,nt", SHAPES)
@pytest.mark.parametrize("nw", [8, 16])
def test_pipe_matches_quad_oracle(S, nb, nt, nw):
    M, x, data = make_flat_problem(nb, nt)
    try:
        ds = S.DeviceStack(M, data, variant=3, consumer_warps=nw)
    except (S.SFHError, ValueError):
        pytest.skip("no pipelined tiling for this shape")
    i = ds.info()
    if not (i.fused and i.pipelined):
        pytest.skip("fell back to another path")
    assert i.ring_slots >= 2 * i.chunks_per_tile
    f, G, comp = ds.eval_fg(x, want_composite=True)
    fq, Gq, gscale, _ = O.fg_quad(x, M, data)
    assert abs(f - fq) <= 1e-12 * abs(fq)
    assert np.all(np.abs(G - Gq) <= 1e-10 * gscale + 1e-300)
    f2, G2, _ = ds.eval_fg(x)
    assert f2 == f and np.array_equal(G, G2)                      # bitwise determinism
    ref = S.DeviceStack(M, data)                                   # the default variant
    fr, Gr, compr = ref.eval_fg(x, want_composite=True)
    assert abs(f - fr) <= 1e-13 * abs(fr) and np.all(np.abs(G - Gr) <= 1e-12 * gscale + 1e-300)
    np.testing.assert_allclose(comp, compr, rtol=1e-13, atol=1e-300)
    fl, none, _ = ds.eval_fg(x, want_G=False)                      # logL-only evaluations use the standard kernel on the same tiling
    assert none is None and abs(fl - fq) <= 1e-12 * abs(fq)


@pytest.mark.parametrize("bt,c", [(8, 2), (8, 4), (16, 4), (16, 8), (32, 8)])
def test_pipe_forced_tilings(S, bt, c):
    M, x, data = make_flat_problem(6000, 1200)
    try:
        ds = S.DeviceStack(M, data, variant=3, tile_bins=bt, cluster=c)
    except (S.SFHError, ValueError):
        pytest.skip("tiling not available")
    if not ds.info().pipelined:
        pytest.skip("tiling not available")
    f, G, _ = ds.eval_fg(x)
    fq, Gq, gscale, _ = O.fg_quad(x, M, data)
    assert abs(f - fq) <= 1e-12 * abs(fq) and np.all(np.abs(G - Gq) <= 1e-10 * gscale + 1e-300)


def test_pipe_f32_storage_and_many_tiles_per_cluster(S):
    M, x, data = make_flat_problem(200000, 500, dtype=np.float32)
    ds = S.DeviceStack(M, data, variant=3)
    if not ds.info().pipelined:
        pytest.skip("no pipelined tiling")
    f, G, _ = ds.eval_fg(x)
    fo, Go, _ = O.fg(x, M.astype(np.float64), data.astype(np.float64))
    assert abs(f - fo) <= 1e-6 * abs(fo)
    assert np.all(np.abs(G - Go) <= 1e-6 * (np.abs(M.astype(np.float64)).T @ np.abs(1 - data / np.maximum(M.astype(np.float64) @ x, 1e-300))) + 1e-300)
    # hierarchical path on top of it
    f2, G2, _ = ds.eval_fg(x * 1.01)
    assert np.isfinite(f2) and np.all(np.isfinite(G2)) and f2 != f


# ---- the BFGS inverse Hessian resident on the device (sfh_bfgs_opts.device_hessian; also written after the GPU budget was spent) ----
@pytest.mark.parametrize("nb,nt", [(10000, 20), (20000, 600)])
def test_device_hessian_bfgs_matches_host_hessian(S, nb, nt):
    M, x, data = make_flat_problem(nb, nt)
    a = S.fit_templates(M, data, x0=np.ones(nt), engine="native")
    b = S.fit_templates(M, data, x0=np.ones(nt), engine="native", device_hessian=True)
    for k in ("map", "mle"):
        assert b[k].result.success == a[k].result.success
        assert np.linalg.norm(a[k].mu - b[k].mu) <= 1e-6 * np.linalg.norm(a[k].mu)
        H = b[k].invH
        assert np.allclose(H, H.T, atol=1e-10 * np.abs(H).max()) and np.all(np.diag(H) > 0)
        assert np.allclose(np.sqrt(np.diag(H)), np.sqrt(np.diag(a[k].invH)), rtol=0.3)
