"""Template construction on the device (sfh_stack_create_from_points) against the CPU oracle's bin_cmd_smooth
(src/StarFormationHistories.jl:574-621) on the same seeded point lists, through the C-ABI."""
import numpy as np
import pytest

import oracle as O
from conftest import make_flat_problem

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def S():
    import sfh_b200
    return sfh_b200


def _points(rng, n, xe, ye, spread=0.15):
    x = rng.uniform(xe[0] - spread, xe[-1] + spread, n)        # some points lie outside the diagram
    y = rng.uniform(ye[0] - spread, ye[-1] + spread, n)
    sx = rng.uniform(0.2, 4.0, n) * (xe[1] - xe[0])
    sy = rng.uniform(0.2, 4.0, n) * (ye[1] - ye[0])
    w = rng.uniform(0.1, 5.0, n)
    return x, y, sx, sy, w


def _oracle_stack(pls, xe, ye):
    nx, ny = len(xe) - 1, len(ye) - 1
    cols = [O.bin_cmd_smooth(*pl[:4], pl[5], pl[4], nx, xe[0], (xe[-1] - xe[0]) / nx, ny, ye[0], (ye[-1] - ye[0]) / ny).reshape(-1, order="F")
            for pl in pls]
    return np.asfortranarray(np.stack(cols, axis=1))


@pytest.mark.parametrize("nx,ny,T,npts", [(40, 50, 7, 60), (75, 100, 24, 300), (200, 300, 5, 500), (33, 1200, 3, 200)])
def test_device_scatter_matches_oracle(S, nx, ny, T, npts):
    rng = np.random.default_rng(nx * 1000 + ny)
    xe, ye = np.linspace(-0.2, 1.2, nx + 1), np.linspace(19.0, 30.0, ny + 1)
    pls = []
    for t in range(T):
        n = int(rng.integers(0, npts)) if t != 1 else 0                 # ragged lists; template 1 is empty
        pls.append((*_points(rng, n, xe, ye), (0, 1, -1)[t % 3]))
    ds = S.DeviceStack.from_points((xe, ye), pls)
    assert ds.shape == (nx * ny, T)
    M, _ = ds.download()
    ref = _oracle_stack(pls, xe, ye)
    scale = np.abs(ref).max(axis=0, keepdims=True) + 1e-300
    # 3e-12 of each template's peak: both sides round pixel-edge coordinates (values ~1, widths ~1e-2) at 1e-16 and a pixel
    # sums hundreds of points; measured 1e-13 .. 1.6e-12 (profiles/r1_templates.txt)
    assert np.all(np.abs(M - ref) <= 3e-12 * scale), np.max(np.abs(M - ref) / scale)
    assert not M[:, 1].any()
    # deterministic: a second build is bit-identical
    M2, _ = S.DeviceStack.from_points((xe, ye), pls).download()
    assert np.array_equal(M, M2)


def test_stack_from_points_is_a_drop_in_for_the_uploaded_stack(S):
    rng = np.random.default_rng(3)
    nx, ny, T = 60, 70, 40
    xe, ye = np.linspace(0.0, 1.5, nx + 1), np.linspace(20.0, 27.0, ny + 1)
    pls = [(*_points(rng, 150, xe, ye, spread=0.0), (0, 1, -1)[t % 3]) for t in range(T)]
    ref = _oracle_stack(pls, xe, ye)
    x = rng.uniform(0.5, 2.0, T)
    data = rng.poisson(ref @ x).astype(np.float64)
    for dtype, rtol in ((np.float64, 1e-11), (np.float32, 1e-5)):
        a = S.DeviceStack.from_points((xe, ye), pls, data=data.reshape((nx, ny), order="F"), dtype=dtype)
        b = S.DeviceStack(ref.astype(dtype), data)
        assert a.info().fused == b.info().fused == 1
        (fa, Ga, _), (fb, Gb, _) = a.eval_fg(x), b.eval_fg(x)
        assert fa == pytest.approx(fb, rel=rtol) and np.allclose(Ga, Gb, rtol=rtol, atol=rtol * np.abs(Gb).max())
    # a bin-row shard builds only its rows (multi-GPU creation path)
    sh = S.DeviceStack.from_points((xe, ye), pls, data=data, rows=(1000, 2500))
    Ms, ds_ = sh.download()
    full, _ = S.DeviceStack.from_points((xe, ye), pls, data=data).download()
    assert np.array_equal(Ms, full[1000:2500]) and np.array_equal(ds_, data[1000:2500])


def test_bin_cmd_smooth_and_partial_cmd_smooth(S):
    """Shapes / non-zero content as the reference tests them (test/templates/template_test.jl:55-107), on a synthetic
    isochrone, plus agreement of the host preparation + device scatter with the oracle scatter."""
    T = S.templates
    m_ini = np.linspace(0.1, 1.4, 400)
    F1 = -0.6 - 8.2 * np.log10(m_ini) + 0.05 * np.sin(6 * m_ini)       # absolute magnitudes: colours F1 - F2 in ~[0.3, 1.1]
    F2 = -1.0 - 7.5 * np.log10(m_ini)
    F3 = -1.5 - 7.0 * np.log10(m_ini)
    dmod = 25.0
    edges = (np.linspace(-0.2, 1.2, 75), np.linspace(dmod - 6.0, dmod + 5.0, 100))
    comp = [lambda m, m50=m50: T.Martin2016_complete(m, 1.0, m50, 0.7) for m50 in (28.5, 27.5, 26.5)]
    err = [lambda m, c=c: np.minimum(T.exp_photerr(m, 1.03, 15.0, c, 0.02), 0.4) for c in (36.0, 35.0, 34.0)]
    imf = lambda m: np.asarray(m) ** -2.35 / 11.0
    for y_index, ci in ((1, (0, 1)), (0, (0, 1)), (2, (0, 1))):
        W, ed = S.partial_cmd_smooth(m_ini, [F1, F2, F3], err, y_index, ci, imf, comp, dmod=dmod, normalize_value=1e7, mean_mass=0.35,
                                     edges=edges)
        assert W.shape == (74, 99) and W.dtype == np.float64 and W.any() and np.all(W >= 0)
        pts = S.template_points(m_ini, [F1, F2, F3], err, y_index, ci, imf, comp, None, dmod, 1e7, 0.35, edges)
        assert pts[5] == (1 if y_index == 1 else (-1 if y_index == 0 else 0))
        ref = O.bin_cmd_smooth(*pts[:4], pts[5], pts[4], 74, edges[0][0], edges[0][1] - edges[0][0], 99, edges[1][0], edges[1][1] - edges[1][0])
        assert np.all(np.abs(W - ref) <= 3e-12 * ref.max())
    with pytest.raises(NotImplementedError):
        S.partial_cmd_smooth(m_ini, [F1, F2], err[:2], 1, (0, 1), imf, comp[:2], binary_model=object(), mean_mass=0.35, edges=edges)
    with pytest.raises(ValueError):
        S.bin_cmd_smooth([0.1, 0.2], [20.0, 21.0], [0.1, 0.1], [0.1, 0.1], 2, edges=edges)
    # whole grid in one pass == template by template
    isos = [(m_ini, [F1 + 0.1 * k, F2 + 0.05 * k]) for k in range(6)]
    ds = S.build_template_stack(isos, err[:2], 1, (0, 1), imf, comp[:2], dmod=dmod, normalize_value=1e7, mean_mass=0.35, edges=edges)
    M, _ = ds.download()
    for k in (0, 5):
        Wk, _ = S.partial_cmd_smooth(*isos[k], err[:2], 1, (0, 1), imf, comp[:2], dmod=dmod, normalize_value=1e7, mean_mass=0.35, edges=edges)
        assert np.array_equal(M[:, k], Wk.reshape(-1, order="F"))
