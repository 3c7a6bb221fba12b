"""Oracle restatement of the template builders (bin_cmd_smooth / addstar! / the two Gaussian kernels,
src/StarFormationHistories.jl:198-205, 315-333, 348-408, 574-621) pinned against independent quadrature, and the host
helpers against the reference's doctests (:45-75, :489-502) and test values (test/templates/template_test.jl:10-33).
The reference's own kernel tests (test/templates/kernel_test.jl) only assert "some pixel is non-zero"."""
import math

import numpy as np
import pytest
from scipy import integrate

import oracle as O
import sfh_b200 as S
from sfh_b200 import templates as T


def test_asymmetric_kernel_is_the_exact_pixel_integral():
    rng = np.random.default_rng(1)
    for _ in range(20):
        dx, dy = rng.normal(0, 2, 2)
        sx, sy = rng.uniform(0.3, 3, 2)
        hx, hy = rng.uniform(0.2, 0.8, 2)
        A = rng.uniform(0.5, 3)
        f = lambda y, x: A / (2 * math.pi * sx * sy) * math.exp(-x * x / (2 * sx * sx) - y * y / (2 * sy * sy))
        want, _ = integrate.dblquad(f, dx - hx, dx + hx, dy - hy, dy + hy, epsabs=1e-13, epsrel=1e-12)
        assert O.gaussian_int_general(dx, dy, hx, hy, sx, sy, A) == pytest.approx(want, rel=1e-9, abs=1e-15)


@pytest.mark.parametrize("cov", [-1.0, 0.0, 1.0])
def test_covariant_kernel_against_quadrature(cov):
    rng = np.random.default_rng(2)
    for _ in range(12):
        x0, y0 = rng.normal(0, 1, 2)
        sx, sy = rng.uniform(0.5, 2, 2)
        hx, hy = 0.05, 0.05                                   # pixels well inside one sigma: 3-point Gauss-Legendre is accurate
        x, y = x0 + rng.normal(0, 1), y0 + rng.normal(0, 1)
        A = 1.7

        def f(yy, xx):
            dyy = yy - y0
            dxx = (xx - x0) + dyy * cov
            return A / (2 * math.pi * sx * sy) * math.exp(-dyy * dyy / (2 * sy * sy) - dxx * dxx / (2 * sx * sx))
        want, _ = integrate.dblquad(f, x - hx, x + hx, y - hy, y + hy, epsabs=1e-14, epsrel=1e-11)
        assert O.gaussian_psf_covariant(x, y, hx, hy, x0, y0, sx, sy, cov, A) == pytest.approx(want, rel=1e-6)


@pytest.mark.parametrize("cov", [0, 1, -1])
def test_bin_cmd_smooth_mass_and_cutouts(cov):
    nx, ny, xf, xs, yf, ys = 60, 80, -0.5, 0.025, 18.0, 0.05
    # a point well inside: the image integrates to its weight (the cut-out is +-5 sigma / +-7.5 sigma)
    img = O.bin_cmd_smooth([0.2], [20.0], [0.04], [0.1], cov, [3.0], nx, xf, xs, ny, yf, ys)
    assert img.sum() == pytest.approx(3.0, rel=2e-3 if cov == 0 else 1e-2) and img.shape == (nx, ny)   # sheared tails clip
    cx, cy = np.unravel_index(img.argmax(), img.shape)
    assert abs(cx - (0.2 - xf) / xs) <= 1.5 and abs(cy - (20.0 - yf) / ys) <= 1.5
    # accumulation is linear in the weights and additive over points, in order
    a = O.bin_cmd_smooth([0.2, 0.3], [20.0, 20.5], [0.04, 0.02], [0.1, 0.07], cov, [1.0, 2.0], nx, xf, xs, ny, yf, ys)
    b = O.bin_cmd_smooth([0.2], [20.0], [0.04], [0.1], cov, [1.0], nx, xf, xs, ny, yf, ys)
    b = O.bin_cmd_smooth([0.3], [20.5], [0.02], [0.07], cov, [2.0], nx, xf, xs, ny, yf, ys, out=b)
    assert np.array_equal(a, b)
    # a point far outside the diagram contributes nothing; one straddling the edge contributes part of its mass
    out = O.bin_cmd_smooth([5.0], [20.0], [0.04], [0.1], cov, [1.0], nx, xf, xs, ny, yf, ys)
    assert not out.any()
    edge = O.bin_cmd_smooth([xf], [20.0], [0.04], [0.1], cov, [1.0], nx, xf, xs, ny, yf, ys)
    assert 0.3 < edge.sum() < 0.7
    with pytest.raises(ValueError):
        O.bin_cmd_smooth([0.2, 0.3], [20.0], [0.04], [0.1], cov, [1.0], nx, xf, xs, ny, yf, ys)


def test_pixel_space_cutout_rule():
    """addstar! (:348-364): cut-out half-widths max(1, ceil(10 sigma_pix) / 2) around round(centroid)."""
    nx, ny = 40, 40
    img = O.bin_cmd_smooth([10.3], [20.6], [0.31], [0.29], 0, [1.0], nx, 0.0, 1.0, ny, 0.0, 1.0)
    ix, iy = np.nonzero(img)
    # x0 = 11.3 -> x = 11, offset = ceil(3.1) // 2 = 2 -> pixels 9..13 (1-based); y0 = 21.6 -> 22, offset max(1, ceil(2.9) // 2) = 1
    assert (ix.min() + 1, ix.max() + 1) == (9, 13) and (iy.min() + 1, iy.max() + 1) == (21, 23)


def test_host_helpers_doctests():
    m = [0.08, 0.10, 0.12, 0.14, 0.16]
    mg = [13.545, 12.899, 12.355, 11.459, 10.947]
    want = [13.545, 13.222, 12.899, 12.626999999999999, 12.355, 11.907, 11.459, 11.203, 10.947]
    assert np.allclose(T.interpolate_mini(m, mg, np.arange(0.08, 0.1601, 0.01)), want)          # :57-66
    r = T.mini_spacing(m, [1.0, 0.99, 0.98, 0.97, 0.96], mg, 0.1)                               # template_test.jl:10-22
    assert len(r) > 5 and np.diff(r).max() < 0.1
    r2, sp = T.mini_spacing(m, [1.0, 0.99, 0.98, 0.97, 0.96], mg, 0.1, True)
    assert np.allclose(sp, np.diff(r2)) and np.array_equal(r, r2)
    assert np.allclose(T.midpoints(np.arange(0.5, 1.01, 0.1)), np.arange(0.55, 0.96, 0.1))      # template_test.jl:26-31
    assert np.allclose(T.midpoints([1.0, 2.0, 2.2, 2.1]), [1.5, 2.1, 2.15])
    assert T.histogram_pix(0.5, np.arange(0, 1.01, 0.1)) == pytest.approx(6)                    # :489-496
    assert T.histogram_pix(0.55, np.arange(0, 1.01, 0.1)) == pytest.approx(6.5)
    assert T.Martin2016_complete(28.5, 1.0, 28.5, 0.7) == pytest.approx(0.5)
    assert T.exp_photerr(36.0, 1.03, 15.0, 36.0, 0.02) == pytest.approx(1.02)
    # the reference's own known answers (test/utilities/utilities_test.jl:44-47, BigFloat literals)
    assert T.Martin2016_complete(20.0, 1.0, 25.0, 1.0) == pytest.approx(0.9933071490757151444406380196186748, rel=1e-15)
    assert T.exp_photerr(20.0, 1.05, 10.0, 32.0, 0.01) == pytest.approx(0.01286605230281143891186877135084309, rel=1e-13)
    xe, ye = T.calculate_edges(None, (1.5, -1.0), (22.0, 27.2), (26, 53))
    assert xe.shape == (26,) and ye.shape == (53,) and xe[0] == -1.0 and ye[-1] == 27.2
    with pytest.raises(ValueError):
        T.calculate_edges(None, (0, 1), (0, 1))
    # calculate_weights: trapezoid of the IMF pdf over each mass segment (:765-776)
    w = T.calculate_weights([1.0, 2.0, 4.0], [0.5, 1.0, 1.0], lambda mm: np.asarray(mm) ** -2.0, 10.0, 2.0)
    assert np.allclose(w, [1.0 * (1 + 0.25) / 2 * 0.5 * 5, 2.0 * (0.25 + 0.0625) / 2 * 1.0 * 5])


def test_template_points_follow_partial_cmd_smooth():
    """The per-point arguments handed to bin_cmd_smooth (src/StarFormationHistories.jl:829-888): kernel choice, colour error,
    completeness product, IMF weights, bias, midpoints -- host numpy, no device."""
    m = np.linspace(0.2, 1.2, 60)
    B, Vm, R = 6.0 - 7.0 * np.log10(m), 5.5 - 6.5 * np.log10(m), 5.0 - 6.0 * np.log10(m)
    edges = (np.linspace(-0.5, 1.5, 41), np.linspace(20.0, 32.0, 61))
    err = [lambda x: 0.01 + 0.0 * x, lambda x: 0.02 + 0.0 * x, lambda x: 0.03 + 0.0 * x]
    comp = [lambda x: 0.9 + 0.0 * x, lambda x: 0.8 + 0.0 * x, lambda x: 0.5 + 0.0 * x]
    bias = [lambda x: 0.1 + 0.0 * x, lambda x: 0.0 * x, lambda x: -0.2 + 0.0 * x]
    imf = lambda mm: np.asarray(mm) ** -2.35
    kw = dict(dmod=24.0, normalize_value=1e4, mean_mass=0.5, edges=edges)
    # y = V, x = B - V: covariant kernel, cov_mult = +1, colour error = sigma_B alone, completeness = product of B and V (:857-869)
    c, y, ce, ye, w, cov = S.template_points(m, [B, Vm, R], err, 1, (0, 1), imf, comp, bias, **kw)
    n = c.shape[0]
    assert cov == 1 and y.shape == ce.shape == ye.shape == w.shape == (n,)
    assert np.allclose(ce, 0.01) and np.allclose(ye, 0.02)
    new_m, sp = T.mini_spacing(m, B - Vm, Vm, min(0.05, 0.2), True)
    assert n == new_m.shape[0] - 1
    pdf = imf(new_m)
    assert np.allclose(w, sp * (pdf[:-1] + pdf[1:]) / 2 * (0.9 * 0.8) * 1e4 / 0.5)                     # :765-776
    Bi, Vi = T.interpolate_mini(m, B, new_m) + 24.0, T.interpolate_mini(m, Vm, new_m) + 24.0
    assert np.allclose(c, T.midpoints((Bi + 0.1) - Vi)) and np.allclose(y, T.midpoints(Vi))           # bias: measured = intrinsic + bias
    # y = B: cov_mult = -1 and the colour error is sigma_V
    _, _, ce, ye, _, cov = S.template_points(m, [B, Vm, R], err, 0, (0, 1), imf, comp, bias, **kw)
    assert cov == -1 and np.allclose(ce, 0.02) and np.allclose(ye, 0.01)
    # y = R (not in the colour): separable kernel, colour error in quadrature, completeness of all three filters (:870-881)
    _, y, ce, ye, w3, cov = S.template_points(m, [B, Vm, R], err, 2, (0, 1), imf, comp, bias, **kw)
    assert cov == 0 and np.allclose(ce, np.hypot(0.01, 0.02)) and np.allclose(ye, 0.03)
    assert w3.sum() == pytest.approx(w.sum() * 0.5, rel=0.05)                                          # extra completeness factor 0.5
    with pytest.raises(ValueError):
        S.template_points(m, [B, Vm], err, 1, (0, 1), imf, comp, bias, **kw)                          # length(mags) mismatch (:841)
    with pytest.raises(ValueError):
        S.template_points(m, [B, Vm, R], err, 1, (0, 1, 2), imf, comp, bias, **kw)                    # length(color_indices) == 2 (:840)


def test_bin_cmd_and_partial_cmd():
    """bin_cmd (:544-553): left-closed bins, weights; partial_cmd (:733-757): total weight = IMF mass of the isochrone span."""
    xe, ye = np.linspace(0.0, 1.0, 5), np.linspace(10.0, 12.0, 3)          # 4 x 2 bins
    c = np.array([0.0, 0.24, 0.25, 0.99, 1.0, -0.1, 0.5])
    y = np.array([10.0, 10.99, 11.0, 11.99, 11.5, 10.5, 12.0])
    H, ed = S.bin_cmd(c, y, edges=(xe, ye))
    want = np.zeros((4, 2)); want[0, 0] = 2; want[1, 1] = 1; want[3, 1] = 1      # (1.0, .) and (., 12.0) sit on the excluded right edges
    assert np.array_equal(H, want) and ed[0] is not None
    H2, _ = S.bin_cmd(c, y, weights=np.arange(7.0), edges=(xe, ye))
    assert H2[0, 0] == 0 + 1 and H2[1, 1] == 2 and H2[3, 1] == 3 and H2.sum() == 6
    with pytest.raises(ValueError):
        S.bin_cmd(c, y[:-1], edges=(xe, ye))
    m = np.linspace(0.2, 1.2, 50)
    col, mag = 1.0 - 0.5 * np.log10(m), 6.0 - 7.0 * np.log10(m)
    imf = lambda mm: np.asarray(mm) ** -2.35
    H3, _ = S.partial_cmd(m, col, mag, imf, dmod=20.0, normalize_value=100.0, mean_mass=0.5, edges=(np.linspace(0.5, 1.6, 23), np.linspace(24.0, 32.0, 81)))
    mass = (1.2 ** -1.35 - 0.2 ** -1.35) / -1.35                         # integral of the pdf over the isochrone's mass range
    assert H3.sum() == pytest.approx(mass * 100.0 / 0.5, rel=2e-3) and np.count_nonzero(H3) > 40
