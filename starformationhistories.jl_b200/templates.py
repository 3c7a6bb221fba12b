"""Template construction: the step before the fitting hot path.

Host mirror of the reference's template builders with the Gaussian-kernel scatter -- ``bin_cmd_smooth`` /
``addstar!`` (src/StarFormationHistories.jl:348-408, 574-621) -- running on the device and, for whole template
grids, writing straight into the device stack (``DeviceStack.from_points``).  The per-point preparation
(isochrone resampling, photometric error / completeness / bias callables, IMF weights; :829-911) is host numpy,
as it is host Julia in the reference; callables must accept numpy arrays.
"""
from __future__ import annotations

import math

import numpy as np

from .fitting import DeviceStack


# ------------------------------------------------------------------------------------------------
def midpoints(v):
    """midpoints (src/StarFormationHistories.jl:417-436): midpoints between consecutive entries."""
    v = np.asarray(v, dtype=np.float64)
    return v[:-1] + np.diff(v) / 2


def calculate_edges(edges=None, xlim=None, ylim=None, nbins=None, xwidth=None, ywidth=None):
    """calculate_edges (:451-479).  Returns (xedges, yedges) as uniform arrays."""
    if edges is not None:
        xe, ye = (np.asarray(e, dtype=np.float64) for e in edges)
        return xe, ye
    xlim, ylim = sorted(xlim), sorted(ylim)
    if nbins is None:
        if xwidth is None or ywidth is None:
            raise ValueError("If `edges` and `nbins` are not provided, then `xwidth` and `ywidth` must be provided.")
        nbins = (int(round((xlim[1] - xlim[0]) / xwidth)), int(round((ylim[1] - ylim[0]) / ywidth)))
    return np.linspace(xlim[0], xlim[1], nbins[0]), np.linspace(ylim[0], ylim[1], nbins[1])   # range(...; length=nbins)


def histogram_pix(d, edges):
    """histogram_pix for uniform edges (:503): fractional 1-based pixel position."""
    e = np.asarray(edges, dtype=np.float64)
    return (d - e[0]) / (e[1] - e[0]) + 1


def interpolate_mini(m_ini, mags, new_mini):
    """interpolate_mini (:77-81): linear interpolation of magnitudes in initial mass, after sorting by mass."""
    m = np.asarray(m_ini, dtype=np.float64)
    idx = np.argsort(m, kind="stable")
    return np.interp(np.asarray(new_mini, dtype=np.float64), m[idx], np.asarray(mags, dtype=np.float64)[idx])


def mini_spacing(m_ini, colors, mags, dmag, ret_spacing=False):
    """mini_spacing (:110-147): resample the initial masses so that adjacent CMD points are closer than `dmag`."""
    m, c, y = (np.asarray(a, dtype=np.float64) for a in (m_ini, colors, mags))
    if not (m.shape == c.shape == y.shape):
        raise ValueError("axes(m_ini) == axes(mags) == axes(colors) must hold")
    first = m[0]                                                          # :116 (before sorting, as in the reference)
    idx = np.argsort(m, kind="stable")
    m, c, y = m[idx], c[idx], y[idx]
    d = np.hypot(np.diff(c), np.diff(y))
    interp = d > dmag                                                      # :128
    n = np.where(interp, np.ceil(d / dmag), 1.0).astype(np.int64)          # round(Int, d / dmag, RoundUp)
    step = np.diff(m) / n
    seg = np.repeat(np.arange(n.shape[0]), n)                              # segment of every new point
    j = np.arange(seg.shape[0]) - np.repeat(np.cumsum(n) - n, n) + 1       # 1..n within the segment
    pts = m[:-1][seg] + step[seg] * j                                      # m_ini[i] + mass_step*j  (:133)
    pts = np.where(interp[seg], pts, m[1:][seg])                           # uninterpolated segments push m_ini[i+1] itself (:136)
    new = np.concatenate([[first], pts])
    _, keep = np.unique(new, return_index=True)                            # unique(): first occurrences, original order
    new = new[np.sort(keep)]
    return (new, np.diff(new)) if ret_spacing else new


def calculate_weights(mini, completeness, imf, normalize_value, mean_mass, mini_spacing_=None):
    """calculate_weights (:765-776): trapezoidal IMF mass per isochrone segment x completeness x normalisation."""
    mini = np.asarray(mini, dtype=np.float64)
    comp = np.asarray(completeness, dtype=np.float64)
    if mini.shape != comp.shape:
        raise ValueError("length(mini) == length(completeness) must hold")
    sp = np.diff(mini) if mini_spacing_ is None else np.asarray(mini_spacing_, dtype=np.float64)
    pdf = np.asarray(imf(mini), dtype=np.float64)
    return sp * (pdf[:-1] + pdf[1:]) / 2 * comp[:-1] * normalize_value / mean_mass


def Martin2016_complete(m, A, m50, rho):
    """src/utilities.jl:205"""
    return A / (1 + np.exp((np.asarray(m, dtype=np.float64) - m50) / rho))


def exp_photerr(m, a, b, c, d):
    """src/utilities.jl:218"""
    return a ** (b * (np.asarray(m, dtype=np.float64) - c)) + d


# ------------------------------------------------------------------------------------------------
def bin_cmd(colors, mags, weights=None, edges=None, xlim=None, ylim=None, nbins=None, xwidth=None, ywidth=None):
    """bin_cmd (:544-553): the weighted 2-D histogram of (colour, magnitude) points with left-closed bins [a, b)
    (StatsBase `closed=:left`: a point on the last right edge is NOT counted).  This is how the observed Hess diagram
    `data` of a fit is made.  Returns (weights matrix (nx, ny), (xedges, yedges))."""
    colors = np.asarray(colors, dtype=np.float64)
    mags = np.asarray(mags, dtype=np.float64)
    w = np.ones(colors.shape) if weights is None else np.asarray(weights, dtype=np.float64)
    if not (colors.shape == mags.shape == w.shape):
        raise ValueError("length(colors) == length(mags) == length(weights) must hold")          # :550
    xe, ye = calculate_edges(edges, xlim if xlim is not None else (colors.min(), colors.max()),
                             ylim if ylim is not None else (mags.min(), mags.max()), nbins, xwidth, ywidth)
    ix = np.searchsorted(xe, colors, side="right") - 1
    iy = np.searchsorted(ye, mags, side="right") - 1
    nx, ny = xe.shape[0] - 1, ye.shape[0] - 1
    ok = (ix >= 0) & (ix < nx) & (iy >= 0) & (iy < ny)
    out = np.zeros((nx, ny))
    np.add.at(out, (ix[ok], iy[ok]), w[ok])
    return out, (xe, ye)


def partial_cmd(m_ini, colors, mags, imf, dmod=0.0, normalize_value=1.0, mean_mass=None, edges=None, xlim=None, ylim=None,
                nbins=None, xwidth=None, ywidth=None):
    """partial_cmd (:733-757): the unsmoothed template -- isochrone resampled to 0.01 mag spacing, IMF-weighted, binned."""
    if mean_mass is None:
        raise ValueError("mean_mass (the mean initial mass of the IMF) is required")
    colors = np.asarray(colors, dtype=np.float64)
    mags = np.asarray(mags, dtype=np.float64)
    new_mini, spacing = mini_spacing(m_ini, colors, mags, 0.01, True)                          # :739
    c = interpolate_mini(m_ini, colors, new_mini)
    y = interpolate_mini(m_ini, mags, new_mini) + dmod
    pdf = np.asarray(imf(new_mini), dtype=np.float64)
    w = spacing * (pdf[:-1] + pdf[1:]) / 2 * normalize_value / mean_mass                        # :748-753
    return bin_cmd(c[:-1], y[:-1], weights=w, edges=edges, xlim=xlim if xlim is not None else (colors.min(), colors.max()),
                   ylim=ylim if ylim is not None else (mags.min(), mags.max()), nbins=nbins, xwidth=xwidth, ywidth=ywidth)


def bin_cmd_smooth(colors, mags, color_err, mag_err, cov_mult=0, weights=None, edges=None, xlim=None, ylim=None,
                   nbins=None, xwidth=None, ywidth=None):
    """bin_cmd_smooth (:574-621) on the device: returns (weights matrix (nx, ny), (xedges, yedges))."""
    colors = np.asarray(colors, dtype=np.float64)
    mags = np.asarray(mags, dtype=np.float64)
    if weights is None:
        weights = np.ones(colors.shape)
    xe, ye = calculate_edges(edges, xlim if xlim is not None else (colors.min(), colors.max()),
                             ylim if ylim is not None else (mags.min(), mags.max()), nbins, xwidth, ywidth)
    ds = DeviceStack.from_points((xe, ye), [(colors, mags, color_err, mag_err, weights, cov_mult)], force_unfused=True)
    M, _ = ds.download()
    return M[:, 0].reshape((xe.shape[0] - 1, ye.shape[0] - 1), order="F"), (xe, ye)


def template_points(m_ini, mags, mag_err_funcs, y_index, color_indices, imf, completeness_funcs=None, bias_funcs=None, dmod=0.0,
                    normalize_value=1.0, mean_mass=None, edges=None):
    """The per-point arguments partial_cmd_smooth hands to bin_cmd_smooth (:829-888, NoBinaries):
    (colors, mags, color_err, mag_err, weights, cov_mult).  Indices are 0-based."""
    nm = len(mags)
    completeness_funcs = [lambda m: np.ones_like(m)] * nm if completeness_funcs is None else completeness_funcs
    bias_funcs = [lambda m: np.zeros_like(m)] * nm if bias_funcs is None else bias_funcs
    if len(color_indices) != 2:
        raise ValueError("length(color_indices) == 2 must hold")                       # :840
    if not (nm == len(mag_err_funcs) == len(completeness_funcs) == len(bias_funcs)):
        raise ValueError("length(mags) == length(mag_err_funcs) == length(completeness_funcs) == length(bias_funcs) must hold")
    if mean_mass is None:
        raise ValueError("mean_mass (the mean initial mass of the IMF) is required")
    xe, ye = edges
    c0, c1 = color_indices
    mags = [np.asarray(v, dtype=np.float64) for v in mags]
    colors = mags[c0] - mags[c1]
    dmag = min(xe[1] - xe[0], ye[1] - ye[0])                                           # :852
    new_mini, spacing = mini_spacing(m_ini, colors, mags[y_index], dmag, True)
    iso = [interpolate_mini(m_ini, v, new_mini) + dmod for v in mags]                   # :855
    err = [np.asarray(mag_err_funcs[i](iso[i]), dtype=np.float64) for i in range(nm)]
    comp = np.asarray(completeness_funcs[c0](iso[c0])) * np.asarray(completeness_funcs[c1](iso[c1]))
    if y_index in color_indices:                                                       # :857-869
        other = c1 if c0 == y_index else c0
        color_err = err[other]
        cov_mult = -1 if y_index == c0 else 1
    else:                                                                              # :870-881
        color_err = np.sqrt(err[c0] ** 2 + err[c1] ** 2)
        comp = comp * np.asarray(completeness_funcs[y_index](iso[y_index]))
        cov_mult = 0
    w = calculate_weights(new_mini, comp, imf, normalize_value, mean_mass, spacing)
    biased = [iso[i] + np.asarray(bias_funcs[i](iso[i]), dtype=np.float64) for i in range(nm)]   # :884
    col = biased[c0] - biased[c1]
    return (midpoints(col), midpoints(biased[y_index]), midpoints(color_err), midpoints(err[y_index]), w, cov_mult)


def partial_cmd_smooth(m_ini, mags, mag_err_funcs, y_index, color_indices, imf, completeness_funcs=None, bias_funcs=None,
                       dmod=0.0, normalize_value=1.0, binary_model=None, mean_mass=None, edges=None, xlim=None, ylim=None,
                       nbins=None, xwidth=None, ywidth=None):
    """partial_cmd_smooth (:829-911) for single stars (NoBinaries): returns (weights matrix (nx, ny), edges)."""
    if binary_model is not None:
        raise NotImplementedError("binary models (binary_hess, :627-731) are outside the accelerated path")
    edges = calculate_edges(edges, xlim, ylim, nbins, xwidth, ywidth)
    pts = template_points(m_ini, mags, mag_err_funcs, y_index, color_indices, imf, completeness_funcs, bias_funcs, dmod,
                          normalize_value, mean_mass, edges)
    return bin_cmd_smooth(*pts[:4], pts[5], weights=pts[4], edges=edges)


def build_template_stack(isochrones, mag_err_funcs, y_index, color_indices, imf, completeness_funcs=None, bias_funcs=None,
                         dmod=0.0, normalize_value=1.0, mean_mass=None, edges=None, data=None, dtype=np.float64, **kws):
    """All templates of a fit in one device pass: `isochrones` is a list of (m_ini, mags) pairs -- one per (age,
    metallicity) -- and the result is the DeviceStack the fitting functions take; no host Hess diagram is ever formed."""
    pts = [template_points(m, mg, mag_err_funcs, y_index, color_indices, imf, completeness_funcs, bias_funcs, dmod,
                           normalize_value, mean_mass, edges) for (m, mg) in isochrones]
    return DeviceStack.from_points(edges, pts, data=data, dtype=dtype, **kws)
