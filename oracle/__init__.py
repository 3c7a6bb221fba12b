"""CPU oracle bindings -- TEST INFRASTRUCTURE ONLY.

ctypes front end to ``oracle/libsfhoracle.so`` (built by ``oracle/Makefile`` from
``sfh_oracle.c``), the plain-C restatement of the StarFormationHistories.jl fitting hot path,
plus numpy restatements of the O(T) *adapters* that sit between the reference's
optimizers/samplers and that path.  Only ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import this package; the
product package never does (tests/test_no_oracle_in_product.py enforces it).

Reference lines restated by the adapters here:
  * HMCModel.logdensity_and_gradient        src/fitting/hmc_sample.jl:24-37
  * MCMCModel callable                      src/fitting/mcmc_sample.jl:12-23
  * HierarchicalOptimizer.logdensity_and_gradient
                                            src/fitting/hierarchical/generic_fitting.jl:90-199
  * exptransform                            src/fitting/hierarchical/transformations.jl:44
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libsfhoracle.so")

POWERLAW_MZR, LINEAR_AMR, LOG_AMR = 0, 1, 2
_DT = {"f32": (np.float32, C.c_float), "f64": (np.float64, C.c_double)}


def build(force: bool = False) -> str:
    """Compile the oracle with gcc (seconds).  Building the checker is not using it."""
    if force or not os.path.exists(_LIB_PATH) or any(
        os.path.getmtime(os.path.join(_HERE, f)) > os.path.getmtime(_LIB_PATH)
        for f in ("sfh_oracle.c", "oracle_impl.inc", "Makefile")
    ):
        subprocess.run(["make", "-C", _HERE, "-B", "libsfhoracle.so"], check=True, capture_output=True)
    return _LIB_PATH


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        _lib = C.CDLL(_LIB_PATH)
        _declare(_lib)
    return _lib


def _p(a, ct):
    return a.ctypes.data_as(C.POINTER(ct))


def _declare(L):
    i64, dbl, flt, pint = C.c_int64, C.c_double, C.c_float, C.POINTER(C.c_int)
    for suf, (_, ct) in _DT.items():
        P = C.POINTER(ct)
        getattr(L, f"sfho_composite_{suf}").argtypes = [P, P, P, i64, i64]
        getattr(L, f"sfho_composite_{suf}").restype = None
        getattr(L, f"sfho_loglikelihood_{suf}").argtypes = [P, P, i64]
        getattr(L, f"sfho_loglikelihood_{suf}").restype = ct
        getattr(L, f"sfho_grad_single_{suf}").argtypes = [P, P, P, i64]
        getattr(L, f"sfho_grad_single_{suf}").restype = ct
        getattr(L, f"sfho_grad_inplace_{suf}").argtypes = [P, P, P, P, i64, i64]
        getattr(L, f"sfho_grad_inplace_{suf}").restype = None
        getattr(L, f"sfho_fg_{suf}").argtypes = [C.c_int, C.c_int, P, P, P, P, P, i64, i64]
        getattr(L, f"sfho_fg_{suf}").restype = ct
        getattr(L, f"sfho_calculate_coeffs_{suf}").argtypes = [C.c_int, ct, ct, P, ct, P, i64, P, P, i64, P]
        getattr(L, f"sfho_calculate_coeffs_{suf}").restype = C.c_int
        getattr(L, f"sfho_fg_hier_{suf}").argtypes = [C.c_int, P, C.c_int, P, pint, P, i64, P, P, P, i64, P, P, i64, P]
        getattr(L, f"sfho_fg_hier_{suf}").restype = ct
        getattr(L, f"sfho_disp_gauss_{suf}").argtypes = [ct, ct, ct]
        getattr(L, f"sfho_disp_gauss_{suf}").restype = ct
        getattr(L, f"sfho_disp_gauss_grad_{suf}").argtypes = [ct, ct, ct, P, P]
        getattr(L, f"sfho_disp_gauss_grad_{suf}").restype = None
        getattr(L, f"sfho_mh_mean_{suf}").argtypes = [C.c_int, ct, ct, P, ct]
        getattr(L, f"sfho_mh_mean_{suf}").restype = ct
        getattr(L, f"sfho_mh_grad_{suf}").argtypes = [C.c_int, ct, ct, P, ct, P, P, P]
        getattr(L, f"sfho_mh_grad_{suf}").restype = None
        getattr(L, f"sfho_fg_omp_{suf}").argtypes = [P, P, P, P, P, i64, i64]
        getattr(L, f"sfho_fg_omp_{suf}").restype = dbl
    PD, PF = C.POINTER(dbl), C.POINTER(flt)
    L.sfho_fg_quad.argtypes = [C.c_int, PD, PD, PD, PD, PD, PD, i64, i64]
    L.sfho_fg_quad.restype = dbl
    L.sfho_fg_quad_f32.argtypes = [C.c_int, PD, PD, PD, PF, PF, i64, i64, dbl]
    L.sfho_fg_quad_f32.restype = dbl
    L.sfho_fg_hier_quad.argtypes = [C.c_int, PD, C.c_int, PD, pint, PD, i64, PD, PD, i64, PD, PD, i64, PD]
    L.sfho_fg_hier_quad.restype = dbl
    L.sfho_mcmc_logl_f64.argtypes = [PD, i64, PD, PD, i64, i64, PD]
    L.sfho_bin_cmd_smooth.argtypes = [i64, PD, PD, PD, PD, PD, C.c_int, i64, dbl, dbl, i64, dbl, dbl, PD]
    L.sfho_bin_cmd_smooth.restype = None
    L.sfho_gaussian_int_general.argtypes = [dbl] * 7
    L.sfho_gaussian_int_general.restype = dbl
    L.sfho_gaussian_psf_covariant.argtypes = [dbl] * 10
    L.sfho_gaussian_psf_covariant.restype = dbl
    L.sfho_mcmc_logl_f64.restype = None
    L.sfho_num_threads.restype = C.c_int
    L.sfho_set_num_threads.restype = None
    L.sfho_set_num_threads.argtypes = [C.c_int]


def _suf(dtype) -> str:
    dtype = np.dtype(dtype)
    if dtype == np.float32:
        return "f32"
    if dtype == np.float64:
        return "f64"
    raise TypeError(f"oracle supports float32/float64, got {dtype}")


def _fcol(M, dt):
    """Column-major (Fortran) contiguous copy in dtype dt: the stack_models layout."""
    return np.asfortranarray(np.asarray(M, dtype=dt))


def stack_models(models):
    """src/fitting/utilities.jl:12-13 -- reduce(hcat, map(vec, models)) with Julia's column-major vec."""
    return np.asfortranarray(np.stack([np.asarray(m).reshape(-1, order="F") for m in models], axis=1))


# ------------------------------------------------------------------ core path
def composite(coeffs, M, dtype=np.float64):
    suf = _suf(dtype); dt, ct = _DT[suf]
    M = _fcol(M, dt); nb, nt = M.shape
    c = np.ascontiguousarray(coeffs, dtype=dt)
    if c.shape[0] != nt:
        raise ValueError("axes(coeffs,1) != axes(models,2)")  # fitting_base.jl:59
    out = np.empty(nb, dtype=dt)
    getattr(lib(), f"sfho_composite_{suf}")(_p(out, ct), _p(c, ct), _p(M, ct), nb, nt)
    return out


def loglikelihood(Cm, data, dtype=np.float64):
    suf = _suf(dtype); dt, ct = _DT[suf]
    Cm = np.ascontiguousarray(np.asarray(Cm).reshape(-1, order="F"), dtype=dt)
    d = np.ascontiguousarray(np.asarray(data).reshape(-1, order="F"), dtype=dt)
    if Cm.shape != d.shape:
        raise ValueError("axes(composite) != axes(data)")  # fitting_base.jl:85
    return dt(getattr(lib(), f"sfho_loglikelihood_{suf}")(_p(Cm, ct), _p(d, ct), Cm.shape[0]))


def grad_single(model, Cm, data, dtype=np.float64):
    suf = _suf(dtype); dt, ct = _DT[suf]
    m = np.ascontiguousarray(np.asarray(model).reshape(-1, order="F"), dtype=dt)
    Cm = np.ascontiguousarray(np.asarray(Cm).reshape(-1, order="F"), dtype=dt)
    d = np.ascontiguousarray(np.asarray(data).reshape(-1, order="F"), dtype=dt)
    return dt(getattr(lib(), f"sfho_grad_single_{suf}")(_p(m, ct), _p(Cm, ct), _p(d, ct), m.shape[0]))


def grad_inplace(Cm, M, data, dtype=np.float64):
    """Returns (G, residual): ∇loglikelihood! leaves 1-n/m in `composite` (fitting_base.jl:219)."""
    suf = _suf(dtype); dt, ct = _DT[suf]
    M = _fcol(M, dt); nb, nt = M.shape
    Cm = np.array(np.asarray(Cm).reshape(-1, order="F"), dtype=dt)
    d = np.ascontiguousarray(np.asarray(data).reshape(-1, order="F"), dtype=dt)
    G = np.empty(nt, dtype=dt)
    getattr(lib(), f"sfho_grad_inplace_{suf}")(_p(G, ct), _p(Cm, ct), _p(M, ct), _p(d, ct), nb, nt)
    return G, Cm


def fg(coeffs, M, data, want_F=True, want_G=True, dtype=np.float64):
    """solvers.jl:20-38.  Returns (-logL or None, G or None, composite-after-call)."""
    suf = _suf(dtype); dt, ct = _DT[suf]
    M = _fcol(M, dt); nb, nt = M.shape
    c = np.ascontiguousarray(coeffs, dtype=dt)
    d = np.ascontiguousarray(np.asarray(data).reshape(-1, order="F"), dtype=dt)
    if c.shape[0] != nt or d.shape[0] != nb:
        raise ValueError("shape mismatch")  # solvers.jl:9-12
    G = np.empty(nt, dtype=dt); Cm = np.empty(nb, dtype=dt)
    r = getattr(lib(), f"sfho_fg_{suf}")(int(want_F), int(want_G), _p(G, ct), _p(c, ct), _p(M, ct), _p(d, ct), _p(Cm, ct), nb, nt)
    return (dt(r) if want_F else None), (G if want_G else None), Cm


def fg_quad(coeffs, M, data, want_G=True):
    """__float128 arbiter on Float64 inputs.  Returns (-logL, G, gscale, composite)."""
    M = _fcol(M, np.float64); nb, nt = M.shape
    c = np.ascontiguousarray(coeffs, dtype=np.float64)
    d = np.ascontiguousarray(np.asarray(data).reshape(-1, order="F"), dtype=np.float64)
    G = np.empty(nt); gs = np.empty(nt); Cm = np.empty(nb)
    D = C.c_double
    r = lib().sfho_fg_quad(int(want_G), _p(G, D), _p(gs, D), _p(c, D), _p(M, D), _p(d, D), _p(Cm, D), nb, nt)
    return r, (G if want_G else None), (gs if want_G else None), Cm


def fg_quad_f32(coeffs, M32, data32, want_G=True, eps=float(np.finfo(np.float32).eps)):
    """__float128 arbiter on Float32-STORED templates/data (coeffs Float64); clamp eps(Float32) like a
    Float32 fit in the reference (fitting_base.jl:86,90)."""
    M = _fcol(M32, np.float32); nb, nt = M.shape
    c = np.ascontiguousarray(coeffs, dtype=np.float64)
    d = np.ascontiguousarray(np.asarray(data32).reshape(-1, order="F"), dtype=np.float32)
    G = np.empty(nt); gs = np.empty(nt)
    D = C.c_double
    r = lib().sfho_fg_quad_f32(int(want_G), _p(G, D), _p(gs, D), _p(c, D), _p(M, C.c_float), _p(d, C.c_float), nb, nt, float(eps))
    return r, (G if want_G else None), (gs if want_G else None)


def fg_omp(coeffs, M, data, dtype=np.float64, G=None, Cm=None):
    """Threaded two-pass CPU baseline (bench.py).  M must already be Fortran-ordered in dtype."""
    suf = _suf(dtype); dt, ct = _DT[suf]
    assert M.flags.f_contiguous and M.dtype == dt
    nb, nt = M.shape
    c = np.ascontiguousarray(coeffs, dtype=dt)
    d = np.ascontiguousarray(data, dtype=dt)
    G = np.empty(nt, dtype=dt) if G is None else G
    Cm = np.empty(nb, dtype=dt) if Cm is None else Cm
    r = getattr(lib(), f"sfho_fg_omp_{suf}")(_p(G, ct), _p(c, ct), _p(M, ct), _p(d, ct), _p(Cm, ct), nb, nt)
    return r, G


def fg_blas(coeffs, M, data, dtype=np.float64):
    """The reference's flat fg! by the route Julia itself takes: `mul!(C, M, coeffs)` and `mul!(G, M', C, -1, false)`
    (fitting_base.jl:60, :283) are BLAS gemv 'N' / 'T' calls into OpenBLAS -- here the OpenBLAS bundled with numpy, all its
    threads -- with the Poisson term (:84-96) and the residual (:274-280) as vectorised loops in between; fg! flips the sign
    (solvers.jl:28-31).  The second CPU baseline bench.py times ("B1"); M must be Fortran-ordered in
    `dtype`.  Returns (-logL, G)."""
    dt = np.dtype(dtype).type
    assert M.flags.f_contiguous and M.dtype == dt
    eps = dt(np.finfo(dt).eps)
    c = np.ascontiguousarray(coeffs, dtype=dt)
    d = np.ascontiguousarray(data, dtype=dt)
    m = M @ c                                                              # gemv 'N'
    np.maximum(m, eps, out=m)                                              # :90 / :277
    with np.errstate(divide="ignore", invalid="ignore"):
        q = d / m
        logl = np.where(d > 0, d - m - d * np.log(q), -m).sum(dtype=dt)    # :92 (accumulates in T like the reference, :87)
    r = dt(1) - q                                                          # :279
    G = r @ M                                                              # gemv 'T' (columns of M are contiguous)
    return (float(-logl) if logl != 0 else float("inf")), G


def blas_threads() -> int:
    try:
        from threadpoolctl import threadpool_info
        return max([int(i["num_threads"]) for i in threadpool_info() if i.get("user_api") == "blas"] or [1])
    except Exception:
        return os.cpu_count() or 1


def num_threads() -> int:
    return int(lib().sfho_num_threads())


def set_num_threads(n: int) -> None:
    """omp_set_num_threads for the OpenMP routes (launchers such as torchrun export OMP_NUM_THREADS=1)."""
    lib().sfho_set_num_threads(int(n))


# ------------------------------------------------------------------ hierarchical
def _fixed(kind, fixed):
    """fixed = (logMstar0,) | (T_max,) | (T_max[, solZ, Y_p, gamma]) padded to 4 doubles."""
    fx = [float(v) for v in np.atleast_1d(np.asarray(fixed, dtype=np.float64))]
    if kind == LOG_AMR:
        defaults = [13.7, 0.01524, 0.2485, 1.78]  # amr.jl:277 T_max; src/utilities.jl:138 solZ, Y_p, gamma
        fx = fx + defaults[len(fx):]
    f = np.zeros(4)
    f[: min(4, len(fx))] = fx[:4]
    return f


def disp_gauss(x, mu, sigma):
    """GaussianDispersion(sigma)(x, mu)  dispersion_models.jl:92."""
    return float(lib().sfho_disp_gauss_f64(x, mu, sigma))


def disp_gauss_grad(x, mu, sigma):
    """gradient(GaussianDispersion(sigma), x, mu) -> (dA/dsigma, dA/dmu)  dispersion_models.jl:95-100."""
    a = C.c_double(); b = C.c_double()
    lib().sfho_disp_gauss_grad_f64(x, mu, sigma, C.byref(a), C.byref(b))
    return a.value, b.value


def mh_mean(kind, alpha, beta, fixed, arg):
    fx = _fixed(kind, fixed)
    return float(lib().sfho_mh_mean_f64(kind, alpha, beta, _p(fx, C.c_double), arg))


def mh_grad(kind, alpha, beta, fixed, arg):
    fx = _fixed(kind, fixed)
    a = C.c_double(); b = C.c_double(); m = C.c_double()
    lib().sfho_mh_grad_f64(kind, alpha, beta, _p(fx, C.c_double), arg, C.byref(a), C.byref(b), C.byref(m))
    return a.value, b.value, m.value


def calculate_coeffs(kind, alpha, beta, fixed, sigma, R, logAge, MH, dtype=np.float64):
    """mzr.jl:50-79 / amr.jl:50-73."""
    suf = _suf(dtype); dt, ct = _DT[suf]
    R = np.ascontiguousarray(R, dtype=dt); la = np.ascontiguousarray(logAge, dtype=dt)
    mh = np.ascontiguousarray(MH, dtype=dt)
    if la.shape != mh.shape:
        raise ValueError("length(logAge) != length(metallicities)")  # mzr.jl:57
    fx = _fixed(kind, fixed).astype(dt)
    out = np.empty(la.shape[0], dtype=dt)
    rc = getattr(lib(), f"sfho_calculate_coeffs_{suf}")(kind, ct(alpha), ct(beta), _p(fx, ct), ct(sigma), _p(R, ct), R.shape[0], _p(la, ct), _p(mh, ct), la.shape[0], _p(out, ct))
    if rc != 0:
        raise ValueError("Length of `mstars` must be the same as `unique(logAge)`.")  # mzr.jl:55-56
    return out


def fg_hier(kind, fixed, free3, variables, M, data, logAge, MH, want_G=True, dtype=np.float64, quad=False):
    """mzr.jl:84-215 / amr.jl:78-173.  Returns (-logL, G or None, fullG or None)."""
    suf = _suf(dtype); dt, ct = _DT[suf]
    if quad:
        dt, ct = np.float64, C.c_double
    M = _fcol(M, dt); nb, nt = M.shape
    v = np.ascontiguousarray(variables, dtype=dt); nj = v.shape[0] - 3
    d = np.ascontiguousarray(np.asarray(data).reshape(-1, order="F"), dtype=dt)
    la = np.ascontiguousarray(logAge, dtype=dt); mh = np.ascontiguousarray(MH, dtype=dt)
    fx = _fixed(kind, fixed).astype(dt)
    fr = np.ascontiguousarray(free3, dtype=np.int32)
    G = np.empty(nj + 3, dtype=dt); fullG = np.empty(nt, dtype=dt)
    if quad:
        r = lib().sfho_fg_hier_quad(int(want_G), _p(G, ct), kind, _p(fx, ct), _p(fr, C.c_int), _p(v, ct), nj, _p(M, ct), _p(d, ct), nb, _p(la, ct), _p(mh, ct), nt, _p(fullG, ct))
    else:
        Cm = np.empty(nb, dtype=dt)
        r = getattr(lib(), f"sfho_fg_hier_{suf}")(int(want_G), _p(G, ct), kind, _p(fx, ct), _p(fr, C.c_int), _p(v, ct), nj, _p(M, ct), _p(d, ct), _p(Cm, ct), nb, _p(la, ct), _p(mh, ct), nt, _p(fullG, ct))
    return dt(r), (G if want_G else None), (fullG if want_G else None)


# transforms(): PowerLawMZR (1,0) mzr.jl:283; LinearAMR (1,0) amr.jl:212; LogarithmicAMR (1,1) amr.jl:297;
# GaussianDispersion (1,) dispersion_models.jl:107
TRANSFORMS = {POWERLAW_MZR: (1, 0, 1), LINEAR_AMR: (1, 0, 1), LOG_AMR: (1, 1, 1)}


def hier_logdensity_and_gradient(kind, fixed, free3, init_params, xvec, M, data, logAge, MH,
                                 jacobian_corrections=True, want_G=True):
    """generic_fitting.jl:90-199.  `xvec` = [log R_1..log R_Nj, transformed FREE params].
    init_params = (alpha0, beta0, sigma0) supply the fixed ones (:134-136).
    Returns (+logp, +grad over free variables)."""
    tf = np.array(TRANSFORMS[kind]); free = np.array(free3, dtype=bool)
    xvec = np.asarray(xvec, dtype=np.float64)
    nfixed = int((~free).sum())
    nbins = xvec.shape[0] - 3 + nfixed                                   # :116-118
    x = np.empty(nbins + 3)
    x[:nbins] = np.exp(xvec[:nbins])                                     # :127
    par = xvec[nbins:]
    tff = tf[free]
    xz = np.where(tff == 1, np.exp(par), np.where(tff == -1, -np.exp(par), par))  # exptransform
    x[nbins:][free] = xz
    x[nbins:][~free] = np.asarray(init_params, dtype=np.float64)[~free]  # :134-136
    nlogL, G2, _ = fg_hier(kind, fixed, free3, x, M, data, logAge, MH, want_G=want_G)
    nlogL = float(nlogL)
    ptf = [i for i in range(3) if tf[i] == 1 and free[i]]
    idxs = list(range(nbins)) + [nbins + i for i in ptf]
    if jacobian_corrections:
        for i in idxs:                                                   # :149-152
            nlogL -= np.log(x[i])
            if want_G:
                G2[i] = G2[i] * x[i] - 1
    elif want_G:
        for i in idxs:                                                   # :162-164
            G2[i] = G2[i] * x[i]
    if not want_G:
        return -nlogL, None
    G = np.empty_like(xvec)
    G[:nbins] = G2[:nbins]
    G[nbins:] = G2[nbins:][free]                                         # :181-189
    return -nlogL, -G                                                    # :193-194


def hmc_logdensity_and_gradient(logx, M, data):
    """hmc_sample.jl:24-37."""
    logx = np.asarray(logx, dtype=np.float64)
    x = np.exp(logx)
    nl, G, _ = fg(x, M, data)
    return -float(nl) + logx.sum(), -G * x + 1


def mcmc_logl(X, M, data):
    """mcmc_sample.jl:12-23 for each column of X (nt x W)."""
    M = _fcol(M, np.float64); nb, nt = M.shape
    X = np.asfortranarray(np.asarray(X, dtype=np.float64).reshape(nt, -1))
    d = np.ascontiguousarray(np.asarray(data).reshape(-1, order="F"), dtype=np.float64)
    out = np.empty(X.shape[1])
    D = C.c_double
    lib().sfho_mcmc_logl_f64(_p(X, D), X.shape[1], _p(M, D), _p(d, D), nb, nt, _p(out, D))
    return out


def bin_cmd_smooth(colors, mags, color_err, mag_err, cov_mult, weights, nx, xfirst, xstep, ny, yfirst, ystep, out=None):
    """bin_cmd_smooth (src/StarFormationHistories.jl:574-621) onto nx x ny uniform bins; returns the (nx, ny) weights matrix."""
    L = lib()
    a = [np.ascontiguousarray(v, dtype=np.float64) for v in (colors, mags, color_err, mag_err, weights)]
    n = a[0].shape[0]
    if any(v.shape != (n,) for v in a):
        raise ValueError("axes(colors) == axes(mags) == axes(color_err) == axes(mag_err) == axes(weights) must hold")
    img = np.zeros((nx, ny), order="F") if out is None else out
    PD = C.POINTER(C.c_double)
    L.sfho_bin_cmd_smooth(n, *[v.ctypes.data_as(PD) for v in a], int(cov_mult), nx, float(xfirst), float(xstep), ny, float(yfirst),
                          float(ystep), img.ctypes.data_as(PD))
    return img


def gaussian_int_general(dx, dy, hx, hy, sx, sy, A):
    return lib().sfho_gaussian_int_general(dx, dy, hx, hy, sx, sy, A)


def gaussian_psf_covariant(x, y, hx, hy, x0, y0, sx, sy, cov_mult, A):
    return lib().sfho_gaussian_psf_covariant(x, y, hx, hy, x0, y0, sx, sy, cov_mult, A)
