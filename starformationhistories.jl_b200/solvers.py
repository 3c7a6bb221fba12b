"""Host-side mirrors of the reference DRIVERS that own the iteration loop (SURVEY.md section 8f "next" rows):

    renormalize_x0           src/fitting/utilities.jl:104-121
    fit_templates_lbfgsb     src/fitting/solvers.jl:70-90     (LBFGSB.jl  -> scipy's L-BFGS-B 3.0, same Fortran lineage)
    fit_templates            src/fitting/solvers.jl:163-221   (Optim BFGS on log-coefficients: MAP then MLE)
    fit_templates_fast       src/fitting/solvers.jl:238-275   (Optim BFGS on sqrt-coefficients: MLE only)
    fit_sfh                  src/fitting/hierarchical/generic_fitting.jl:242-411 (BFGS on the HierarchicalOptimizer)
    mcmc_sample              src/fitting/mcmc_sample.jl:97-130 (KissMCMC.emcee -> affine-invariant stretch move;
                             each half-ensemble is ONE batched device call instead of W/2 host gemv's)
    hmc_sample               src/fitting/hmc_sample.jl:105-143 (DynamicHMC NUTS -> a compact NUTS with dual averaging)

The optimisation / sampling ENGINES are third-party in the reference too (LBFGSB, Optim, DynamicHMC, KissMCMC);
they are replaced by scipy or by short textbook implementations and are NOT part of the parity claim --
per-iterate trajectories differ between engines, converged answers and posterior moments do not.  Every
objective / gradient / log-likelihood evaluation goes through the device path (`fg_`, `HierarchicalOptimizer`,
`MCMCModel.batch`, `HMCModel`); nothing here evaluates the model on the host.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import ctypes as C

import numpy as np
from scipy import optimize

from . import _lib as L
from .fitting import DeviceStack, DeviceStackGroup, composite_, device_stack, fg_ as _fg_flat
from .hierarchical import HierarchicalOptimizer, calculate_coeffs, logtransform, exptransform
from .sampling import HMCModel, MCMCModel


# ---------------------------------------------------------------------------------------------
def renormalize_x0(data, models, x0, full_coeffs=None):
    """Scale x0 so that sum(composite(full_coeffs)) == sum(data)  (fitting/utilities.jl:104-115)."""
    x0 = np.asarray(x0, dtype=np.float64)
    full = x0 if full_coeffs is None else np.asarray(full_coeffs, dtype=np.float64)
    ds = device_stack(models, data)
    if isinstance(ds, DeviceStackGroup):
        csum = float(full @ ds.column_sums())       # sum_i (M full)_i without gathering the sharded composite
    else:
        comp = np.empty(ds.rows)
        composite_(comp, full, ds)
        csum = comp.sum()
    if csum == 0:
        return x0.copy()                                                   # :112
    return x0 * (np.asarray(data, dtype=np.float64).sum() / csum)          # :113-114


def _check_sizes(x0, ds):
    if np.asarray(x0).shape[0] != ds.shape[1]:
        raise ValueError("axes(coeffs,1) != axes(models,2)")               # solvers.jl:10


# ---------------------------------------------------------------------------------------------
def _lbfgsb_opts(m, factr, pgtol, maxiter, maxfun):
    o = L.sfh_lbfgsb_opts()
    o.struct_size = C.sizeof(L.sfh_lbfgsb_opts)
    o.m, o.factr, o.pgtol, o.maxiter, o.maxfun = int(m), float(factr), float(pgtol), int(maxiter), int(maxfun)
    return o


def native_lbfgsb(fun, x0, lb=None, ub=None, m=10, factr=1e-12, pgtol=1e-5, maxiter=100000, maxfun=100000):
    """The library's L-BFGS-B loop (sfh_minimize_lbfgsb) on a Python objective ``fun(x) -> (f, grad)``; `factr` is in units of
    machine epsilon like LBFGSB.jl's.  Returns (x, f, info) with info = {"nit", "funcalls", "pg_norm", "status"}."""
    x = np.array(x0, dtype=np.float64)
    n = x.shape[0]
    err = []

    def cb(_user, xp, nn, fp, gp):
        try:
            f, g = fun(np.ctypeslib.as_array(xp, shape=(nn,)).copy())
            fp[0] = float(f)
            np.ctypeslib.as_array(gp, shape=(nn,))[:] = g
            return 0
        except Exception as e:
            err.append(e)
            return L.SFH_ERR_INVALID_ARG
    dp = C.POINTER(C.c_double)
    lo = np.ascontiguousarray(np.broadcast_to(np.asarray(lb, dtype=np.float64), (n,))) if lb is not None else None
    hi = np.ascontiguousarray(np.broadcast_to(np.asarray(ub, dtype=np.float64), (n,))) if ub is not None else None
    rep, o = L.sfh_lbfgsb_report(), _lbfgsb_opts(m, factr, pgtol, maxiter, maxfun)
    st = L.lib.sfh_minimize_lbfgsb(L.sfh_objective_fn(cb), None, n, x.ctypes.data_as(dp), lo.ctypes.data_as(dp) if lo is not None else None,
                                   hi.ctypes.data_as(dp) if hi is not None else None, C.byref(o), C.byref(rep))
    if err:
        raise err[0]
    L.check(st)
    return x, rep.f, {"nit": int(rep.iterations), "funcalls": int(rep.f_calls), "pg_norm": rep.pg_norm, "status": int(rep.status)}


def fit_templates_lbfgsb(models, data, x0=None, factr=1e-12, pgtol=1e-5, iprint=0, engine="scipy", **kws):
    """Returns (-logL, coeffs): box-constrained (coeffs >= 0) L-BFGS-B on fg!  (solvers.jl:82-90).
    engine="native": the whole optimisation is one call into the library (sfh_fit_templates_lbfgsb)."""
    _check_engine(engine)
    ds = device_stack(models, data)
    x0 = np.ones(ds.shape[1]) if x0 is None else np.asarray(x0, dtype=np.float64)
    _check_sizes(x0, ds)
    x0 = renormalize_x0(data, ds, x0)                                      # :86
    if engine == "native":
        x = np.array(x0, dtype=np.float64)
        rep = L.sfh_lbfgsb_report()
        o = _lbfgsb_opts(kws.get("m", 10), factr, pgtol, kws.get("maxiter", 100000), kws.get("maxfun", 100000))
        L.check(L.lib.sfh_fit_templates_lbfgsb(ds.ctx().handle, x.ctypes.data_as(C.POINTER(C.c_double)), C.byref(o), C.byref(rep)))
        return rep.f, x
    G = np.empty(ds.shape[1])

    def fg(x):                                                             # :88
        return float(_fg_flat(True, G, x, ds, data)), G.copy()

    kws.setdefault("m", 10)
    kws.setdefault("maxfun", 100000)
    kws.setdefault("maxiter", 100000)
    # (`iprint` is accepted for signature parity; recent scipy dropped the L-BFGS-B print switch)
    x, f, info = optimize.fmin_l_bfgs_b(fg, x0, bounds=[(0.0, None)] * ds.shape[1], factr=factr, pgtol=pgtol, **kws)
    return f, x


@dataclass
class LogTransformFTResult:
    """solvers.jl:115-129: mu (natural units), sigma = sqrt(diag(invH)) * mu, invH in log space, engine result."""
    mu: np.ndarray
    sigma: np.ndarray
    invH: np.ndarray
    result: object

    def rand(self, rng, n):
        z = rng.multivariate_normal(np.log(self.mu), (self.invH + self.invH.T) / 2, size=n)   # :127-128
        return np.exp(z).T


def _bfgs(fun, x0, gtol=1e-8, maxiter=5000):
    return optimize.minimize(fun, x0, jac=True, method="BFGS", options={"gtol": gtol, "maxiter": maxiter})


@dataclass
class NativeBFGSResult:
    """What the library's BFGS loop (include/sfhcuda.h: sfh_minimize_bfgs / sfh_fit_*_bfgs) returns, with scipy's field names."""
    x: np.ndarray
    hess_inv: np.ndarray
    fun: float
    g_norm: float
    nit: int
    nfev: int
    success: bool
    status: int          # 0 converged; 1 iteration limit; 2 line search failed; 3 start not finite


def _bfgs_opts(gtol, maxiter, alphaguess=0, device_hessian=False):
    o = L.sfh_bfgs_opts()
    o.struct_size = C.sizeof(L.sfh_bfgs_opts)
    o.g_abstol, o.maxiter, o.alphaguess, o.device_hessian = float(gtol), int(maxiter), int(alphaguess), int(bool(device_hessian))
    return o


def _native_result(x, invH, rep):
    return NativeBFGSResult(x, invH, rep.f, rep.g_norm, int(rep.iterations), int(rep.f_calls), bool(rep.converged), int(rep.status))


def native_bfgs(fun, x0, gtol=1e-8, maxiter=5000, alphaguess=0):
    """The library's BFGS loop on an arbitrary Python objective ``fun(x) -> (f, grad)`` (sfh_minimize_bfgs through a ctypes
    callback).  The drivers below bind it to the device objectives natively (no callback, one C call per optimisation);
    this generic form exists for user-defined objectives and for testing the engine itself."""
    x = np.array(x0, dtype=np.float64)
    n = x.shape[0]
    err = []

    def cb(_user, xp, nn, fp, gp):
        try:
            f, g = fun(np.ctypeslib.as_array(xp, shape=(nn,)).copy())
            fp[0] = float(f)
            np.ctypeslib.as_array(gp, shape=(nn,))[:] = g
            return 0
        except Exception as e:          # an exception must not unwind through the C frames
            err.append(e)
            return L.SFH_ERR_INVALID_ARG
    invH = np.empty((n, n), order="F")
    rep = L.sfh_bfgs_report()
    o = _bfgs_opts(gtol, maxiter, alphaguess)
    st = L.lib.sfh_minimize_bfgs(L.sfh_objective_fn(cb), None, n, x.ctypes.data_as(C.POINTER(C.c_double)), C.byref(o), C.byref(rep),
                                 invH.ctypes.data_as(C.POINTER(C.c_double)))
    if err:
        raise err[0]
    L.check(st)
    return _native_result(x, invH, rep)


def native_fit_sfh_generic(inner_fg, n_ages, params0, transforms, free, xstart, jacobian_corrections=True, gtol=1e-8, maxiter=5000):
    """fit_sfh's transformed objective and BFGS loop (sfh_fit_sfh_bfgs_generic) around a caller-supplied hierarchical
    ``inner_fg(variables) -> (-logL, gradient)`` over the natural variables [R_1..R_n_ages, model parameters] -- the route for
    user-defined metallicity / dispersion models, whose chain rule is host code (SURVEY.md section 8b "GENERIC")."""
    x = np.array(xstart, dtype=np.float64)
    p0 = np.ascontiguousarray(params0, dtype=np.float64)
    tf = np.ascontiguousarray(transforms, dtype=np.int32)
    fr = np.ascontiguousarray(free, dtype=np.uint8)
    if not (p0.shape == tf.shape == fr.shape) or x.shape[0] != n_ages + int(fr.sum()):
        raise ValueError("length(x0) != n_ages + number of free parameters")
    err = []

    def cb(_user, xp, nn, fp, gp):
        try:
            f, g = inner_fg(np.ctypeslib.as_array(xp, shape=(nn,)).copy())
            fp[0] = float(f)
            np.ctypeslib.as_array(gp, shape=(nn,))[:] = g
            return 0
        except Exception as e:
            err.append(e)
            return L.SFH_ERR_INVALID_ARG
    invH = np.empty((x.shape[0],) * 2, order="F")
    rep, o, dp = L.sfh_bfgs_report(), _bfgs_opts(gtol, maxiter), C.POINTER(C.c_double)
    st = L.lib.sfh_fit_sfh_bfgs_generic(L.sfh_objective_fn(cb), None, int(n_ages), p0.shape[0], p0.ctypes.data_as(dp),
                                        tf.ctypes.data_as(C.POINTER(C.c_int32)), fr.ctypes.data_as(C.POINTER(C.c_uint8)),
                                        int(bool(jacobian_corrections)), x.ctypes.data_as(dp), C.byref(o), C.byref(rep), invH.ctypes.data_as(dp))
    if err:
        raise err[0]
    L.check(st)
    return _native_result(x, invH, rep)


def _native_fit_templates(ds, transform, theta0, gtol, maxiter, device_hessian=False):
    theta = np.array(theta0, dtype=np.float64)
    n = theta.shape[0]
    invH = np.empty((n, n), order="F")
    rep = L.sfh_bfgs_report()
    o = _bfgs_opts(gtol, maxiter, 0, device_hessian)
    dp = C.POINTER(C.c_double)
    L.check(L.lib.sfh_fit_templates_bfgs(ds.ctx().handle, transform, theta.ctypes.data_as(dp), C.byref(o), C.byref(rep), invH.ctypes.data_as(dp)))
    return _native_result(theta, invH, rep)


def _check_engine(engine):
    if engine not in ("scipy", "native"):
        raise ValueError("engine must be 'scipy' or 'native'")


def fit_templates(models, data, x0=None, g_abstol=1e-8, iterations=5000, engine="scipy", device_hessian=False):
    """Returns {"map": LogTransformFTResult, "mle": ...}: BFGS on log-coefficients (solvers.jl:172-221).
    engine="native": the whole optimisation is one call into the library (sfh_fit_templates_bfgs)."""
    _check_engine(engine)
    ds = device_stack(models, data)
    x0 = np.ones(ds.shape[1]) if x0 is None else np.asarray(x0, dtype=np.float64)
    _check_sizes(x0, ds)
    x0 = np.log(renormalize_x0(data, ds, x0))                              # :175-176
    G = np.empty(ds.shape[1])

    def fg_map(logx):                                                      # :178-186
        x = np.exp(logx)
        f = float(_fg_flat(True, G, x, ds, data)) - logx.sum()
        return f, G * x - 1

    def fg_mle(logx):                                                      # :187-195
        x = np.exp(logx)
        f = float(_fg_flat(True, G, x, ds, data))
        return f, G * x

    if engine == "native":
        # device_hessian (experimental): the T x T inverse Hessian lives in HBM, its update and products run as kernels
        rmap = _native_fit_templates(ds, L.SFH_FIT_LOG_MAP, x0, g_abstol, iterations, device_hessian)
        rmle = _native_fit_templates(ds, L.SFH_FIT_LOG_MLE, rmap.x, g_abstol, iterations, device_hessian)
    else:
        rmap = _bfgs(fg_map, x0, g_abstol, iterations)                     # :206
        rmle = _bfgs(fg_mle, rmap.x, g_abstol, iterations)                 # :207 (seeded from the MAP)
    out = {}
    for key, r in (("map", rmap), ("mle", rmle)):
        mu = np.exp(r.x)
        out[key] = LogTransformFTResult(mu, np.sqrt(np.abs(np.diag(r.hess_inv))) * mu, np.asarray(r.hess_inv), r)
    return out


def fit_templates_fast(models, data, x0=None, g_abstol=1e-8, iterations=5000, engine="scipy"):
    """Returns (coeffs, result): BFGS on theta with coeffs = theta^2  (solvers.jl:248-275).
    engine="native": one call into the library (sfh_fit_templates_bfgs, SFH_FIT_SQRT_MLE)."""
    _check_engine(engine)
    ds = device_stack(models, data)
    x0 = np.ones(ds.shape[1]) if x0 is None else np.asarray(x0, dtype=np.float64)
    _check_sizes(x0, ds)
    x0 = np.sqrt(renormalize_x0(data, ds, x0))                             # :251-252
    G = np.empty(ds.shape[1])

    def fg_mle(sqrtx):                                                     # :254-261
        f = float(_fg_flat(True, G, sqrtx ** 2, ds, data))
        return f, G * 2 * sqrtx

    r = _native_fit_templates(ds, L.SFH_FIT_SQRT_MLE, x0, g_abstol, iterations) if engine == "native" else _bfgs(fg_mle, x0, g_abstol, iterations)
    return r.x ** 2, r


# ---------------------------------------------------------------------------------------------
@dataclass
class BFGSResult:
    """hierarchical/bfgs_result.jl:25-41: mu / sigma in natural units, invH in the transformed fitting space."""
    mu: np.ndarray
    sigma: np.ndarray
    invH: np.ndarray
    result: object
    MH_model: object = None
    disp_model: object = None

    def __len__(self):          # bfgs_result.jl:38
        return self.mu.shape[0]

    def mode(self):             # :39
        return self.mu

    def median(self):           # :40
        return self.mu

    def std(self):              # :41
        return self.sigma

    def rand(self, rng, n, invH=None):
        """`rand(result, N)` (bfgs_result.jl:43-75): draws from MvNormal(minimizer, invH) in the fitting space, transformed
        back to natural units; fixed parameters are written in at their values.  Returns (length(mu), n)."""
        npar = self.MH_model.nparams() + self.disp_model.nparams()
        nj = self.mu.shape[0] - npar
        tf = np.array(list(self.MH_model.transforms()) + list(self.disp_model.transforms()))
        free = np.array(list(self.MH_model.free_params()) + list(self.disp_model.free_params()), dtype=bool)
        cov = np.asarray(self.invH if invH is None else invH, dtype=np.float64)
        z = rng.multivariate_normal(np.asarray(self.result.x, dtype=np.float64), (cov + cov.T) / 2, size=n).T
        out = np.empty((self.mu.shape[0], n))
        out[:nj] = np.exp(z[:nj])                                          # transformations.jl:70-72
        tfree = tf[free]
        zp = z[nj:]
        vals = np.where(tfree[:, None] == 1, np.exp(zp), np.where(tfree[:, None] == -1, -np.exp(zp), zp))   # :73-90
        out[nj:][free] = vals
        par = np.array(list(self.MH_model.fittable_params()) + list(self.disp_model.fittable_params()), dtype=np.float64)
        out[nj:][~free] = par[~free][:, None]
        return out


def rand_result(result, rng, n):
    """`rand` for either a BFGSResult or the {"map", "mle"} pair fit_sfh returns (CompositeBFGSResult,
    bfgs_result.jl:114-150: MLE best-fit values with the MAP inverse Hessian)."""
    if isinstance(result, dict):
        return result["mle"].rand(rng, n, invH=result["map"].invH)
    return result.rand(rng, n)


def result_mode(result):
    """`mode` / `median` of a BFGSResult or of the {"map", "mle"} pair (CompositeBFGSResult: the MLE, bfgs_result.jl:39-40,109-110)."""
    return (result["mle"] if isinstance(result, dict) else result).mu


result_median = result_mode


def result_std(result):
    """`std` (bfgs_result.jl:41,111): the standard errors -- of the MAP for the {"map", "mle"} pair, whose inverse Hessian is the
    better conditioned one."""
    return (result["map"] if isinstance(result, dict) else result).sigma


def construct_x0(logAge, T_max, normalize_value=1.0):
    """construct_x0 (fitting/utilities.jl:31-47): starting coefficients of a constant star-formation rate whose total
    is `normalize_value`, `logAge` being left bin edges and `T_max` [Gyr] the final right edge; order-independent."""
    la = np.asarray(logAge, dtype=np.float64)
    max_logAge = math.log10(T_max) + 9
    if not max_logAge > la.max():
        raise ValueError("log10(T_max) + 9 > maximum(logAge) must hold")                # :33
    ua, inv, cnt = np.unique(la, return_inverse=True, return_counts=True)
    edges = np.concatenate([10.0 ** ua, [10.0 ** max_logAge]])
    sfr = normalize_value / (10.0 ** max_logAge - 10.0 ** la.min())
    return sfr * np.diff(edges)[inv] / cnt[inv]


def _unique_first(a):
    """unique(a) in first-appearance order (Julia's `unique`)."""
    a = np.asarray(a, dtype=np.float64)
    _, first = np.unique(a, return_index=True)
    return a[np.sort(first)]


def construct_x0_mdf(logAge, *args, normalize_value=1.0):
    """construct_x0_mdf (hierarchical/construct_x0_mdf.jl:58-117): one starting stellar mass per unique(logAge), in that order.
    `construct_x0_mdf(logAge, T_max)`: constant star-formation rate; `construct_x0_mdf(logAge, cum_sfh, T_max)`: from a
    cumulative SFH given on unique(logAge), or as a pair (logAge grid, cumulative SFH) that is interpolated onto it.
    (The reference's index arithmetic -- the position in unique(logAge) of the i-th SORTED age indexes the SORTED widths,
    :67-70 -- is restated as written: it is the intended pairing for ascending or descending ages.)"""
    la = np.asarray(logAge, dtype=np.float64)
    if len(args) == 1:
        cum, T_max = None, args[0]
    elif len(args) == 2:
        cum, T_max = args
    else:
        raise TypeError("construct_x0_mdf(logAge, [cum_sfh,] T_max)")
    max_logAge = math.log10(T_max) + 9
    if not max_logAge > la.max():
        raise ValueError("log10(T_max) + 9 > maximum(logAge) must hold")                # :59, :79
    ua = _unique_first(la)
    order = np.argsort(ua, kind="stable")
    sorted_ul = np.concatenate([ua[order], [max_logAge]])
    pos = {float(v): k for k, v in enumerate(ua)}
    idx = np.array([pos[float(sorted_ul[i])] for i in range(ua.shape[0])])
    if cum is None:
        sfr = normalize_value / (10.0 ** max_logAge - 10.0 ** la.min())
        return sfr * np.diff(10.0 ** sorted_ul)[idx]
    if len(cum) == 2 and np.ndim(cum[0]) == 1 and np.ndim(cum[1]) == 1 and len(cum[0]) == len(cum[1]) and len(cum[0]) != 1 \
            and not np.isscalar(cum[0]):
        gla, gcs = np.asarray(cum[0], dtype=np.float64), np.asarray(cum[1], dtype=np.float64)
        if gcs.max() > 1:
            raise ValueError("Maximum of cumulative SFH must be less than or equal to one")   # :106
        o = np.argsort(gla, kind="stable")
        cum = np.interp(ua, np.concatenate([gla[o], [max_logAge]]), np.concatenate([gcs[o], [0.0]]))   # Flat() extrapolation
    cum = np.asarray(cum, dtype=np.float64)
    if cum.min() < 0:
        raise ValueError("minimum(cum_sfh) >= 0 must hold")                              # :75
    if cum.shape[0] != ua.shape[0]:
        raise ValueError("`length(unique(logAge))` not equal to `length(cum_sfh)`.")     # :82
    sc = np.concatenate([cum[order], [0.0]])
    if np.any(np.diff(sc) > 0):
        raise ValueError("Provided `cum_sfh` must be monotonically increasing as `logAge` decreases.")   # :87
    return normalize_value * (sc[idx] - sc[idx + 1])


def truncate_relweights(relweightsmin, relweights, logAge):
    """truncate_relweights (hierarchical/fixed_amr.jl:229-243): indices (0-based) of the templates whose relative weight is
    at least `relweightsmin` times the largest one of their age, grouped by unique(logAge) in first-appearance order."""
    rw, la = np.asarray(relweights, dtype=np.float64), np.asarray(logAge, dtype=np.float64)
    if rw.shape != la.shape:
        raise ValueError("length(relweights) == length(logAge) must hold")
    if relweightsmin == 0:
        return np.arange(rw.shape[0])
    keep = []
    for a in _unique_first(la):
        good = np.nonzero(la == a)[0]
        keep.append(good[rw[good] >= relweightsmin * rw[good].max()])
    return np.concatenate(keep)


def fixed_amr(models, data, logAge, metallicities, relweights, relweightsmin=0, x0=None, g_abstol=1e-8, iterations=5000,
              engine="scipy"):
    """fixed_amr (hierarchical/fixed_amr.jl:41-180): one stellar-mass coefficient per unique(logAge) under externally imposed
    relative weights; BFGS on log-coefficients, MAP (Jacobian term) then MLE seeded from it.  Every objective evaluation is
    one fused `fg!` on the device; the per-age contraction of the gradient (:119-121, :147-151) is O(T) on the host.
    Returns {"map": ..., "mle": ...} with mu, sigma (sqrt of diag(invH), log space, :173-174), invH, result."""
    la = np.asarray(logAge, dtype=np.float64)
    mh = np.asarray(metallicities, dtype=np.float64)
    rw = np.array(relweights, dtype=np.float64)
    ua = _unique_first(la)
    if x0 is None:
        x0 = construct_x0_mdf(la, 13.7)
    x0 = np.asarray(x0, dtype=np.float64)
    if x0.shape[0] != ua.shape[0]:
        raise ValueError("length(x0) == length(unique(logAge)) must hold")               # :50
    if not (la.shape == mh.shape == rw.shape):
        raise ValueError("size(models,2) == length(logAge) == length(metallicities) == length(relweights) must hold")
    if np.any(rw < 0):
        raise ValueError("all relative weights must be >= 0")                            # :53
    if relweightsmin < 0:
        raise ValueError("relweightsmin >= 0 must hold")                                 # :54
    if relweightsmin != 0:                                                              # :58-64: a narrower stack
        keep = truncate_relweights(relweightsmin, rw, la)
        if isinstance(models, DeviceStack):
            Mh, _ = models.download()
            data = models.download_data() if data is None else data
        else:
            Mh = np.asarray(models) if not isinstance(models, (list, tuple)) else np.stack([np.asarray(m).reshape(-1, order="F") for m in models], axis=1)
        models = np.asfortranarray(Mh[:, keep])
        data = np.asarray(data).reshape(-1, order="F")
        la, mh, rw = la[keep], mh[keep], rw[keep]
    ds = device_stack(models, data)
    if ds.shape[1] != la.shape[0]:
        raise ValueError("size(models,2) == length(logAge) must hold")                   # :52
    inv = np.array([int(np.nonzero(ua == a)[0][0]) for a in la])                        # template -> age (idxlogAge, :80)
    sums = np.bincount(inv, weights=rw, minlength=ua.shape[0])
    if not np.allclose(sums, 1.0):                                                      # :66-77
        if relweightsmin == 0:
            import warnings
            warnings.warn("The relative weights provided to `fixed_amr` do not sum to 1 for every logAge and will be renormalized.")
        rw = rw / sums[inv]
    x0 = renormalize_x0(data, ds, x0, rw * x0[inv])                                      # :83-88

    def make(jac):
        def fun(xvec):
            x = np.exp(xvec)
            coeffs = rw * x[inv]                                                         # :106-108
            nl, G, _ = ds.eval_fg(coeffs)                                                # -logL and +M'r = -fullG
            g = np.bincount(inv, weights=G * coeffs, minlength=ua.shape[0])              # -sum(fullG[j] coeffs[j])  (:120, :150)
            if jac:
                return nl - xvec.sum(), g - 1                                            # :112, :120
            return nl, g
        return fun
    _check_engine(engine)
    res = {}
    start = np.log(x0)
    for key, jac in (("map", True), ("mle", False)):                                     # :166-167
        if engine == "native":                                                          # one library call per optimisation
            theta = np.array(start, dtype=np.float64)
            invH = np.empty((theta.shape[0],) * 2, order="F")
            rep, o, dp = L.sfh_bfgs_report(), _bfgs_opts(g_abstol, iterations), C.POINTER(C.c_double)
            rwc, inv32 = np.ascontiguousarray(rw, dtype=np.float64), np.ascontiguousarray(inv, dtype=np.int32)
            L.check(L.lib.sfh_fit_fixed_amr_bfgs(ds.ctx().handle, rwc.ctypes.data_as(dp), inv32.ctypes.data_as(C.POINTER(C.c_int32)),
                                                 ua.shape[0], int(jac), theta.ctypes.data_as(dp), C.byref(o), C.byref(rep),
                                                 invH.ctypes.data_as(dp)))
            r = _native_result(theta, invH, rep)
        else:
            r = _bfgs(make(jac), start, g_abstol, iterations)
        start = r.x
        invH = np.asarray(r.hess_inv)
        res[key] = {"mu": np.exp(r.x), "sigma": np.sqrt(np.abs(np.diag(invH))), "invH": invH, "result": r}
    return res


def calculate_cum_sfr(coeffs, logAge, MH, T_max, normalize_value=1, sorted=False):
    """calculate_cum_sfr (fitting/utilities.jl:153-195): (unique_logAge ascending, cumulative SFH normalised to 1 at the
    youngest bin, SFR per bin with `logAge` as left edges and `T_max` [Gyr] the last right edge, mass-weighted <[M/H]>)."""
    c = np.asarray(coeffs, dtype=np.float64) * normalize_value
    la = np.asarray(logAge, dtype=np.float64)
    mh = np.asarray(MH, dtype=np.float64)
    if not (c.shape == la.shape == mh.shape):
        raise ValueError("axes(coeffs) == axes(logAge) == axes(MH) must hold")          # :154
    max_logAge = math.log10(T_max) + 9
    if not max_logAge > la.max():
        raise ValueError("log10(T_max) + 9 > maximum(logAge) must hold")                # :155
    mtot = c.sum()
    if not sorted:
        idx = np.argsort(la, kind="stable")
        la, c, mh = la[idx], c[idx], mh[idx]
    ua, inv = np.unique(la, return_inverse=True)
    dt = np.diff(np.concatenate([10.0 ** ua, [10.0 ** max_logAge]]))
    mstar = np.bincount(inv, weights=c, minlength=ua.shape[0])
    wsum = np.bincount(inv, weights=c * mh, minlength=ua.shape[0])
    cnt = np.bincount(inv, minlength=ua.shape[0])
    plain = np.bincount(inv, weights=mh, minlength=ua.shape[0]) / cnt
    mean_mh = np.empty(ua.shape[0])
    for i in range(ua.shape[0]):                                                      # :176-188
        if mstar[i] == 0:
            mean_mh[i] = plain[i] if i == 0 else mean_mh[i - 1]
        else:
            mean_mh[i] = wsum[i] / mstar[i]
    cum = np.cumsum(mstar[::-1])[::-1] / mtot
    return ua, cum, mstar / dt, mean_mh


def cum_sfr_quantiles(result, logAge, MH, T_max, Nsamples, q, rng=None, **kws):
    """cum_sfr_quantiles (fitting/utilities.jl:239-302): draw `Nsamples` SFHs from a fit_sfh result, expand each with
    calculate_coeffs, and return per-age quantiles `q` of the cumulative SFH, the SFR and <[M/H]> plus the draws.
    Samples whose coefficients are not all finite are dropped, as in the reference (:267-270)."""
    rng = np.random.default_rng() if rng is None else rng
    best = result["mle"] if isinstance(result, dict) else result
    samples = rand_result(result, rng, int(Nsamples))
    npm, npd = best.MH_model.nparams(), best.disp_model.nparams()
    nj = samples.shape[0] - npm - npd
    la, mh = np.asarray(logAge, dtype=np.float64), np.asarray(MH, dtype=np.float64)
    rows = []
    for i in range(samples.shape[1]):
        r = samples[:, i]
        mm = best.MH_model.update_params(r[nj:nj + npm])
        dm = best.disp_model.update_params(r[nj + npm:])
        with np.errstate(all="ignore"):
            co = calculate_coeffs(mm, dm, r[:nj], la, mh)
        if not np.all(np.isfinite(co)):
            continue
        rows.append(calculate_cum_sfr(co, la, mh, T_max, **kws)[1:])
    qq = np.atleast_1d(np.asarray(q, dtype=np.float64))
    stack = [np.array([row[k] for row in rows]) for k in range(3)]        # each (ngood, Nj)
    cum_q, sfr_q, mh_q = (np.quantile(a, qq, axis=0).T for a in stack)
    return {"cum_sfh": cum_q, "sfrs": sfr_q, "mean_mh": mh_q, "samples": samples, "n_good": len(rows)}


def tau_interp(unique_logAge, max_logAge, cum_sfh):
    """tau_interp (fitting/utilities.jl:311-336): piecewise-linear map from the fraction of the total stellar mass formed to the
    lookback time [Gyr].  Knots: the cumulative SFH normalised to its maximum, with (0, max_logAge) prepended -- the moment the
    galaxy had no mass; either ordering of the inputs is accepted; equal knots are separated by one ulp each
    (`deduplicate_knots!(...; move_knots=true)`).  Like the reference's gridded interpolant the callable refuses fractions
    outside [0, 1] (ValueError for Julia's BoundsError) instead of clamping."""
    la, cum = np.asarray(unique_logAge, dtype=np.float64), np.asarray(cum_sfh, dtype=np.float64)
    if la.ndim != 1 or la.shape != cum.shape:
        raise ValueError("length(unique_logAge) != length(cum_sfh)")                                       # :312
    if not la.max() < max_logAge:
        raise ValueError("`max_logAge` must be greater than the maximum of `unique_logAge`.")              # :313
    ascending = lambda a: bool(np.all(a[1:] >= a[:-1]))
    if not ascending(cum):                                                                                 # :314-317
        cum, la = cum[::-1], la[::-1]
    if not ascending(cum):
        raise ValueError("`cum_sfh` must be sorted in ascending or descending order.")                     # :319
    if not ascending(-la):
        raise ValueError("`unique_logAge` must be sorted in the same order as `cum_sfh`.")                 # :320
    if cum[0] != 0:
        cum = np.concatenate([[0.0], cum])                                                                 # :322-324
    if la[0] != max_logAge:
        la = np.concatenate([[max_logAge], la])                                                            # :326-328
    if cum.shape != la.shape:
        raise ValueError("knots and lookback times differ in length (a cum_sfh that starts at 0 must come with max_logAge)")
    knots = cum / cum.max()                                                                                # :330
    for i in range(1, knots.shape[0]):                                                                     # :332
        if knots[i] <= knots[i - 1]:
            knots[i] = np.nextafter(knots[i - 1], np.inf)
    t = 10.0 ** la / 1e9                                                                                   # :333

    def itp(frac):
        f = np.asarray(frac, dtype=np.float64)
        if np.any(f < knots[0]) or np.any(f > knots[-1]) or np.any(np.isnan(f)):
            raise ValueError(f"tau = {frac} outside the interpolation range [{knots[0]}, {knots[-1]}]")
        out = np.interp(f, knots, t)
        return float(out) if out.ndim == 0 else out

    itp.knots, itp.values = knots, t
    return itp


def tau(*args, Nsamples=10_000, q=(0.16, 0.5, 0.84), rng=None, **kws):
    """tau (fitting/utilities.jl:416-468), the reference's three methods:

    ``tau(τ, unique_logAge, max_logAge, cum_sfh)``                 lookback time [Gyr] at which the fraction τ of the mass had formed
    ``tau(τ, unique_logAge, max_logAge, cum_sfh, lower, upper)``   (length(τ), 3): columns lower / best / upper                   :420-430
    ``tau(result, τ, logAge, MH, max_logAge; Nsamples, q, kws...)`` the same from `cum_sfr_quantiles` of a fit_sfh result          :456-468
    """
    if len(args) == 4:
        frac, ula, max_logAge, cum = args
        return tau_interp(ula, max_logAge, cum)(frac)
    if len(args) == 6:
        frac, ula, max_logAge, cum, lower, upper = args
        cols = [np.atleast_1d(tau_interp(ula, max_logAge, c)(frac)) for c in (lower, cum, upper)]
        return np.stack(cols, axis=1)
    if len(args) == 5:
        result, frac, logAge, MH, max_logAge = args
        if len(q) != 3:
            raise ValueError("q must be a tuple of three quantiles (e.g., (0.16, 0.5, 0.84))")            # :457
        out = cum_sfr_quantiles(result, logAge, MH, 10.0 ** max_logAge / 1e9, Nsamples, q, rng=rng, **kws)
        cm = out["cum_sfh"]
        return tau(frac, np.unique(np.asarray(logAge, dtype=np.float64)), max_logAge, cm[:, 1], cm[:, 0], cm[:, 2])
    raise TypeError("tau takes (τ, unique_logAge, max_logAge, cum_sfh[, lower, upper]) or (result, τ, logAge, MH, max_logAge)")


def fit_sfh(MH_model0, disp_model0, models, data, logAge, metallicities, x0=None, g_abstol=1e-8, iterations=5000, engine="scipy",
            alphaguess=0):
    """BFGS on [log R_j, transformed free parameters]: MAP (Jacobian corrections on) then MLE seeded from it
    (generic_fitting.jl:242-409).  Returns {"map": BFGSResult, "mle": BFGSResult}; mu holds
    [R_1..R_Nj, alpha, beta, sigma] with fixed parameters at their initial values."""
    _check_engine(engine)
    ds = device_stack(models, data)
    la, mh = np.asarray(logAge, float), np.asarray(metallicities, float)
    _, first = np.unique(la, return_index=True)
    nj = first.shape[0]
    tf = np.array(list(MH_model0.transforms()) + list(disp_model0.transforms()))
    free = np.array(list(MH_model0.free_params()) + list(disp_model0.free_params()), dtype=bool)
    par0 = np.array(list(MH_model0.fittable_params()) + list(disp_model0.fittable_params()), dtype=np.float64)
    if x0 is None:
        x0 = np.ones(nj)
    x0 = np.asarray(x0, dtype=np.float64)
    if x0.shape[0] != nj:
        raise ValueError("length(x0) != length(unique(logAge))")
    full = calculate_coeffs(MH_model0, disp_model0, x0, la, mh, models=ds)  # :260 (device prologue kernel)
    if np.isnan(full).any():
        raise ValueError("initial metallicity-model parameters give NaN coefficients (generic_fitting.jl:261-283)")
    x0 = renormalize_x0(data, ds, x0, full)                                # :283
    xstart = np.concatenate([np.log(x0), logtransform(par0, tf)[free]])    # :285-294
    res = {}
    start = xstart
    for key, jac in (("map", True), ("mle", False)):                       # :304-327
        opt = HierarchicalOptimizer(MH_model0, disp_model0, ds, data, la, mh, True, True, jac)

        def fun(X, opt=opt):
            lp, g = opt.logdensity_and_gradient(X)
            return -lp, -g

        if engine == "native":                                             # the loop runs inside the library (sfh_fit_sfh_bfgs)
            r = opt.native_bfgs(start, g_abstol, iterations, alphaguess)   # alphaguess: sfh_bfgs_opts (0/1 = the reference's InitialStatic)
        else:
            r = _bfgs(fun, start, g_abstol, iterations)
        start = r.x                                                        # MLE starts from the MAP minimiser
        mu = np.empty(nj + tf.shape[0])
        mu[:nj] = np.exp(r.x[:nj])
        mu[nj:][free] = exptransform(r.x[nj:], tf[free])
        mu[nj:][~free] = par0[~free]
        # sigma: delta method through the log transforms (generic_fitting.jl:352-407)
        sd = np.sqrt(np.abs(np.diag(r.hess_inv)))
        sigma = np.zeros_like(mu)
        sigma[:nj] = sd[:nj] * mu[:nj]
        sfree = sd[nj:]
        tfree = tf[free]
        sigma[nj:][free] = np.where(tfree == 0, sfree, sfree * np.abs(mu[nj:][free]))
        res[key] = BFGSResult(mu, sigma, np.asarray(r.hess_inv), r,
                              MH_model0.update_params(mu[nj:nj + 2]), disp_model0.update_params(mu[nj + 2:]))
    return res


# ---------------------------------------------------------------------------------------------
def stretch_move_ensemble(logp_batch, x0, nsteps, a_scale=2.0, rng=None, thin=1):
    """Goodman & Weare affine-invariant ensemble sampler ("emcee", KissMCMC.emcee's algorithm) with the two
    half-ensembles updated alternately, each half by ONE call of `logp_batch(X)` (X: (npar, W/2)).
    x0: (npar, nwalkers).  Returns (chain (nsteps//thin, npar, nwalkers), logp, acceptance fraction)."""
    rng = np.random.default_rng() if rng is None else rng
    X = np.array(x0, dtype=np.float64, order="F")
    npar, W = X.shape
    if W % 2 or W < 2:
        raise ValueError("need an even number of walkers")
    half = W // 2
    lp = logp_batch(X)
    chain = np.empty((nsteps // thin, npar, W))
    lps = np.empty((nsteps // thin, W))
    acc = 0
    for step in range(nsteps):
        for h in (0, 1):
            act = slice(0, half) if h == 0 else slice(half, W)
            oth = slice(half, W) if h == 0 else slice(0, half)
            z = ((a_scale - 1.0) * rng.random(half) + 1.0) ** 2 / a_scale  # g(z) ~ 1/sqrt(z) on [1/a, a]
            partner = X[:, oth][:, rng.integers(0, half, size=half)]
            prop = partner + z[None, :] * (X[:, act] - partner)
            lpp = logp_batch(np.asfortranarray(prop))
            with np.errstate(invalid="ignore"):
                lnr = (npar - 1) * np.log(z) + lpp - lp[act]               # acceptance of the stretch move
            ok = np.log(rng.random(half)) < lnr
            ok &= np.isfinite(lpp)
            Xa = X[:, act]
            Xa[:, ok] = prop[:, ok]
            X[:, act] = Xa
            la = lp[act]
            la[ok] = lpp[ok]
            lp[act] = la
            acc += int(ok.sum())
        if (step + 1) % thin == 0:
            chain[(step + 1) // thin - 1] = X
            lps[(step + 1) // thin - 1] = lp
    return chain, lps, acc / (nsteps * W)


def mcmc_sample(models, data, x0, nsteps, nburnin=0, nthin=1, a_scale=2.0, rng=None, engine="device"):
    """mcmc_sample(models, data, x0, nwalkers-implied, nsteps; nburnin, nthin, a_scale)  (mcmc_sample.jl:97-108).
    x0: (npar, nwalkers) or list of walker vectors.  Returns samples with shape (nsteps, npar, nwalkers) like
    convert_kissmcmc (:30-44), the log-likelihoods and the acceptance fraction.
    engine="device": the whole ensemble sampler runs on the GPU (sfh_mcmc_run: proposal, K6, accept/reject; Philox
    streams seeded from `rng`); engine="host": numpy proposals around one K6 call per half-ensemble."""
    ds = device_stack(models, data)
    X0 = np.asarray(x0, dtype=np.float64)
    if X0.ndim == 2 and X0.shape[0] != ds.shape[1] and X0.shape[1] == ds.shape[1]:
        X0 = X0.T                                                          # list of walker vectors
    if X0.shape[0] != ds.shape[1]:
        raise ValueError("length of each walker != number of templates")
    if engine == "host":
        model = MCMCModel(ds, data)
        chain, lps, acc = stretch_move_ensemble(model.batch, X0, nsteps + nburnin, a_scale, rng, 1)
        return chain[nburnin::nthin], lps[nburnin::nthin], acc
    if engine != "device":
        raise ValueError("engine must be 'device' or 'host'")
    if X0.shape[1] % 2 or X0.shape[1] < 2:
        raise ValueError("need an even number of walkers")
    rng = np.random.default_rng() if rng is None else rng
    seeds = rng.integers(0, 2**63, size=2)
    acc_b = 0.0
    if nburnin > 0:                                                        # burn-in: nothing stored, nothing copied back
        _, _, X0, _, acc_b = ds.mcmc_run(X0, nburnin, 1, a_scale, int(seeds[0]), store=False)
    chain, lps, _, _, acc = ds.mcmc_run(X0, nsteps, nthin, a_scale, int(seeds[1]), store=True)
    tot = nburnin + nsteps
    return chain, lps, (acc_b * nburnin + acc * nsteps) / tot if tot else 0.0


# ---------------------------------------------------------------------------------------------
def mdf_amr(coeffs, logAge, metallicities, models=None):
    """mdf_amr (src/fitting/mdf.jl:23-37 mass-weighted; :54-74 number-weighted with `models`).  Returns
    (unique_MH sorted ascending, mdf).  With `models` (a DeviceStack) the per-metallicity composite sums
    sum_i (M[:, idxs] * coeffs[idxs])_i are formed from the resident stack's column sums (one device pass)."""
    coeffs = np.asarray(coeffs, dtype=np.float64)
    mh = np.asarray(metallicities, dtype=np.float64)
    if not (coeffs.shape[0] == np.asarray(logAge).shape[0] == mh.shape[0]):
        raise ValueError("length(coeffs) == length(logAge) == length(metallicities) violated")     # mdf.jl:27 / :59
    _, first = np.unique(mh, return_index=True)
    umh = mh[np.sort(first)]                                                # unique() in first-appearance order
    if models is None:
        w = coeffs
    else:
        ds = device_stack(models, None)
        if ds.shape[1] != coeffs.shape[0]:
            raise ValueError("length(coeffs) != size(models, 2)")            # mdf.jl:59
        w = coeffs * ds.column_sums()                                       # sum(mul!(composite, M[:,idxs], c[idxs]))  :68-69
    mdf = np.array([w[mh == m].sum() for m in umh])
    if models is None:
        mdf = mdf / mdf.sum()                                               # :34 (only the mass-weighted form normalises)
    p = np.argsort(umh, kind="stable")                                      # :35 / :71
    return umh[p], mdf[p]


# ---------------------------------------------------------------------------------------------
def _vel(inv_mass, r):
    """M^-1 r for a diagonal (vector) or dense (matrix) inverse mass."""
    return inv_mass * r if inv_mass.ndim == 1 else inv_mass @ r


def _leapfrog(theta, r, grad, eps, inv_mass):
    """One leapfrog step as a coroutine: yields the position whose (logp, gradient) it needs."""
    r = r + 0.5 * eps * grad
    theta = theta + eps * _vel(inv_mass, r)
    lp, grad = yield theta
    r = r + 0.5 * eps * grad
    return theta, r, lp, grad


def nuts_chain(theta0, nsteps, nwarmup=200, max_depth=8, delta=0.8, rng=None, inv_mass=None, eps0=None):
    """No-U-Turn sampler (Hoffman & Gelman 2014, algorithm 6: slice NUTS with dual-averaging step size) written as a
    coroutine: it *yields* every position at which it needs the log-density and gradient and is *sent* `(logp, grad)`
    back; its return value is `(samples, logps, step_size)`.  Stands in for DynamicHMC.mcmc_with_warmup
    (hmc_sample.jl:111).  `inv_mass` is M^-1 of the Gaussian kinetic energy: a vector (diagonal) or a dense matrix, e.g. the
    inverse Hessian of a fit (DynamicHMC.GaussianKineticEnergy(MAP.invH), generic_fitting.jl:479-482); it is not adapted.
    `eps0` fixes the initial step size (the reference's ϵ) instead of the doubling heuristic.  `nuts_sample` drives one chain; `run_chains_batched`
    drives many, serving each round of requests with one batched device pass."""
    rng = np.random.default_rng() if rng is None else rng
    theta = np.asarray(theta0, dtype=np.float64).copy()
    d = theta.shape[0]
    inv_mass = np.ones(d) if inv_mass is None else np.asarray(inv_mass, dtype=np.float64)
    if inv_mass.ndim == 2:                                                 # r ~ N(0, M), M = inv_mass^-1 = L^-T L^-1 with inv_mass = L L^T
        Lc = np.linalg.cholesky((inv_mass + inv_mass.T) / 2)
        draw = lambda: np.linalg.solve(Lc.T, rng.standard_normal(d))
    else:
        draw = lambda: rng.standard_normal(d) / np.sqrt(inv_mass)
    kin = lambda r: 0.5 * np.dot(r, _vel(inv_mass, r))
    lp, grad = yield theta

    # heuristic initial step size
    eps = 0.1 / math.sqrt(d)
    r0 = draw()
    _, r1, lp1, _ = yield from _leapfrog(theta, r0, grad, eps, inv_mass)
    H0 = lp - kin(r0); H1 = lp1 - kin(r1)
    a = 1.0 if (np.isfinite(H1) and H1 - H0 > math.log(0.5)) else -1.0
    for _ in range(50 if eps0 is None else 0):
        _, r1, lp1, _ = yield from _leapfrog(theta, r0, grad, eps, inv_mass)
        H1 = lp1 - kin(r1)
        if not np.isfinite(H1):
            H1 = -np.inf
        if a * (H1 - H0) <= -a * math.log(2):
            break
        eps *= 2.0 ** a
    if eps0 is not None:
        eps = float(eps0)
    mu, ebar, Hbar, gamma, t0, kappa = math.log(10 * eps), 1.0, 0.0, 0.05, 10.0, 0.75

    def build(theta, r, grad, logu, v, j, eps, H0):
        if j == 0:
            th, rr, lpn, g = yield from _leapfrog(theta, r, grad, v * eps, inv_mass)
            Hn = lpn - kin(rr)
            if not np.isfinite(Hn):
                Hn = -np.inf
            n = int(logu <= Hn)
            s = int(logu < Hn + 1000.0)
            return th, rr, g, th, rr, g, th, lpn, g, n, s, min(1.0, math.exp(min(0.0, Hn - H0))), 1
        thm, rm, gm, thp, rp, gp, th1, lp1, g1, n1, s1, a1, na1 = yield from build(theta, r, grad, logu, v, j - 1, eps, H0)
        if s1:
            if v == -1:
                thm, rm, gm, _, _, _, th2, lp2, g2, n2, s2, a2, na2 = yield from build(thm, rm, gm, logu, v, j - 1, eps, H0)
            else:
                _, _, _, thp, rp, gp, th2, lp2, g2, n2, s2, a2, na2 = yield from build(thp, rp, gp, logu, v, j - 1, eps, H0)
            if n1 + n2 > 0 and rng.random() < n2 / (n1 + n2):
                th1, lp1, g1 = th2, lp2, g2
            dth = thp - thm
            s1 = s2 * int(np.dot(dth, _vel(inv_mass, rm)) >= 0) * int(np.dot(dth, _vel(inv_mass, rp)) >= 0)
            n1 += n2; a1 += a2; na1 += na2
        return thm, rm, gm, thp, rp, gp, th1, lp1, g1, n1, s1, a1, na1

    samples = np.empty((nsteps, d))
    lps = np.empty(nsteps)
    for m in range(1, nwarmup + nsteps + 1):
        r0 = draw()
        H0 = lp - kin(r0)
        logu = H0 + math.log(rng.random())
        thm = thp = theta; rm = rp = r0; gm = gp = grad
        j, n, s = 0, 1, 1
        alpha, nalpha = 0.0, 1
        while s and j < max_depth:
            v = -1 if rng.random() < 0.5 else 1
            if v == -1:
                thm, rm, gm, _, _, _, th1, lp1, g1, n1, s1, alpha, nalpha = yield from build(thm, rm, gm, logu, v, j, eps, H0)
            else:
                _, _, _, thp, rp, gp, th1, lp1, g1, n1, s1, alpha, nalpha = yield from build(thp, rp, gp, logu, v, j, eps, H0)
            if s1 and rng.random() < min(1.0, n1 / n):
                theta, lp, grad = th1, lp1, g1
            n += n1
            dth = thp - thm
            s = s1 * int(np.dot(dth, _vel(inv_mass, rm)) >= 0) * int(np.dot(dth, _vel(inv_mass, rp)) >= 0)
            j += 1
        if m <= nwarmup:                                                   # dual averaging
            Hbar = (1 - 1 / (m + t0)) * Hbar + (delta - alpha / nalpha) / (m + t0)
            leps = mu - math.sqrt(m) / gamma * Hbar
            eta = m ** (-kappa)
            ebar = math.exp(eta * leps + (1 - eta) * math.log(ebar))
            eps = math.exp(leps)
            if m == nwarmup:
                eps = ebar
        else:
            samples[m - nwarmup - 1] = theta
            lps[m - nwarmup - 1] = lp
    return samples, lps, eps


def nuts_sample(logdensity_and_gradient, theta0, nsteps, nwarmup=200, max_depth=8, delta=0.8, rng=None, inv_mass=None, eps0=None):
    """One NUTS chain: every request of `nuts_chain` is answered by `logdensity_and_gradient(theta) -> (logp, grad)`."""
    chain = nuts_chain(theta0, nsteps, nwarmup, max_depth, delta, rng, inv_mass, eps0)
    try:
        req = next(chain)
        while True:
            req = chain.send(logdensity_and_gradient(req))
    except StopIteration as done:
        return done.value


class ChainStats:
    n_batches = 0
    n_evals = 0


def run_chains_batched(batch_fn, theta0s, nsteps, nwarmup=200, max_depth=8, rngs=None, chain=None):
    """Run len(theta0s) sampler coroutines side by side.  The reference runs HMC chains on separate threads, each
    evaluating its own `fg!` (hmc_sample.jl:123-141, generic_fitting.jl:617-626); here every round collects the pending
    request of each live chain and serves them all with ONE call of
    `batch_fn(Theta[npar, C]) -> (logp[C], grad[npar, C])` (sfh_eval_fg_batched).  A chain's answers do not depend on how
    requests were grouped, so it follows the trajectory it would have followed alone; chains that finish early drop out.
    Returns ([(samples, logps, step_size) per chain], ChainStats)."""
    nchains = len(theta0s)
    rngs = list(np.random.default_rng().spawn(nchains)) if rngs is None else rngs
    chain = nuts_chain if chain is None else chain
    chains = [chain(theta0s[c], nsteps, nwarmup, max_depth, rng=rngs[c]) for c in range(nchains)]
    out, stats = [None] * nchains, ChainStats()
    pending = {}
    for c, ch in enumerate(chains):
        try:
            pending[c] = next(ch)
        except StopIteration as done:
            out[c] = done.value
    while pending:
        ids = sorted(pending)
        Theta = np.empty((pending[ids[0]].shape[0], len(ids)), order="F")
        for k, c in enumerate(ids):
            Theta[:, k] = pending[c]
        lp, gr = batch_fn(Theta)
        stats.n_batches += 1
        stats.n_evals += len(ids)
        for k, c in enumerate(ids):
            try:
                pending[c] = chains[c].send((float(lp[k]), np.array(gr[:, k], dtype=np.float64)))
            except StopIteration as done:
                out[c] = done.value
                del pending[c]
    return out, stats


def _nuts_opts(nwarmup, max_depth, delta, eps0, seed, inv_mass):
    o = L.sfh_nuts_opts()
    o.struct_size = C.sizeof(L.sfh_nuts_opts)
    o.max_depth, o.nwarmup, o.delta = int(max_depth), int(nwarmup), float(delta)
    o.eps0 = 0.0 if eps0 is None else float(eps0)
    o.seed = int(seed) & (2**64 - 1)
    im = None
    if inv_mass is not None:
        im = np.asarray(inv_mass, dtype=np.float64)
        o.mass_kind = 1 if im.ndim == 1 else 2
        im = np.asfortranarray(im)
    return o, im


def _nuts_call(entry, head_args, theta0s, nsteps, nwarmup, max_depth, delta, eps0, seed, inv_mass):
    """Common tail of the three native NUTS entry points: returns ([(samples (nsteps_c, n), logps, step_size) per chain], ChainStats)."""
    Th = np.asfortranarray(np.stack([np.asarray(t, dtype=np.float64) for t in theta0s], axis=1))
    n, nch = Th.shape
    lens = np.ascontiguousarray(np.broadcast_to(np.asarray(nsteps, dtype=np.int64), (nch,)))
    tot = int(lens.sum())
    samples = np.empty((n, max(tot, 1)), order="F")
    logps = np.empty(max(tot, 1))
    steps = np.empty(nch)
    o, im = _nuts_opts(nwarmup, max_depth, delta, eps0, seed, inv_mass)
    if im is not None and im.shape not in ((n,), (n, n)):
        raise ValueError("inv_mass must have shape (n,) or (n, n)")
    nb, ne = C.c_int64(0), C.c_int64(0)
    dp, i64p = C.POINTER(C.c_double), C.POINTER(C.c_int64)
    st = entry(*head_args, nch, Th.ctypes.data_as(dp), lens.ctypes.data_as(i64p), im.ctypes.data_as(dp) if im is not None else None,
               C.byref(o), samples.ctypes.data_as(dp), logps.ctypes.data_as(dp), steps.ctypes.data_as(dp), C.byref(nb), C.byref(ne))
    out, a = [], 0
    for c in range(nch):
        out.append((samples[:, a:a + lens[c]].T.copy(), logps[a:a + lens[c]].copy(), float(steps[c])))
        a += int(lens[c])
    stats = ChainStats()
    stats.n_batches, stats.n_evals = nb.value, ne.value
    return st, out, stats


def native_nuts(batch_fn, theta0s, nsteps, nwarmup=200, max_depth=8, delta=0.8, eps0=None, seed=0, inv_mass=None):
    """The library's multi-chain NUTS (sfh_nuts_run, csrc/sfh_nuts.h) around a Python ``batch_fn(Theta[n, C]) -> (logp[C],
    grad[n, C])``: same algorithm and draw order as :func:`nuts_chain`, Philox streams keyed by ``seed``.  ``nsteps`` may be one
    length or one per chain.  Returns ([(samples, logps, step_size) per chain], ChainStats).  `hmc_sample` / `sample_sfh` /
    `tsample_sfh` with engine="native" bind the device log-densities natively instead (no callback)."""
    err = []

    def cb(_user, thp, n, Cn, lpp, gp):
        try:
            Th = np.ctypeslib.as_array(thp, shape=(Cn, n)).T.copy(order="F")          # n x C column-major
            lp, g = batch_fn(Th)
            np.ctypeslib.as_array(lpp, shape=(Cn,))[:] = lp
            np.ctypeslib.as_array(gp, shape=(Cn, n))[:] = np.asarray(g, dtype=np.float64).T
            return 0
        except Exception as e:
            err.append(e)
            return L.SFH_ERR_INVALID_ARG
    n = np.asarray(theta0s[0]).shape[0]
    st, out, stats = _nuts_call(L.lib.sfh_nuts_run, (L.sfh_batch_logdensity_fn(cb), None, n), theta0s, nsteps, nwarmup, max_depth, delta,
                                eps0, seed, inv_mass)
    if err:
        raise err[0]
    L.check(st)
    return out, stats


def native_sample_sfh_generic(inner_fg_batched, n_ages, params0, transforms, free, theta0s, nsteps, nwarmup=0, max_depth=8, eps0=None,
                              seed=0, inv_mass=None):
    """sample_sfh / tsample_sfh's chains (sfh_sample_sfh_nuts_generic) around a caller-supplied batched hierarchical
    ``inner_fg_batched(V[(n_ages + npar), C]) -> (-logL[C], G)`` over the natural variables -- user-defined models."""
    p0 = np.ascontiguousarray(params0, dtype=np.float64)
    tf = np.ascontiguousarray(transforms, dtype=np.int32)
    fr = np.ascontiguousarray(free, dtype=np.uint8)
    err = []

    def cb(_user, vp, nv, Cn, nlp, gp):
        try:
            V = np.ctypeslib.as_array(vp, shape=(Cn, nv)).T.copy(order="F")
            nl, G = inner_fg_batched(V)
            np.ctypeslib.as_array(nlp, shape=(Cn,))[:] = nl
            np.ctypeslib.as_array(gp, shape=(Cn, nv))[:] = np.asarray(G, dtype=np.float64).T
            return 0
        except Exception as e:
            err.append(e)
            return L.SFH_ERR_INVALID_ARG
    dp = C.POINTER(C.c_double)
    head = (L.sfh_batch_logdensity_fn(cb), None, int(n_ages), p0.shape[0], p0.ctypes.data_as(dp), tf.ctypes.data_as(C.POINTER(C.c_int32)),
            fr.ctypes.data_as(C.POINTER(C.c_uint8)))
    st, out, stats = _nuts_call(L.lib.sfh_sample_sfh_nuts_generic, head, theta0s, nsteps, nwarmup, max_depth, 0.8, eps0, seed, inv_mass)
    if err:
        raise err[0]
    L.check(st)
    return out, stats


def hmc_sample(models, data, nsteps, nchains=1, nwarmup=200, rng=None, x0=None, max_depth=8, batched=None, engine="host"):
    """hmc_sample(models, data, nsteps[, nchains])  (hmc_sample.jl:105-143): NUTS on theta = log(coeffs) with the
    Jacobian-corrected log-density of HMCModel.  Returns natural-unit samples of shape (nsteps, npar, nchains).
    With nchains > 1 (threads in the reference, hmc_sample.jl:123-141) the chains run as coroutines and, when `batched`,
    share one sfh_eval_fg_batched pass per round of gradient requests; each chain draws from its own spawned RNG.
    `batched=None` picks the batched pass when it pays: a stack of >= 64 MB (below that one fused single-vector evaluation
    per chain is cheaper) and at least 4 chains (profiles/r1_results.md)."""
    ds = device_stack(models, data)
    model = HMCModel(ds, None, data)
    rng = np.random.default_rng() if rng is None else rng
    x0 = renormalize_x0(data, ds, np.ones(ds.shape[1]) if x0 is None else np.asarray(x0, float))
    out = np.empty((nsteps, ds.shape[1], nchains))
    if engine == "native":                                                 # chains + batching inside the library (sfh_hmc_sample_nuts)
        st, res, _ = _nuts_call(L.lib.sfh_hmc_sample_nuts, (ds.ctx().handle,), [np.log(x0)] * nchains, nsteps, nwarmup, max_depth, 0.8,
                                None, int(rng.integers(0, 2**63)), None)
        L.check(st)
        for c in range(nchains):
            out[:, :, c] = np.exp(res[c][0])
        return out
    if engine != "host":
        raise ValueError("engine must be 'host' or 'native'")
    if nchains == 1:
        s, _, _ = nuts_sample(model.logdensity_and_gradient, np.log(x0), nsteps, nwarmup, max_depth, rng=rng)
        out[:, :, 0] = np.exp(s)                                           # back to natural units
        return out
    rngs = list(rng.spawn(nchains))
    if batched is None:
        batched = nchains >= 4 and ds.shape[0] * ds.shape[1] * np.dtype(ds.dtype).itemsize >= (64 << 20)
    if batched:
        res, _ = run_chains_batched(model.logdensity_and_gradient_batched, [np.log(x0)] * nchains, nsteps, nwarmup, max_depth, rngs)
    else:
        res = [nuts_sample(model.logdensity_and_gradient, np.log(x0), nsteps, nwarmup, max_depth, rng=rngs[c]) for c in range(nchains)]
    for c in range(nchains):
        out[:, :, c] = np.exp(res[c][0])
    return out


# ---------------------------------------------------------------------------------------------
def _expand_posterior(best, Z):
    """Transformed free-variable samples (nfree_total, N) -> natural-unit samples over ALL variables with the fixed
    parameters written in (generic_fitting.jl:640-658, exptransform_samples! transformations.jl:57-93)."""
    npar = best.MH_model.nparams() + best.disp_model.nparams()
    nj = best.mu.shape[0] - npar
    tf = np.array(list(best.MH_model.transforms()) + list(best.disp_model.transforms()))
    free = np.array(list(best.MH_model.free_params()) + list(best.disp_model.free_params()), dtype=bool)
    out = np.empty((best.mu.shape[0], Z.shape[1]))
    out[:nj] = np.exp(Z[:nj])
    tfree = tf[free]
    Zp = Z[nj:]
    out[nj:][free] = np.where(tfree[:, None] == 1, np.exp(Zp), np.where(tfree[:, None] == -1, -np.exp(Zp), Zp))
    par = np.array(list(best.MH_model.fittable_params()) + list(best.disp_model.fittable_params()), dtype=np.float64)
    out[nj:][~free] = par[~free][:, None]
    return out


def _native_sample_sfh(inst, starts, lens, cov, eps, nwarmup, max_depth, seed):
    """sfh_sample_sfh_nuts on the device-bound HierarchicalOptimizer log-density of `inst` (Jacobian corrections on)."""
    from .hierarchical import _bind
    ctx = _bind(inst.models, inst.logAge, inst.metallicities)
    tf = np.ascontiguousarray(list(inst.MH_model0.transforms()) + list(inst.disp_model0.transforms()), dtype=np.int32)
    free = np.ascontiguousarray(list(inst.MH_model0.free_params()) + list(inst.disp_model0.free_params()), dtype=np.uint8)
    init = np.ascontiguousarray(list(inst.MH_model0.fittable_params()) + list(inst.disp_model0.fittable_params()), dtype=np.float64)
    fx = np.ascontiguousarray(inst.MH_model0.fixed(), dtype=np.float64)
    dp = C.POINTER(C.c_double)
    head = (ctx.handle, inst.MH_model0.kind, fx.ctypes.data_as(dp), inst.disp_model0.kind, init.ctypes.data_as(dp),
            tf.ctypes.data_as(C.POINTER(C.c_int32)), free.ctypes.data_as(C.POINTER(C.c_uint8)))
    st, res, stats = _nuts_call(L.lib.sfh_sample_sfh_nuts, head, starts, lens, nwarmup, max_depth, 0.8, eps, seed, cov)
    L.check(st)
    return res, stats


def sample_sfh(bfgs_result, models, data, logAge, metallicities, Nsteps, eps=0.05, rng=None, nwarmup=0, max_depth=8, engine="host"):
    """sample_sfh (generic_fitting.jl:456-556): one NUTS chain over the hierarchical model started at the MLE with the MAP
    inverse Hessian as the Gaussian kinetic energy's M^-1 and initial step size `eps`.  Returns
    {"posterior_matrix": (nvariables, Nsteps) in natural units, "logp": (Nsteps,), "step_size": float}."""
    rng = np.random.default_rng() if rng is None else rng
    MAP, MLE = bfgs_result["map"], bfgs_result["mle"]
    inst = HierarchicalOptimizer(MLE.MH_model, MLE.disp_model, device_stack(models, data), data, logAge, metallicities, True, True, True)
    if engine == "native":
        res, _ = _native_sample_sfh(inst, [np.asarray(MLE.result.x, dtype=np.float64)], [Nsteps], np.asarray(MAP.invH, dtype=np.float64), eps,
                                    nwarmup, max_depth, int(rng.integers(0, 2**63)))
        s, lps, step = res[0]
        return {"posterior_matrix": _expand_posterior(MLE, s.T), "logp": lps, "step_size": step}
    s, lps, step = nuts_sample(inst.logdensity_and_gradient, np.asarray(MLE.result.x, dtype=np.float64), Nsteps, nwarmup, max_depth,
                               rng=rng, inv_mass=np.asarray(MAP.invH, dtype=np.float64), eps0=eps)
    return {"posterior_matrix": _expand_posterior(MLE, s.T), "logp": lps, "step_size": step}


def tsample_sfh(bfgs_result, models, data, logAge, metallicities, Nsteps, eps=0.05, rng=None, chain_length=100, nwarmup=0,
                max_depth=8, batched=True, engine="host"):
    """tsample_sfh (generic_fitting.jl:564-665): ceil(Nsteps / chain_length) short chains, each started from a draw of
    MvNormal(MLE minimizer, MAP.invH) (:586, :619).  The reference spawns one task per chain, every task evaluating its own
    hierarchical `fg!`; here the chains are coroutines whose gradient requests are served `batched` -- one
    sfh_eval_fg_hier_batched pass per round for all live chains.  Returns the same dictionary as `sample_sfh`."""
    rng = np.random.default_rng() if rng is None else rng
    MAP, MLE = bfgs_result["map"], bfgs_result["mle"]
    x0 = np.asarray(MLE.result.x, dtype=np.float64)
    cov = np.asarray(MAP.invH, dtype=np.float64)
    cov = (cov + cov.T) / 2
    lens = [min(chain_length, Nsteps - a) for a in range(0, Nsteps, chain_length)]   # Iterators.partition (:607)
    starts = rng.multivariate_normal(x0, cov, size=len(lens))
    rngs = list(rng.spawn(len(lens)))
    inst = HierarchicalOptimizer(MLE.MH_model, MLE.disp_model, device_stack(models, data), data, logAge, metallicities, True, True, True)

    def chain(th0, nsteps, nw, md, rng=None, _n=iter(lens)):
        return nuts_chain(th0, next(_n), nw, md, rng=rng, inv_mass=cov, eps0=eps)
    if engine == "native":                                                 # chain threads + batching inside the library
        res, stats = _native_sample_sfh(inst, list(starts), lens, cov, eps, nwarmup, max_depth, int(rng.integers(0, 2**63)))
    elif batched and len(lens) > 1:
        res, stats = run_chains_batched(inst.logdensity_and_gradient_batched, list(starts), 0, nwarmup, max_depth, rngs, chain=chain)
    else:
        res = [nuts_sample(inst.logdensity_and_gradient, starts[k], lens[k], nwarmup, max_depth, rng=rngs[k], inv_mass=cov, eps0=eps)
               for k in range(len(lens))]
    Z = np.concatenate([r[0] for r in res], axis=0).T                                  # reduce(hcat, ...) (:633)
    return {"posterior_matrix": _expand_posterior(MLE, Z), "logp": np.concatenate([r[1] for r in res]),
            "step_size": float(np.mean([r[2] for r in res]))}
