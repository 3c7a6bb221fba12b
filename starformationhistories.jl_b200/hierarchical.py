"""Host-side mirror of the reference's hierarchical (metallicity-model) interface (L3/L4 of SURVEY.md):

    AbstractMZR / PowerLawMZR            src/fitting/hierarchical/mzr.jl:6-48, 263-284
    AbstractAMR / LinearAMR / LogarithmicAMR   src/fitting/hierarchical/amr.jl:6-48, 181-213, 250-298
    GaussianDispersion                   src/fitting/hierarchical/dispersion_models.jl:80-110
    calculate_coeffs                     mzr.jl:50-79, amr.jl:50-73
    fg! (hierarchical method)            mzr.jl:84-215, amr.jl:78-173           -> fg_
    HierarchicalOptimizer + logdensity_and_gradient   generic_fitting.jl:44-58, 90-199

The model classes carry the reference's small model API (callable, nparams, fittable_params, gradient,
update_params, transforms, free_params) as plain scalar host code -- it is metadata.  The O(Nb*T)
evaluation and the O(T) coefficient expansion / chain rule of every iteration run on the device
through ``sfh_eval_fg_hier``; only Nj+3 numbers cross the host boundary per evaluation.
"""
from __future__ import annotations

import ctypes as C
import math

import numpy as np

from . import _lib as L
from .fitting import DeviceStack, _dp, device_stack

LOGTEN = math.log(10.0)


# ---------------------------------------------------------------------------------------------
# dispersion model
# ---------------------------------------------------------------------------------------------
class GaussianDispersion:
    kind = L.SFH_DISP_GAUSSIAN

    def __init__(self, sigma, free=(True,)):
        if sigma <= 0:
            raise ValueError("σ must be > 0")                              # dispersion_models.jl:84
        self.sigma = float(sigma)
        self.free = tuple(bool(f) for f in free)

    def __call__(self, x, mu):
        return math.exp(-(((x - mu) / self.sigma) ** 2) / 2)              # :92

    def gradient(self, x, mu):
        A = self(x, mu)
        return (A * (x - mu) ** 2 / self.sigma ** 3, A * (x - mu) / self.sigma ** 2)   # :95-100 (σ, μ)

    def nparams(self): return 1
    def fittable_params(self): return (self.sigma,)
    def update_params(self, new): return GaussianDispersion(new[0], self.free)
    def transforms(self): return (1,)
    def free_params(self): return self.free
    def __eq__(self, o): return isinstance(o, GaussianDispersion) and (self.sigma, self.free) == (o.sigma, o.free)


# ---------------------------------------------------------------------------------------------
# metallicity models
# ---------------------------------------------------------------------------------------------
class _MHModel:
    is_mzr = False

    def nparams(self): return 2
    def fittable_params(self): return (self.alpha, self.beta)
    def free_params(self): return self.free


class PowerLawMZR(_MHModel):
    """[M/H](M*) = MH0 + α (log10 M* - logMstar0)   (mzr.jl:263-284)"""
    kind = L.SFH_MH_POWERLAW_MZR
    is_mzr = True

    def __init__(self, alpha, MH0, logMstar0=6.0, free=(True, True)):
        if alpha < 0:
            raise ValueError("α must be ≥ 0")                              # mzr.jl:268
        self.alpha, self.beta, self.logMstar0 = float(alpha), float(MH0), float(logMstar0)
        self.free = tuple(bool(f) for f in free)

    MH0 = property(lambda self: self.beta)

    def __call__(self, Mstar): return self.beta + self.alpha * (math.log10(Mstar) - self.logMstar0)

    def gradient(self, Mstar):
        return (math.log10(Mstar) - self.logMstar0, 1.0, self.alpha / Mstar / LOGTEN)

    def update_params(self, new): return PowerLawMZR(new[0], new[1], self.logMstar0, self.free)
    def transforms(self): return (1, 0)
    def fixed(self): return np.array([self.logMstar0, 0, 0, 0], dtype=np.float64)
    def __eq__(self, o):
        return isinstance(o, PowerLawMZR) and (self.alpha, self.beta, self.logMstar0, self.free) == (o.alpha, o.beta, o.logMstar0, o.free)


class LinearAMR(_MHModel):
    """μ_j = β + α (T_max - t_j [Gyr])   (amr.jl:181-213)"""
    kind = L.SFH_MH_LINEAR_AMR

    def __init__(self, alpha, beta, T_max=13.7, free=(True, True)):
        if alpha < 0:
            raise ValueError("α must be ≥ 0")                              # amr.jl:187
        if T_max <= 0:
            raise ValueError("T_max must be > 0")                          # amr.jl:189
        self.alpha, self.beta, self.T_max = float(alpha), float(beta), float(T_max)
        self.free = tuple(bool(f) for f in free)

    @classmethod
    def from_constraints(cls, constraint1, constraint2, T_max=13.7, free=(True, True)):
        """LinearAMR(constraint1, constraint2, T_max) (amr.jl:229-245): the line through two ([M/H], lookback time [Gyr]) points."""
        if len(constraint1) != 2 or len(constraint2) != 2:
            raise ValueError("length(constraint1) == length(constraint2) == 2 must hold")
        times, mhs = (constraint1[1], constraint2[1]), (constraint1[0], constraint2[0])
        dt, dmh = times[1] - times[0], mhs[1] - mhs[0]
        if dt == 0:
            raise ValueError("Constraints are given at identical times.")
        if not (np.sign(dt) != np.sign(dmh)):
            raise ValueError("The constraints indicate metallicity is decreasing towards present-day; not allowed under LinearAMR.")
        alpha = dmh / -dt
        return cls(alpha, min(mhs) - alpha * (T_max - max(times)), T_max, free)

    def __call__(self, logAge): return self.beta + self.alpha * (self.T_max - 10.0 ** (logAge - 9))
    def gradient(self, logAge): return (self.T_max - 10.0 ** (logAge - 9), 1.0)
    def update_params(self, new): return LinearAMR(new[0], new[1], self.T_max, self.free)
    def __eq__(self, o):
        return isinstance(o, LinearAMR) and (self.alpha, self.beta, self.T_max, self.free) == (o.alpha, o.beta, o.T_max, o.free)
    def transforms(self): return (1, 0)
    def fixed(self): return np.array([self.T_max, 0, 0, 0], dtype=np.float64)


def Y_from_Z(Z, Y_p=0.2485, gamma=1.78):
    """src/utilities.jl:118"""
    return Y_p + gamma * Z


def X_from_Z(Z, Y_p=0.2485, gamma=1.78):
    """src/utilities.jl:123-125"""
    return 1 - (Y_from_Z(Z, Y_p, gamma) + Z)


def Z_from_MH(MH, solZ=0.01524, Y_p=0.2485, gamma=1.78):
    """src/utilities.jl:160-176: inverse of MH_from_Z under Y = Y_p + gamma Z."""
    zoverx = 10.0 ** (MH + math.log10(solZ / X_from_Z(solZ, Y_p, gamma)))
    return (1 - Y_p) * zoverx / (1 + (1 + gamma) * zoverx)


def dZ_dMH(MH, solZ=0.01524, Y_p=0.2485, gamma=1.78):
    """src/utilities.jl:182-187"""
    prefac = 10.0 ** MH * solZ
    X = X_from_Z(solZ, Y_p, gamma)
    return -prefac * X * (Y_p - 1) * LOGTEN / (X + prefac * (1 + gamma)) ** 2


def MH_from_Z(Z, solZ=0.01524, Y_p=0.2485, gamma=1.78):
    """src/utilities.jl:138-146 (NaN instead of a DomainError when X <= 0)."""
    X = 1 - ((Y_p + gamma * Z) + Z)
    Xs = 1 - ((Y_p + gamma * solZ) + solZ)
    return math.log10(Z / (X * solZ) * Xs) if X > 0 else float("nan")


def dMH_dZ(Z, solZ=0.01524, Y_p=0.2485, gamma=1.78):
    """src/utilities.jl:152-155"""
    return (Y_p - 1) / (LOGTEN * Z * (Y_p + Z + gamma * Z - 1))


class LogarithmicAMR(_MHModel):
    """Z_j = β + α (T_max - t_j), μ_j = MH_from_Z(Z_j)   (amr.jl:250-298).  Only the default
    MH_from_Z / dMH_dZ pair (PARSEC: solZ, Y_p, γ) has a device-side chain rule."""
    kind = L.SFH_MH_LOG_AMR

    def __init__(self, alpha, beta, T_max=13.7, free=(True, True), solZ=0.01524, Y_p=0.2485, gamma=1.78):
        if alpha < 0:
            raise ValueError("α must be ≥ 0")                              # amr.jl:262
        if beta < 0:
            raise ValueError("β must be ≥ 0")                              # amr.jl:264
        if T_max <= 0:
            raise ValueError("T_max must be > 0")                          # amr.jl:266
        self.alpha, self.beta, self.T_max = float(alpha), float(beta), float(T_max)
        self.solZ, self.Y_p, self.gamma = float(solZ), float(Y_p), float(gamma)
        self.free = tuple(bool(f) for f in free)

    @classmethod
    def from_constraints(cls, constraint1, constraint2, T_max=13.7, free=(True, True), solZ=0.01524, Y_p=0.2485, gamma=1.78):
        """LogarithmicAMR(constraint1, constraint2, T_max) (amr.jl:317-338): Z linear in lookback time through two
        ([M/H], lookback time [Gyr]) points, with the default Z_from_MH conversion."""
        if len(constraint1) != 2 or len(constraint2) != 2:
            raise ValueError("length(constraint1) == length(constraint2) == 2 must hold")
        times = (constraint1[1], constraint2[1])
        zs = (Z_from_MH(constraint1[0], solZ, Y_p, gamma), Z_from_MH(constraint2[0], solZ, Y_p, gamma))
        dt, dz = times[1] - times[0], zs[1] - zs[0]
        if dt == 0:
            raise ValueError("Constraints are given at identical times.")
        if not (np.sign(dt) != np.sign(dz)):
            raise ValueError("The constraints indicate metallicity is decreasing towards present-day; not allowed under LogarithmicAMR.")
        alpha = dz / -dt
        beta = min(zs) - alpha * (T_max - max(times))
        if beta < 0:
            raise ValueError("Given constraints result in a metal mass fraction Z < 0 at T_max. Please revise arguments.")
        return cls(alpha, beta, T_max, free, solZ, Y_p, gamma)

    def __eq__(self, o):
        return isinstance(o, LogarithmicAMR) and (self.alpha, self.beta, self.T_max, self.free, self.solZ, self.Y_p, self.gamma) == \
            (o.alpha, o.beta, o.T_max, o.free, o.solZ, o.Y_p, o.gamma)

    def __call__(self, logAge):
        return MH_from_Z(self.beta + self.alpha * (self.T_max - 10.0 ** (logAge - 9)), self.solZ, self.Y_p, self.gamma)

    def gradient(self, logAge):
        age = 10.0 ** (logAge - 9)
        d = dMH_dZ(self.beta + self.alpha * (self.T_max - age), self.solZ, self.Y_p, self.gamma)
        return (d * (self.T_max - age), d)

    def update_params(self, new):
        return LogarithmicAMR(new[0], new[1], self.T_max, self.free, self.solZ, self.Y_p, self.gamma)

    def transforms(self): return (1, 1)
    def fixed(self): return np.array([self.T_max, self.solZ, self.Y_p, self.gamma], dtype=np.float64)


def nparams(*models): return sum(m.nparams() for m in models)     # hierarchical_models.jl:17


# ---------------------------------------------------------------------------------------------
# device calls
# ---------------------------------------------------------------------------------------------
def _bind(ds: DeviceStack, logAge, MH):
    """sfh_hier_bind once per (context, logAge, MH): the grouping of mzr.jl:131-140."""
    ctx = ds.ctx()
    # fast path: the same array OBJECTS as last time (every driver passes the arrays it captured once) with
    # unchanged end points -- skips hashing 2 x T doubles on every evaluation
    last = getattr(ctx, "bound_objs", None)
    if (last is not None and last[0] is logAge and last[1] is MH and isinstance(logAge, np.ndarray) and isinstance(MH, np.ndarray)
            and logAge.shape == last[2] and (logAge[0], logAge[-1], MH[0], MH[-1]) == last[3]):
        return ctx
    ctx.bound_objs = None
    la = np.ascontiguousarray(logAge, dtype=np.float64)
    mh = np.ascontiguousarray(MH, dtype=np.float64)
    if la.shape != mh.shape:
        raise ValueError("length(logAge) != length(metallicities)")        # mzr.jl:57
    if la.shape[0] != ds.shape[1]:
        raise ValueError("length(logAge) != number of templates")
    key = (la.tobytes(), mh.tobytes())
    if ctx.bound_key != key:
        n = C.c_int64()
        L.check(L.lib.sfh_hier_bind(ctx.handle, _dp(la), _dp(mh), C.byref(n)))
        ctx.bound_key, ctx.n_ages = key, int(n.value)
    if isinstance(logAge, np.ndarray) and isinstance(MH, np.ndarray) and la.shape[0] > 0:
        ctx.bound_objs = (logAge, MH, logAge.shape, (logAge[0], logAge[-1], MH[0], MH[-1]))
    return ctx


def calculate_coeffs(MHmodel, dispmodel, mstars, logAge=None, metallicities=None, models=None):
    """``calculate_coeffs(MHmodel, dispmodel, R, logAge, MH)`` (mzr.jl:50-79 / amr.jl:50-73), and the three-argument method
    ``calculate_coeffs(result, logAge, MH)`` for a fit_sfh result (bfgs_result.jl:81-90).

    With ``models`` (a DeviceStack) the expansion runs through the device prologue kernel that every
    hierarchical evaluation uses.  Without it -- the reference signature has no stack argument, and this
    form is only used for one-shot set-up such as building x0 -- the same O(T) formulae are evaluated
    on the host in float64."""
    if logAge is None and metallicities is None:
        # calculate_coeffs(result, logAge, MH) (bfgs_result.jl:81-90, 151-154): best-fit models and stellar masses of a
        # fit_sfh result; the MLE of a {"map", "mle"} pair (CompositeBFGSResult)
        res = MHmodel["mle"] if isinstance(MHmodel, dict) else MHmodel
        nj = res.mu.shape[0] - res.MH_model.nparams() - res.disp_model.nparams()
        return calculate_coeffs(res.MH_model, res.disp_model, res.mu[:nj], dispmodel, mstars, models=models)
    la = np.asarray(logAge, dtype=np.float64)
    mh = np.asarray(metallicities, dtype=np.float64)
    R = np.asarray(mstars, dtype=np.float64)
    if la.shape != mh.shape:
        raise ValueError("length(logAge) != length(metallicities)")        # mzr.jl:57
    _, first, inv = np.unique(la, return_index=True, return_inverse=True)
    order = np.argsort(first, kind="stable")                               # unique(logAge): first-appearance order (mzr.jl:54)
    uniq = la[first[order]]
    if R.shape[0] != uniq.shape[0]:
        raise ValueError("Length of `mstars` must be the same as `unique(logAge)`.")   # mzr.jl:55-56
    if models is not None:
        ds = device_stack(models, None)
        ctx = _bind(ds, la, mh)
        v = np.concatenate([R, [MHmodel.alpha, MHmodel.beta, dispmodel.sigma]])
        out = np.empty(ds.shape[1])
        fx = MHmodel.fixed()
        L.check(L.lib.sfh_calculate_coeffs(ctx.handle, MHmodel.kind, _dp(fx), dispmodel.kind, _dp(v), _dp(out)))
        return out
    if MHmodel.is_mzr:
        s = np.argsort(-uniq, kind="stable")                               # sortperm(rev=true)  mzr.jl:61
        cum = np.empty_like(R)
        cum[s] = np.cumsum(R[s])                                           # mzr.jl:66
        mu = np.array([MHmodel(c) for c in cum])
    else:
        mu = np.array([MHmodel(a) for a in uniq])
    rank = np.empty(order.shape[0], dtype=np.int64)
    rank[order] = np.arange(order.shape[0])
    jidx = rank[inv.reshape(-1)]                                           # template -> its age's position in unique(logAge)  (:71)
    A = np.exp(-(((mh - mu[jidx]) / dispmodel.sigma) ** 2) / 2)            # :73
    Asum = np.bincount(jidx, weights=A, minlength=uniq.shape[0])
    return A * R[jidx] / Asum[jidx]                                        # :74-76


def fg_(F, G, MHmodel0, dispmodel0, variables, models, data, composite, logAge, metallicities):
    """Hierarchical ``fg!`` (mzr.jl:84-215 / amr.jl:78-173): returns -logL if ``F is not None``; fills ``G``
    (length Nj + nparams) with d(-logL)/d variables if ``G is not None``.  ``composite`` is accepted for
    signature parity and ignored (scratch lives on the device)."""
    ds = device_stack(models, data)
    ctx = _bind(ds, logAge, metallicities)
    v = np.ascontiguousarray(variables, dtype=np.float64)
    if v.shape[0] != ctx.n_ages + 3:
        raise ValueError("length(variables) != length(unique(logAge)) + nparams")      # mzr.jl:55, :126
    if G is not None and np.asarray(G).shape[0] != v.shape[0]:
        raise ValueError("axes(G) != axes(variables)")                     # mzr.jl:126
    free = np.array(list(MHmodel0.free_params()) + list(dispmodel0.free_params()) + [0], dtype=np.uint8)
    nl = C.c_double()
    g = np.empty(v.shape[0]) if G is not None else None
    fx = MHmodel0.fixed()
    L.check(L.lib.sfh_eval_fg_hier(ctx.handle, MHmodel0.kind, _dp(fx), dispmodel0.kind, _dp(v),
                                   free.ctypes.data_as(C.POINTER(C.c_uint8)), C.byref(nl), _dp(g) if g is not None else None))
    if G is not None:
        G[...] = g
    return nl.value if F is not None else None


def fg_batched_(MHmodel0, dispmodel0, V, models, data, logAge, metallicities, want_G=True):
    """Hierarchical ``fg!`` for every column of V ((Nj + nparams) x C, natural units) in one device pass
    (sfh_eval_fg_hier_batched): returns (-logL[C], G[(Nj + nparams), C] or None)."""
    ds = device_stack(models, data)
    ctx = _bind(ds, logAge, metallicities)
    V = np.asfortranarray(V, dtype=np.float64)
    if V.ndim != 2 or V.shape[0] != ctx.n_ages + 3:
        raise ValueError("size(V,1) != length(unique(logAge)) + nparams")
    free = np.array(list(MHmodel0.free_params()) + list(dispmodel0.free_params()) + [0], dtype=np.uint8)
    nl = np.empty(V.shape[1])
    G = np.empty(V.shape, order="F") if want_G else None
    fx = MHmodel0.fixed()
    L.check(L.lib.sfh_eval_fg_hier_batched(ctx.handle, MHmodel0.kind, _dp(fx), dispmodel0.kind, _dp(V), V.shape[1],
                                           free.ctypes.data_as(C.POINTER(C.c_uint8)), _dp(nl), _dp(G) if want_G else None))
    return nl, G


def exptransform(params, transforms):
    """src/fitting/hierarchical/transformations.jl:44"""
    return np.array([math.exp(p) if t == 1 else (-math.exp(p) if t == -1 else p) for p, t in zip(params, transforms)])


def logtransform(params, transforms):
    """src/fitting/hierarchical/transformations.jl:18"""
    return np.array([math.log(p) if t == 1 else (math.log(-p) if t == -1 else p) for p, t in zip(params, transforms)])


class HierarchicalOptimizer:
    """generic_fitting.jl:44-58 -- the LogDensityProblems adapter used by fit_sfh / sample_sfh."""

    def __init__(self, MH_model0, disp_model0, models, data, logAge, metallicities, F=True, G=True,
                 jacobian_corrections=True):
        self.MH_model0, self.disp_model0 = MH_model0, disp_model0
        self.models = device_stack(models, data)
        self.data, self.logAge, self.metallicities = data, np.asarray(logAge, float), np.asarray(metallicities, float)
        self.F, self.G, self.jacobian_corrections = F, G, bool(jacobian_corrections)

    def dimension(self):
        free = list(self.MH_model0.free_params()) + list(self.disp_model0.free_params())
        return _bind(self.models, self.logAge, self.metallicities).n_ages + sum(free)

    def _layout(self, nx):
        """Index sets of generic_fitting.jl:107-145 for an input of length nx (cached: they only depend on the models)."""
        lay = getattr(self, "_lay", None)
        if lay is None or lay[0] != nx:
            tf = np.array(list(self.MH_model0.transforms()) + list(self.disp_model0.transforms()))
            free = np.array(list(self.MH_model0.free_params()) + list(self.disp_model0.free_params()), dtype=bool)
            npar = tf.shape[0]
            nbins = nx - npar + int((~free).sum())                         # :116-118
            init = np.array(list(self.MH_model0.fittable_params()) + list(self.disp_model0.fittable_params()), dtype=np.float64)
            pos = np.zeros(nbins + npar, dtype=bool); pos[:nbins] = True
            pos[nbins:] = (tf == 1) & free                                 # :143-145
            neg = np.zeros(nbins + npar, dtype=bool); neg[nbins:] = (tf == -1) & free
            lay = (nx, tf, free, npar, nbins, init, pos, neg, tf[free])
            self._lay = lay
        return lay

    def logdensity_and_gradient(self, xvec):
        """generic_fitting.jl:90-199.  Returns (+logp, +grad) over the FREE transformed variables
        (only ``logp`` / only ``grad`` when ``G`` / ``F`` is None)."""
        xvec = np.asarray(xvec, dtype=np.float64)
        ret_F, ret_G = self.F is not None, self.G is not None
        _, tf, free, npar, nbins, init, pos, neg, tfree = self._layout(xvec.shape[0])
        x = np.empty(nbins + npar)
        x[:nbins] = np.exp(xvec[:nbins])                                   # :127
        xp = xvec[nbins:]
        x[nbins:][free] = np.where(tfree == 1, np.exp(xp), np.where(tfree == -1, -np.exp(xp), xp))   # :129-131
        x[nbins:][~free] = init[~free]                                     # :134-136
        G2 = np.empty_like(x) if ret_G else None
        nlogL = fg_(self.F, G2, self.MH_model0, self.disp_model0, x, self.models, self.data, None,
                    self.logAge, self.metallicities)                       # :140
        has_neg = bool(neg.any())
        if self.jacobian_corrections:                                      # :148-160
            if ret_F:
                nlogL -= float(np.log(x[pos]).sum())
                if has_neg:
                    nlogL += float(np.log(x[neg]).sum())
            if ret_G:
                G2[pos] = G2[pos] * x[pos] - 1
                if has_neg:
                    G2[neg] = -G2[neg] * x[neg] + 1
        elif ret_G:                                                        # :161-169 (with the +Nbins the reference forgets)
            G2[pos] = G2[pos] * x[pos]
            if has_neg:
                G2[neg] = -G2[neg] * x[neg]
        if not ret_G:
            return -nlogL if ret_F else None                               # :172-178
        G = np.empty_like(xvec)
        G[:nbins] = G2[:nbins]
        G[nbins:] = G2[nbins:][free]                                       # :181-189
        return (-nlogL, -G) if ret_F else -G                               # :193-197

    def native_bfgs(self, xstart, g_abstol=1e-8, iterations=5000, alphaguess=0):
        """Minimise -logdensity over the free transformed variables with the library's BFGS loop (sfh_fit_sfh_bfgs): the
        objective of fg_map! / fg_mle! (generic_fitting.jl:306-325) evaluated, transformed and iterated natively -- one C
        call per optimisation.  Returns an object with scipy's field names (x, hess_inv, fun, nit, nfev, success)."""
        from .solvers import _bfgs_opts, _native_result
        x = np.array(xstart, dtype=np.float64)
        _, tf, free, npar, nbins, init, pos, neg, tfree = self._layout(x.shape[0])
        ctx = _bind(self.models, self.logAge, self.metallicities)
        if nbins != ctx.n_ages:
            raise ValueError("length(x0) != length(unique(logAge)) + number of free parameters")
        invH = np.empty((x.shape[0],) * 2, order="F")
        rep, o, dp = L.sfh_bfgs_report(), _bfgs_opts(g_abstol, iterations, alphaguess), C.POINTER(C.c_double)
        fx = self.MH_model0.fixed()
        tf32, free8 = np.ascontiguousarray(tf, dtype=np.int32), np.ascontiguousarray(free, dtype=np.uint8)
        L.check(L.lib.sfh_fit_sfh_bfgs(ctx.handle, self.MH_model0.kind, _dp(fx), self.disp_model0.kind, _dp(np.ascontiguousarray(init)),
                                       tf32.ctypes.data_as(C.POINTER(C.c_int32)), free8.ctypes.data_as(C.POINTER(C.c_uint8)),
                                       int(self.jacobian_corrections), x.ctypes.data_as(dp), C.byref(o), C.byref(rep), invH.ctypes.data_as(dp)))
        return _native_result(x, invH, rep)

    def logdensity_and_gradient_batched(self, X):
        """The same for C chains at once: X is (dimension, C); one device pass (sfh_eval_fg_hier_batched) serves all chains.
        Returns (+logp[C], +grad[dimension, C])  (generic_fitting.jl:90-199 per column)."""
        X = np.asarray(X, dtype=np.float64)
        tf = np.array(list(self.MH_model0.transforms()) + list(self.disp_model0.transforms()))
        free = np.array(list(self.MH_model0.free_params()) + list(self.disp_model0.free_params()), dtype=bool)
        npar = tf.shape[0]
        nbins = X.shape[0] - npar + int((~free).sum())
        Cn = X.shape[1]
        V = np.empty((nbins + npar, Cn))
        V[:nbins] = np.exp(X[:nbins])                                      # :127
        tfree = tf[free]
        Xp = X[nbins:]
        V[nbins:][free] = np.where(tfree[:, None] == 1, np.exp(Xp), np.where(tfree[:, None] == -1, -np.exp(Xp), Xp))   # :129-131
        init = np.array(list(self.MH_model0.fittable_params()) + list(self.disp_model0.fittable_params()))
        V[nbins:][~free] = init[~free][:, None]                            # :134-136
        nl, G2 = fg_batched_(self.MH_model0, self.disp_model0, V, self.models, self.data, self.logAge, self.metallicities)
        pos = np.zeros(nbins + npar, dtype=bool); pos[:nbins] = True
        pos[nbins:] = (tf == 1) & free                                     # :143-145
        neg = np.zeros(nbins + npar, dtype=bool); neg[nbins:] = (tf == -1) & free
        if self.jacobian_corrections:                                      # :148-160
            nl = nl - np.log(V[pos]).sum(axis=0)
            if neg.any():                                                  # (no built-in model has a -1 transform)
                nl = nl + np.log(V[neg]).sum(axis=0)
            G2[pos] = G2[pos] * V[pos] - 1
            G2[neg] = -G2[neg] * V[neg] + 1
        else:                                                              # :161-169
            G2[pos] = G2[pos] * V[pos]
            G2[neg] = -G2[neg] * V[neg]
        G = np.empty_like(X)
        G[:nbins] = G2[:nbins]
        G[nbins:] = G2[nbins:][free]                                       # :181-189
        return -nl, -G                                                     # :193-197
