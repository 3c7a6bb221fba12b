"""Host-side mirror of the reference's core evaluation interface (L1/L2 of SURVEY.md):

    stack_models            src/fitting/utilities.jl:12-13
    composite!              src/fitting/fitting_base.jl:20-33, 55-65      -> composite_
    loglikelihood           src/fitting/fitting_base.jl:84-96, 107-125
    grad-loglikelihood      src/fitting/fitting_base.jl:144-160, 171-182, 193-211  -> grad_loglikelihood
    grad-loglikelihood!     src/fitting/fitting_base.jl:227-256, 265-285  -> grad_loglikelihood_
    fg!                     src/fitting/solvers.jl:20-38                  -> fg_

Same argument order, meaning and error behaviour (the reference's ``@argcheck`` ArgumentErrors are
ValueErrors here); Julia's ``!`` suffix becomes a trailing underscore and ``nothing`` becomes
``None``.  Every function is a thin call through the C-ABI of ``libsfhcuda.so``; nothing is
computed in Python.  ``models`` is a :class:`DeviceStack` (the device-resident mirror of the
``stack_models`` matrix, uploaded once per fit) or a plain array / list of matrices, for which a
DeviceStack is created on first use and cached by identity (mechanism 2).
"""
from __future__ import annotations

import ctypes as C
import threading
import weakref

import numpy as np

from . import _lib as L

_DT = {np.dtype(np.float32): L.SFH_F32, np.dtype(np.float64): L.SFH_F64, np.dtype(np.int64): L.SFH_I64}


def _dp(a: np.ndarray):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def stack_models(models):
    """``reduce(hcat, map(vec, models))`` -- column-major ``vec`` of each matrix, one template per column."""
    return np.asfortranarray(np.stack([np.asarray(m).reshape(-1, order="F") for m in models], axis=1))


def _as_stack_matrix(models):
    """Accept the flat (Nb x T) layout or the vector-of-matrices layout; return F-ordered (Nb, T)."""
    if isinstance(models, (list, tuple)):
        return stack_models(models)
    M = np.asarray(models)
    if M.ndim != 2:
        raise ValueError("models must be an (nbins x ntemplates) matrix or a list of matrices")
    return M


class _Ctx:
    """One sfh_ctx per host thread (TaskLocalValue analogue, hmc_sample.jl:127)."""

    def __init__(self, stack_handle, stream=None):
        h = C.c_void_p()
        L.check(L.lib.sfh_ctx_create(stack_handle, C.c_void_p(stream) if stream else None, C.byref(h)))
        self.handle = h
        self.bound_key = None
        self.n_ages = 0
        self._fin = weakref.finalize(self, L.lib.sfh_ctx_destroy, h)

    def close(self):
        self._fin()


class DeviceStack:
    """Device-resident template stack + observed Hess diagram.

    Parameters
    ----------
    models : (Nb, T) array (column-major preferred) or list of T matrices -- what ``stack_models`` builds.
    data   : (Nb,) vector or matrix of the same shape as each template (float32/float64/int64).
    dtype  : storage dtype on device (default: ``models.dtype``; float32 stacks accumulate in FP64).
    rows   : optional ``(row_begin, row_end)`` bin-row shard held by this process (multi-GPU).
    """

    def __init__(self, models, data, dtype=None, device=0, rows=None, clamp_eps=0.0, tile_bins=0, cluster=0,
                 force_unfused=False, consumer_warps=0, variant=0):
        M = _as_stack_matrix(models)
        dt = np.dtype(dtype) if dtype is not None else M.dtype
        if dt not in (np.dtype(np.float32), np.dtype(np.float64)):
            dt = np.dtype(np.float64)
        M = np.asfortranarray(M, dtype=dt)
        d = np.asarray(data)
        d = d.reshape(-1, order="F")
        if d.dtype not in _DT:
            d = d.astype(np.float64)
        d = np.ascontiguousarray(d)
        if d.shape[0] != M.shape[0]:
            raise ValueError("axes(models,1) != axes(data,1)")            # solvers.jl:11
        self.shape = M.shape
        self.dtype = dt
        o = L.sfh_opts()
        o.struct_size = C.sizeof(L.sfh_opts)
        o.device = device
        if rows is not None:
            o.row_begin, o.row_end = int(rows[0]), int(rows[1])
        o.clamp_eps = clamp_eps
        o.tile_bins, o.cluster, o.force_unfused, o.consumer_warps = tile_bins, cluster, int(force_unfused), consumer_warps
        o.variant = variant
        h = C.c_void_p()
        L.check(L.lib.sfh_stack_create(C.byref(h), M.ctypes.data_as(C.c_void_p), M.shape[0], M.shape[1], _DT[dt],
                                       d.ctypes.data_as(C.c_void_p), _DT[d.dtype], C.byref(o)))
        self._finish_init(h)
        self._data_ref = _data_ref(data)

    def _finish_init(self, h):
        self.handle = h
        self._tls = threading.local()
        self._ctxs = []
        self._fin = weakref.finalize(self, DeviceStack._destroy, h, self._ctxs)
        self.rows = self.info().row_end - self.info().row_begin

    @classmethod
    def from_points(cls, edges, point_lists, data=None, dtype=np.float64, device=0, rows=None, force_unfused=False):
        """Build the stack on the device from per-template point lists (include/sfhcuda.h: sfh_stack_create_from_points).

        edges       : (xedges, yedges) uniform bin edges (what ``calculate_edges`` returns), nx+1 and ny+1 values.
        point_lists : one ``(colors, mags, color_err, mag_err, weights, cov_mult)`` tuple per template -- the arguments
                      ``bin_cmd_smooth`` (src/StarFormationHistories.jl:574-621) receives for that template.
        data        : observed Hess diagram (nx, ny) or its vec; zeros if omitted."""
        xe, ye = (np.asarray(e, dtype=np.float64) for e in edges)
        nx, ny = xe.shape[0] - 1, ye.shape[0] - 1
        if nx < 2 or ny < 2:
            raise ValueError("need at least 2 x 2 Hess bins")
        xstep, ystep = (xe[-1] - xe[0]) / nx, (ye[-1] - ye[0]) / ny
        for e, st in ((xe, xstep), (ye, ystep)):
            if not (st > 0 and np.allclose(np.diff(e), st, rtol=1e-9, atol=0)):
                raise ValueError("edges must be uniform, increasing ranges")       # addstar! :372-374
        offs = np.zeros(len(point_lists) + 1, dtype=np.int64)
        cols = [[] for _ in range(5)]
        cov = np.zeros(len(point_lists), dtype=np.int32)
        for t, pl in enumerate(point_lists):
            arrs = [np.ascontiguousarray(a, dtype=np.float64).reshape(-1) for a in pl[:5]]
            if any(a.shape != arrs[0].shape for a in arrs):
                raise ValueError("axes(colors) == axes(mags) == axes(color_err) == axes(mag_err) == axes(weights) must hold")  # :580
            if int(pl[5]) not in (-1, 0, 1):
                raise ValueError("cov_mult must be -1, 0 or 1")                     # :579
            cov[t] = int(pl[5])
            offs[t + 1] = offs[t] + arrs[0].shape[0]
            for k in range(5):
                cols[k].append(arrs[k])
        flat = [np.concatenate(c) if c else np.zeros(0) for c in cols]
        self = cls.__new__(cls)
        dt = np.dtype(dtype)
        self.shape = (nx * ny, len(point_lists))
        self.dtype = dt
        self.edges = (xe, ye)
        o = L.sfh_opts()
        o.struct_size = C.sizeof(L.sfh_opts)
        o.device = device
        if rows is not None:
            o.row_begin, o.row_end = int(rows[0]), int(rows[1])
        o.force_unfused = int(force_unfused)
        d = None
        if data is not None:
            d = np.asarray(data).reshape(-1, order="F")
            if d.dtype not in _DT:
                d = d.astype(np.float64)
            d = np.ascontiguousarray(d)
            if d.shape[0] != nx * ny:
                raise ValueError("axes(models,1) != axes(data,1)")
        h = C.c_void_p()
        i32p = C.POINTER(C.c_int32)
        L.check(L.lib.sfh_stack_create_from_points(
            C.byref(h), nx, ny, float(xe[0]), float(xstep), float(ye[0]), float(ystep), len(point_lists),
            offs.ctypes.data_as(C.POINTER(C.c_int64)), *[_dp(a) for a in flat], cov.ctypes.data_as(i32p), _DT[dt],
            d.ctypes.data_as(C.c_void_p) if d is not None else None, _DT[d.dtype] if d is not None else _DT[np.dtype(np.float64)],
            C.byref(o)))
        self._finish_init(h)
        self._data_ref = _data_ref(data)
        return self

    @classmethod
    def synthetic(cls, nbins, ntemplates, dtype, seed, scale, x_true, device=0, rows=None, tile_bins=0, cluster=0,
                  force_unfused=False, consumer_warps=0, variant=0):
        """On-device Philox/Poisson stack (include/sfhcuda.h: sfh_stack_create_synthetic)."""
        self = cls.__new__(cls)
        dt = np.dtype(dtype)
        self.shape = (int(nbins), int(ntemplates))
        self.dtype = dt
        o = L.sfh_opts()
        o.struct_size = C.sizeof(L.sfh_opts)
        o.device = device
        if rows is not None:
            o.row_begin, o.row_end = int(rows[0]), int(rows[1])
        o.tile_bins, o.cluster, o.force_unfused, o.consumer_warps = tile_bins, cluster, int(force_unfused), consumer_warps
        o.variant = variant
        x = np.ascontiguousarray(x_true, dtype=np.float64)
        if x.shape[0] != ntemplates:
            raise ValueError("len(x_true) != ntemplates")
        h = C.c_void_p()
        L.check(L.lib.sfh_stack_create_synthetic(C.byref(h), nbins, ntemplates, _DT[dt], C.c_uint64(seed), float(scale),
                                                 _dp(x), C.byref(o)))
        self._finish_init(h)
        self._data_ref = None
        return self

    @classmethod
    def from_file(cls, path, device=0, rows=None, verify=False, tile_bins=0, cluster=0, force_unfused=False, consumer_warps=0,
                  variant=0):
        """Upload a stack written by :meth:`save` (include/sfhcuda.h: sfh_stack_create_from_file).  ``rows`` selects the
        bin-row shard of this process; only those rows of the memory-mapped file are read.  ``verify`` recomputes the
        payload checksums first (reads the whole file).  ``logAge`` / ``MH`` / ``hess_shape`` are restored when stored."""
        from .io import KIND_STACK, SFHFile
        with SFHFile(path) as f:
            if f.kind != KIND_STACK or "models" not in f:
                raise ValueError(f"{path} is not a stack file")
            nb_total, fb, fe, nx, ny = f.attrs[:5]
            dm = f.describe("models")
            logAge = f.read("logAge") if "logAge" in f else None
            MH = f.read("MH") if "MH" in f else None
        self = cls.__new__(cls)
        self.shape = (int(nb_total), int(dm["shape"][1]))
        self.dtype = np.dtype(dm["dtype"])
        self.logAge, self.MH = logAge, MH
        self.hess_shape = (int(nx), int(ny)) if nx * ny else None
        o = L.sfh_opts()
        o.struct_size = C.sizeof(L.sfh_opts)
        o.device = device
        if rows is not None:
            o.row_begin, o.row_end = int(rows[0]), int(rows[1])
        o.tile_bins, o.cluster, o.force_unfused, o.consumer_warps = tile_bins, cluster, int(force_unfused), consumer_warps
        o.variant = variant
        h = C.c_void_p()
        L.check(L.lib.sfh_stack_create_from_file(C.byref(h), str(path).encode(), int(bool(verify)), C.byref(o)))
        self._finish_init(h)
        self._data_ref = None
        return self

    def save(self, path, logAge=None, MH=None, hess_shape=None):
        """Write this stack (the bin-row shard it holds) and its data to ``path`` (sfh_stack_save): the device copy goes
        straight into the memory-mapped file in column blocks, no host copy of the stack is made."""
        if (logAge is None) != (MH is None):
            raise ValueError("logAge and MH go together")
        la = np.ascontiguousarray(logAge, dtype=np.float64) if logAge is not None else None
        mh = np.ascontiguousarray(MH, dtype=np.float64) if MH is not None else None
        if la is not None and not (la.shape == mh.shape == (self.shape[1],)):
            raise ValueError("size(models,2) == length(logAge) == length(metallicities) must hold")
        nx, ny = (int(hess_shape[0]), int(hess_shape[1])) if hess_shape is not None else (0, 0)
        L.check(L.lib.sfh_stack_save(self.handle, str(path).encode(), nx, ny, _dp(la) if la is not None else None,
                                     _dp(mh) if mh is not None else None))

    @staticmethod
    def _destroy(h, ctxs):
        for c in ctxs:
            c.close()
        L.lib.sfh_stack_destroy(h)

    def close(self):
        self._fin()

    # -- plumbing -------------------------------------------------------------------------
    def ctx(self) -> _Ctx:
        c = getattr(self._tls, "ctx", None)
        if c is None:
            c = _Ctx(self.handle)
            self._tls.ctx = c
            self._ctxs.append(c)
        return c

    def new_ctx(self, stream=None) -> _Ctx:
        c = _Ctx(self.handle, stream)
        self._ctxs.append(c)
        return c

    def info(self) -> L.sfh_info:
        i = L.sfh_info()
        L.check(L.lib.sfh_stack_info(self.handle, C.byref(i)))
        return i

    def set_data(self, data):
        d = np.asarray(data).reshape(-1, order="F")
        if d.dtype not in _DT:
            d = d.astype(np.float64)
        d = np.ascontiguousarray(d)
        if d.shape[0] != self.shape[0]:
            raise ValueError("axes(models,1) != axes(data,1)")
        L.check(L.lib.sfh_stack_set_data(self.handle, d.ctypes.data_as(C.c_void_p), _DT[d.dtype]))
        self._data_ref = _data_ref(data)

    def download(self):
        i = self.info()
        rows = i.row_end - i.row_begin
        M = np.empty((rows, self.shape[1]), dtype=self.dtype, order="F")
        d = np.empty(rows, dtype=np.float64)
        L.check(L.lib.sfh_stack_download(self.handle, M.ctypes.data_as(C.c_void_p), _dp(d)))
        return M, d

    def download_data(self):
        """The observed Hess diagram bound to this stack (float64), without copying the templates back."""
        i = self.info()
        d = np.empty(i.row_end - i.row_begin, dtype=np.float64)
        L.check(L.lib.sfh_stack_download(self.handle, None, _dp(d)))
        return d

    # -- raw calls (all host-synchronous) ----------------------------------------------------
    def eval_fg(self, coeffs, want_F=True, want_G=True, want_composite=False):
        x = np.ascontiguousarray(coeffs, dtype=np.float64)
        if x.ndim != 1 or x.shape[0] != self.shape[1]:
            raise ValueError("axes(coeffs,1) != axes(models,2)")           # fitting_base.jl:59 / solvers.jl:10
        nl = C.c_double()
        G = np.empty(self.shape[1]) if want_G else None
        comp = np.empty(self.rows) if want_composite else None
        L.check(L.lib.sfh_eval_fg(self.ctx().handle, _dp(x), C.byref(nl) if want_F else None,
                                  _dp(G) if want_G else None, _dp(comp) if want_composite else None))
        return (nl.value if want_F else None), G, comp

    def eval_logl_batched(self, X):
        X = np.asarray(X, dtype=np.float64)
        if X.ndim == 1:
            X = X[:, None]
        if X.shape[0] != self.shape[1]:
            raise ValueError("size(X,1) != size(models,2)")
        X = np.asfortranarray(X)
        out = np.empty(X.shape[1])
        L.check(L.lib.sfh_eval_logl_batched(self.ctx().handle, _dp(X), X.shape[1], _dp(out)))
        return out

    def eval_fg_batched(self, X, want_G=True):
        """fg! for every column of X (T x C) in one device pass: returns (-logL[C], G[T, C] or None)."""
        X = np.asarray(X, dtype=np.float64)
        if X.ndim == 1:
            X = X[:, None]
        if X.shape[0] != self.shape[1]:
            raise ValueError("size(X,1) != size(models,2)")
        X = np.asfortranarray(X)
        nl = np.empty(X.shape[1])
        G = np.empty(X.shape, order="F") if want_G else None
        L.check(L.lib.sfh_eval_fg_batched(self.ctx().handle, _dp(X), X.shape[1], _dp(nl), _dp(G) if want_G else None))
        return nl, G

    def mcmc_run(self, X0, nsteps, nthin=1, a_scale=2.0, seed=0, store=True):
        """Device-resident stretch-move ensemble sampler (sfh_mcmc_run).  X0: (T, W) starting walkers, W even.
        Returns (chain (nsteps//nthin, T, W) or None, logl_chain (nsteps//nthin, W) or None, X_final, logl_final,
        acceptance fraction)."""
        X = np.array(X0, dtype=np.float64, order="F")
        if X.ndim != 2 or X.shape[0] != self.shape[1]:
            raise ValueError("length of each walker != number of templates")
        T, W = X.shape
        nstore = int(nsteps) // int(nthin)
        chain_buf = np.empty((nstore, W, T)) if store else None    # each stored step is T x W column-major = W rows of T
        lchain = np.empty((nstore, W)) if store else None
        lfin = np.empty(W)
        acc = C.c_double(0.0)
        L.check(L.lib.sfh_mcmc_run(self.ctx().handle, _dp(X), W, int(nsteps), int(nthin), float(a_scale), int(seed) & (2**64 - 1),
                                   _dp(chain_buf) if store else None, _dp(lchain) if store else None, _dp(lfin), C.byref(acc)))
        chain = chain_buf.transpose(0, 2, 1) if store else None      # view: (nstore, T, W)
        return chain, lchain, X, lfin, acc.value

    def column_sums(self):
        """colsum_j = sum_i M_ij of the resident stack (one device pass)."""
        out = np.empty(self.shape[1])
        L.check(L.lib.sfh_column_sums(self.ctx().handle, _dp(out)))
        return out

    def time_fg(self, coeffs, reps=10, want_G=True, flush_l2=True):
        x = np.ascontiguousarray(coeffs, dtype=np.float64)
        ms, msk = C.c_double(), C.c_double()
        L.check(L.lib.sfh_time_fg(self.ctx().handle, _dp(x), reps, int(want_G), int(flush_l2), C.byref(ms), C.byref(msk)))
        return ms.value, msk.value


class _GroupCtx:
    """The primary context of a multi-GPU group: owned by the group (never destroyed from here)."""

    def __init__(self, handle):
        self.handle = handle
        self.bound_key = None
        self.n_ages = 0

    def close(self):
        pass


class DeviceStackGroup(DeviceStack):
    """A template stack sharded by bin rows over several GPUs of THIS process (include/sfhcuda.h: sfh_group_*): what the
    reference's single-process callers (fit_sfh, fit_templates, ...) need for stacks that do not fit one GPU.  Drop-in for a
    :class:`DeviceStack` wherever the fused evaluation is used (``fg_``, the hierarchical ``fg_``, ``fit_templates*`` and
    ``fit_sfh`` with ``engine="native"``); the batched-walker / sampler paths raise ``SFHError`` (unsupported on a group)."""

    def __init__(self, models, data, devices=None, ndev=None, dtype=None, clamp_eps=0.0, tile_bins=0, cluster=0, variant=0):
        M = _as_stack_matrix(models)
        dt = np.dtype(dtype) if dtype is not None else M.dtype
        if dt not in (np.dtype(np.float32), np.dtype(np.float64)):
            dt = np.dtype(np.float64)
        M = np.asfortranarray(M, dtype=dt)
        d = np.asarray(data).reshape(-1, order="F")
        if d.dtype not in _DT:
            d = d.astype(np.float64)
        d = np.ascontiguousarray(d)
        if d.shape[0] != M.shape[0]:
            raise ValueError("axes(models,1) != axes(data,1)")
        devs, n = self._devices(devices, ndev)
        o = L.sfh_opts()
        o.struct_size = C.sizeof(L.sfh_opts)
        o.clamp_eps = clamp_eps
        o.tile_bins, o.cluster, o.variant = tile_bins, cluster, variant
        g = C.c_void_p()
        L.check(L.lib.sfh_group_create(C.byref(g), M.ctypes.data_as(C.c_void_p), M.shape[0], M.shape[1], _DT[dt],
                                       d.ctypes.data_as(C.c_void_p), _DT[d.dtype], devs, n, C.byref(o)))
        self._finish_group(g, M.shape, dt, n)
        self._data_ref = None      # the bound data cannot be re-bound on a group: pass the group itself on every call

    @staticmethod
    def _devices(devices, ndev):
        if devices is None:
            n = int(ndev) if ndev is not None else L.device_count()
            return None, n
        arr = (C.c_int * len(devices))(*[int(v) for v in devices])
        return arr, len(devices)

    def _finish_group(self, g, shape, dt, n):
        self.group = g
        self.shape = (int(shape[0]), int(shape[1]))
        self.dtype = dt
        self.ndev = n
        self.rows = self.shape[0]
        h = C.c_void_p()
        L.check(L.lib.sfh_group_ctx(g, C.byref(h)))
        self._primary = _GroupCtx(h)
        self.handle = None
        self._fin = weakref.finalize(self, L.lib.sfh_group_destroy, g)

    @classmethod
    def synthetic(cls, nbins, ntemplates, dtype, seed, scale, x_true, devices=None, ndev=None, tile_bins=0, cluster=0, variant=0):
        self = cls.__new__(cls)
        devs, n = cls._devices(devices, ndev)
        o = L.sfh_opts()
        o.struct_size = C.sizeof(L.sfh_opts)
        o.tile_bins, o.cluster, o.variant = tile_bins, cluster, variant
        x = np.ascontiguousarray(x_true, dtype=np.float64)
        if x.shape[0] != ntemplates:
            raise ValueError("len(x_true) != ntemplates")
        g = C.c_void_p()
        L.check(L.lib.sfh_group_create_synthetic(C.byref(g), int(nbins), int(ntemplates), _DT[np.dtype(dtype)], C.c_uint64(seed),
                                                 float(scale), _dp(x), devs, n, C.byref(o)))
        self._finish_group(g, (nbins, ntemplates), np.dtype(dtype), n)
        self._data_ref = None
        return self

    def close(self):
        self._fin()

    def ctx(self):
        return self._primary

    def new_ctx(self, stream=None):
        raise L.SFHError("a multi-GPU group has one (primary) context")

    def infos(self):
        n = C.c_int()
        arr = (L.sfh_info * self.ndev)()
        L.check(L.lib.sfh_group_info(self.group, C.byref(n), arr))
        return list(arr)

    def info(self):
        return self.infos()[0]

    def set_data(self, data):
        raise L.SFHError("re-binding data on a multi-GPU group is not supported: create a new group")

    def download(self):
        raise L.SFHError("a multi-GPU group does not gather its shards back")

    download_data = download

    def time_fg(self, coeffs, reps=10, want_G=True, flush_l2=False):
        """(ms per all-reduced evaluation, same) -- device-timed, max over the GPUs."""
        x = np.ascontiguousarray(coeffs, dtype=np.float64)
        ms = C.c_double()
        L.check(L.lib.sfh_group_time_fg(self.group, _dp(x), int(reps), int(want_G), C.byref(ms)))
        return ms.value, ms.value


# ---------------------------------------------------------------------------------------------
# identity-keyed cache for callers that pass bare arrays on every iteration
# ---------------------------------------------------------------------------------------------
_cache: dict = {}
_cache_lock = threading.Lock()


def _data_ref(data):
    """A reference that pins the IDENTITY of the bound data object (id() alone is recycled by the allocator once the
    original array is collected, so `DeviceStack(M, d1)` followed by a call with a fresh `d2` at the same address silently kept
    d1).  Weak when the object allows it, strong otherwise (lists, scalars); None when no data was bound by the caller."""
    if data is None:
        return None
    try:
        return weakref.ref(data)
    except TypeError:
        return lambda: data


def _weak_ids(obj):
    """Weak references that pin down the IDENTITY of `obj` (an ndarray or a list of ndarrays): id() alone can be
    recycled by the allocator once the original array is garbage collected."""
    items = list(obj) if isinstance(obj, (list, tuple)) else [obj]
    refs = []
    for it in items:
        try:
            refs.append(weakref.ref(it))
        except TypeError:
            return None  # not weak-referenceable (e.g. a nested list): never cache
    return refs


def _same(refs, obj):
    items = list(obj) if isinstance(obj, (list, tuple)) else [obj]
    return refs is not None and len(refs) == len(items) and all(r() is it for r, it in zip(refs, items))


def device_stack(models, data) -> DeviceStack:
    if isinstance(models, DeviceStack):
        # the stack re-binds `data` when the caller passes a DIFFERENT object than the one bound (identity through a weak
        # reference: a dead referent never compares equal).  In-place mutation of the bound array is NOT detected -- the reference
        # re-reads `data` on every call, a device stack cannot; call ds.set_data(data) after mutating it.
        if data is not None and models._data_ref is not None and models._data_ref() is not data:
            models.set_data(data)
        return models
    key = (id(models), id(data))
    with _cache_lock:
        ent = _cache.get(key)
        if ent is not None:
            ds, mrefs, drefs = ent
            if _same(mrefs, models) and _same(drefs, data):
                return ds
            _cache.pop(key)
            ds.close()
        if len(_cache) >= 8:
            _cache.pop(next(iter(_cache)))[0].close()
        ds = DeviceStack(models, data)
        mrefs, drefs = _weak_ids(models), _weak_ids(data)
        if mrefs is not None and drefs is not None:
            _cache[key] = (ds, mrefs, drefs)
        return ds


def clear_cache():
    with _cache_lock:
        for ds, _, _ in _cache.values():
            ds.close()
        _cache.clear()


# ---------------------------------------------------------------------------------------------
# the reference's functions
# ---------------------------------------------------------------------------------------------
def _flat_out(arr):
    """View of a user-supplied output array as a flat column-major vector we can fill in place."""
    a = np.asarray(arr)
    return a


def composite_(composite, coeffs, models, data=None):
    """``composite!(composite, coeffs, models)``: composite <- sum_j coeffs[j] * models[:, j].  Returns None.

    (``data`` is not part of the reference signature; it is only needed when ``models`` is a bare array
    that has not been seen before, because a DeviceStack binds data at creation.)"""
    if isinstance(models, DeviceStack):
        ds = models
    else:
        M = _as_stack_matrix(models)
        ds = device_stack(models, data if data is not None else np.zeros(M.shape[0]))
    comp = np.asarray(composite)
    if comp.size != ds.shape[0]:
        raise ValueError("axes(composite,1) != axes(models,1)")            # fitting_base.jl:58
    x = np.ascontiguousarray(coeffs, dtype=np.float64)
    if x.shape[0] != ds.shape[1]:
        raise ValueError("axes(coeffs,1) != axes(models,2)")               # fitting_base.jl:59
    out = np.empty(ds.rows)
    L.check(L.lib.sfh_composite(ds.ctx().handle, _dp(x), _dp(out)))
    composite[...] = out.reshape(comp.shape, order="F").astype(comp.dtype, copy=False)
    return None


def loglikelihood(*args):
    """``loglikelihood(composite, data)`` or ``loglikelihood(coeffs, models, data)``.

    The two-argument form needs the device copy of ``data``: pass a DeviceStack as ``data`` or use the
    three-argument form.  Returned scalar has the promoted element type (fitting_core_test.jl:40)."""
    if len(args) == 2:
        composite, data = args
        if not isinstance(data, DeviceStack):
            raise TypeError("loglikelihood(composite, data): `data` must be the DeviceStack holding the observed "
                            "Hess diagram (there is no host implementation)")
        comp = np.ascontiguousarray(np.asarray(composite).reshape(-1, order="F"), dtype=np.float64)
        if comp.shape[0] != data.rows:
            raise ValueError("axes(composite) != axes(data)")              # fitting_base.jl:85
        out = C.c_double()
        L.check(L.lib.sfh_loglikelihood(data.ctx().handle, _dp(comp), C.byref(out)))
        T = np.promote_types(np.asarray(composite).dtype, data.dtype) if np.asarray(composite).dtype.kind == "f" else data.dtype
        return T.type(out.value)
    coeffs, models, data = args
    ds = device_stack(models, data)
    x = np.ascontiguousarray(coeffs, dtype=np.float64)
    if x.shape[0] != ds.shape[1]:
        raise ValueError("axes(coeffs,1) != axes(models,2)")               # fitting_base.jl:120
    out = C.c_double()
    L.check(L.lib.sfh_loglikelihood_coeffs(ds.ctx().handle, _dp(x), C.byref(out)))
    return ds.dtype.type(out.value)


def grad_loglikelihood_(G, composite, models, data=None):
    """``grad-loglikelihood!(G, composite, models, data)``: G <- -M' (1 - n/max(composite,eps)); ``composite`` is
    overwritten with the residual (documented side effect, fitting_base.jl:219).  Returns G."""
    ds = device_stack(models, data)
    comp = np.asarray(composite)
    if comp.size != ds.rows:
        raise ValueError("axes(models,1) != axes(composite,1)")            # fitting_base.jl:272
    if np.asarray(G).shape[0] != ds.shape[1]:
        raise ValueError("axes(G,1) != axes(models,2)")                    # fitting_base.jl:271
    buf = np.ascontiguousarray(comp.reshape(-1, order="F"), dtype=np.float64).copy()
    g = np.empty(ds.shape[1])
    L.check(L.lib.sfh_grad_loglikelihood(ds.ctx().handle, _dp(buf), _dp(g)))
    composite[...] = buf.reshape(comp.shape, order="F").astype(comp.dtype, copy=False)
    G[...] = g.astype(np.asarray(G).dtype, copy=False)
    return G


def _grad_single(model, composite, data):
    """``grad-loglikelihood(model, composite, data)`` for ONE template (fitting_base.jl:144-160): the scalar
    ``sum_i -c_i (1 - n_i / max(m_i, eps(T)))`` with ``T = promote_type(eltype(model), eltype(composite), eltype(data))``.
    On the device this is ``sfh_grad_loglikelihood`` on a one-column stack."""
    m, c, d = np.asarray(model), np.asarray(composite), np.asarray(data)
    if not (m.shape == c.shape == d.shape):
        raise ValueError("axes(model) == axes(composite) == axes(data) must hold")      # fitting_base.jl:147
    T = np.result_type(*[a.dtype if a.dtype.kind == "f" else np.float32 for a in (m, c, d)])   # integers promote to the float type
    if T not in (np.dtype(np.float32), np.dtype(np.float64)):
        T = np.dtype(np.float64)
    ds = DeviceStack(np.asfortranarray(m.reshape(-1, 1, order="F"), dtype=T), d.reshape(-1, order="F"), dtype=T)
    try:
        buf = np.ascontiguousarray(c.reshape(-1, order="F"), dtype=np.float64).copy()
        g = np.empty(1)
        L.check(L.lib.sfh_grad_loglikelihood(ds.ctx().handle, _dp(buf), _dp(g)))
    finally:
        ds.close()
    return T.type(g[0])


def grad_loglikelihood(*args):
    """``grad-loglikelihood(model, composite, data)`` -> scalar (one template, fitting_base.jl:144-160);
    ``(models, composite, data)`` / ``(coeffs, models, data)`` -> new gradient vector (:171-182, :193-211)."""
    a, b, data = args
    if not isinstance(a, (DeviceStack, list, tuple)) and not isinstance(b, (DeviceStack, list, tuple)):
        sa, sb = np.shape(a), np.shape(b)
        if sa == sb and len(sa) >= 1:
            # same axes for the first two arguments: the single-template method (a coefficient vector never has the stack's shape,
            # and a stack never has the composite's).  Mismatched `data` raises like the reference's @argcheck.
            return _grad_single(a, b, data)
    if isinstance(b, DeviceStack) or (not isinstance(a, DeviceStack) and np.asarray(a).ndim == 1 and not isinstance(a, (list, tuple))):
        coeffs, models = a, b                                              # (coeffs, models, data)  :193-211
        ds = device_stack(models, data)
        _, G, _ = ds.eval_fg(coeffs, want_F=False, want_G=True)
        return (-G).astype(ds.dtype, copy=False)
    models, composite = a, b                                               # (models, composite, data) :171-182
    ds = device_stack(models, data)
    G = np.empty(ds.shape[1], dtype=ds.dtype)
    comp = np.array(np.asarray(composite).reshape(-1, order="F"), dtype=np.float64)
    return grad_loglikelihood_(G, comp, ds, data)


def fg_(F, G, coeffs, models, data, composite=None):
    """``fg!(F, G, coeffs, models, data, composite)`` -- returns -logL when ``F is not None``; fills ``G`` with the
    gradient of -logL when ``G is not None``; with both ``None`` only the composite is formed (solvers.jl:20-38).
    ``composite`` (optional scratch, as in the reference) receives what the reference leaves in it."""
    ds = device_stack(models, data)
    want_F, want_G = F is not None, G is not None
    if want_G and np.asarray(G).shape[0] != ds.shape[1]:
        raise ValueError("axes(G,1) != axes(models,2)")
    if composite is not None and np.asarray(composite).size != ds.rows:
        raise ValueError("axes(models,1) != axes(composite,1)")            # solvers.jl:11
    nl, g, comp = ds.eval_fg(coeffs, want_F=want_F, want_G=want_G, want_composite=composite is not None)
    if want_G:
        G[...] = g.astype(np.asarray(G).dtype, copy=False)
    if composite is not None:
        ca = np.asarray(composite)
        composite[...] = comp.reshape(ca.shape, order="F").astype(ca.dtype, copy=False)
    if want_F:
        return ds.dtype.type(nl)
    return None
