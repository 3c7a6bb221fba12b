"""On-disk container for template stacks and fit results.

The reference has no file format of its own (its notebook pickles the template matrices with ``Serialization``,
examples/fitting1.ipynb cell 96); this is the binding of the one ``libsfhcuda.so`` defines (include/sfhcuda.h,
``sfh_file_*``; byte layout in csrc/sfh_file.h): a memory-mappable table of named column-major arrays with per-array
checksums.  File handling is plain host I/O and works without a GPU; moving a stack between a file and the device
(``DeviceStack.save`` / ``DeviceStack.from_file``) goes through ``sfh_stack_save`` / ``sfh_stack_create_from_file``.

    write_arrays(path, {"name": array, ...}, kind=..., attrs=...)     sfh_file_write
    SFHFile(path)  -> .names, .kind, .attrs, [name] (zero-copy views), .verify()
    read_arrays(path, verify=True) -> dict of copies
    save_result / load_result       fit results (mu, sigma, invH, minimiser, model parameters) and sampler chains
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib as L

KIND_GENERIC, KIND_STACK, KIND_RESULT = 0, 1, 2
_NP2SFH = {np.dtype(np.float32): L.SFH_F32, np.dtype(np.float64): L.SFH_F64, np.dtype(np.int64): L.SFH_I64,
           np.dtype(np.uint8): L.SFH_U8}
_SFH2NP = {v: k for k, v in _NP2SFH.items()}


def checksum64(a) -> int:
    """The container's order-independent 64-bit checksum of an array's bytes (sfh_checksum64)."""
    a = np.ascontiguousarray(a)
    out = C.c_uint64(0)
    L.check(L.lib.sfh_checksum64(a.ctypes.data_as(C.c_void_p), a.nbytes, C.byref(out)))
    return out.value


def write_arrays(path, arrays, kind=KIND_GENERIC, attrs=None):
    """Write ``arrays`` (mapping name -> ndarray of float32/float64/int64/uint8, at most 4-d; stored column-major) to
    ``path`` atomically.  ``attrs``: up to 8 integers kept in the header."""
    names = list(arrays)
    n = len(names)
    descs = (L.sfh_array_desc * max(n, 1))()
    ptrs = (C.c_void_p * max(n, 1))()
    keep = []
    for i, name in enumerate(names):
        a = np.asarray(arrays[name])
        if a.dtype not in _NP2SFH:
            if a.dtype.kind in "iub":
                a = a.astype(np.int64)
            elif a.dtype.kind == "f":
                a = a.astype(np.float64)
            else:
                raise ValueError(f"array '{name}': dtype {a.dtype} cannot be stored")
        if a.ndim == 0:
            a = a.reshape(1)
        if a.ndim > 4:
            raise ValueError(f"array '{name}': more than 4 dimensions")
        a = np.asfortranarray(a)
        keep.append(a)
        b = name.encode("utf-8")
        if not 0 < len(b) < 48:
            raise ValueError(f"array name '{name}' must be 1..47 bytes")
        descs[i].name = b
        descs[i].dtype = _NP2SFH[a.dtype]
        descs[i].ndim = a.ndim
        for k in range(4):
            descs[i].dims[k] = a.shape[k] if k < a.ndim else 1
        ptrs[i] = a.ctypes.data
    at = (C.c_int64 * 8)(*([int(v) for v in attrs] + [0] * (8 - len(attrs)))) if attrs is not None else None
    L.check(L.lib.sfh_file_write(str(path).encode(), int(kind), at, n, descs, ptrs))


class SFHFile:
    """Read-only mapping of a container file.  ``f[name]`` is a zero-copy, read-only, column-major view that is valid
    until ``close()``; ``f.read(name)`` copies."""

    _h = None

    def __init__(self, path):
        h = C.c_void_p()
        L.check(L.lib.sfh_file_open(str(path).encode(), C.byref(h)))
        self._h = h
        kind, n = C.c_int(), C.c_int()
        at = (C.c_int64 * 8)()
        L.check(L.lib.sfh_file_info(h, C.byref(kind), C.byref(n), at))
        self.kind, self.attrs = kind.value, [int(v) for v in at]
        self._descs = []
        for i in range(n.value):
            d = L.sfh_array_desc()
            L.check(L.lib.sfh_file_array(h, i, C.byref(d), None))
            self._descs.append(d)
        self.names = [d.name.decode("utf-8") for d in self._descs]

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def close(self):
        if self._h is not None:
            L.lib.sfh_file_close(self._h)
            self._h = None

    def __del__(self):
        self.close()

    def __contains__(self, name):
        return name in self.names

    def index(self, name) -> int:
        i = C.c_int(-1)
        L.check(L.lib.sfh_file_find(self._h, name.encode("utf-8"), C.byref(i)))
        if i.value < 0:
            raise KeyError(name)
        return i.value

    def describe(self, name):
        d = self._descs[self.index(name)]
        return {"dtype": _SFH2NP[d.dtype], "shape": tuple(d.dims[k] for k in range(d.ndim)), "nbytes": d.nbytes, "checksum": d.checksum}

    def __getitem__(self, name):
        i = self.index(name)
        d = self._descs[i]
        p = C.c_void_p()
        L.check(L.lib.sfh_file_array(self._h, i, None, C.byref(p)))
        shape = tuple(d.dims[k] for k in range(d.ndim))
        dt = _SFH2NP[d.dtype]
        if d.nbytes == 0:
            a = np.empty(shape, dtype=dt, order="F")
            a.flags.writeable = False
            return a
        buf = (C.c_char * d.nbytes).from_address(p.value)
        buf._owner = self                       # the mapping must outlive the view
        a = np.frombuffer(buf, dtype=dt).reshape(shape, order="F")
        a.flags.writeable = False
        return a

    def read(self, name):
        return np.array(self[name], order="F")

    def verify(self, name=None):
        """Recompute payload checksums (all arrays, or one); raises SFHError(SFH_ERR_IO) on a mismatch."""
        L.check(L.lib.sfh_file_verify(self._h, -1 if name is None else self.index(name)))
        return True


def read_arrays(path, verify=True):
    with SFHFile(path) as f:
        if verify:
            f.verify()
        return {n: f.read(n) for n in f.names}


# ---------------------------------------------------------------------------------------------
# fit results and chains
# ---------------------------------------------------------------------------------------------
def _model_record(prefix, m, out):
    out[prefix + "_params"] = np.asarray(list(m.fittable_params()), dtype=np.float64)
    out[prefix + "_free"] = np.asarray(list(m.free_params()), dtype=np.uint8)
    out[prefix + "_class"] = np.frombuffer(type(m).__name__.encode(), dtype=np.uint8)
    fixed = getattr(m, "fixed", None)          # logMstar0 / T_max, solZ, Y_p, gamma (the C-ABI's mh_fixed)
    if callable(fixed):
        out[prefix + "_fixed"] = np.asarray(list(fixed()), dtype=np.float64)


def save_result(path, result):
    """Store what a driver returned: the ``{"map", "mle"}`` pair of BFGSResult / LogTransformFTResult objects
    (fit_sfh, fit_templates; bfgs_result.jl:25-41, solvers.jl:115-129), a single such object, or a dict of arrays
    (e.g. the ``posterior_matrix`` / ``logp`` of sample_sfh, or ``{"chain": ..., "logl": ...}`` of mcmc_sample)."""
    arrays = {}

    def one(prefix, r):
        if isinstance(r, dict):
            for k, v in r.items():
                if k == "result":
                    arrays[prefix + "x"] = np.asarray(v.x, dtype=np.float64)
                elif isinstance(v, (np.ndarray, list, tuple, float, int)):
                    arrays[prefix + k] = np.asarray(v)
            return
        arrays[prefix + "mu"] = np.asarray(r.mu, dtype=np.float64)
        arrays[prefix + "sigma"] = np.asarray(r.sigma, dtype=np.float64)
        arrays[prefix + "invH"] = np.asarray(r.invH, dtype=np.float64)
        if getattr(r, "result", None) is not None and hasattr(r.result, "x"):
            arrays[prefix + "x"] = np.asarray(r.result.x, dtype=np.float64)
        if getattr(r, "MH_model", None) is not None:
            _model_record(prefix + "MH", r.MH_model, arrays)
            _model_record(prefix + "disp", r.disp_model, arrays)

    if isinstance(result, dict) and set(result) >= {"map", "mle"} and not isinstance(result["map"], np.ndarray):
        one("map/", result["map"])
        one("mle/", result["mle"])
    else:
        one("", result)
    write_arrays(path, arrays, kind=KIND_RESULT)


def load_result(path, verify=True):
    """Inverse of :func:`save_result`: a dict of arrays, nested one level by the ``map/`` / ``mle/`` prefixes; the
    ``*_class`` records come back as strings."""
    flat = read_arrays(path, verify=verify)
    out: dict = {}
    for k, v in flat.items():
        if k.endswith("_class"):
            v = bytes(v).decode()
        if "/" in k:
            a, b = k.split("/", 1)
            out.setdefault(a, {})[b] = v
        else:
            out[k] = v
    return out
