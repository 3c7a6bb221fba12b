"""ctypes binding of ``libsfhcuda.so`` -- one prototype per symbol declared in ``include/sfhcuda.h``.

The library is loaded at import time from THIS directory (built in-tree by ``csrc/Makefile`` /
``__graft_entry__.build()``).  If it is missing the import fails loudly: there is no Python, numpy
or CPU implementation of the hot path behind this package.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SFH_LIB") or os.path.join(_HERE, "libsfhcuda.so")   # SFH_LIB: A/B experiment builds

if not os.path.exists(LIB_PATH):
    raise ImportError(
        f"{LIB_PATH} not found: build it with `make -C {os.path.join(_HERE, 'csrc')}` "
        "(or __graft_entry__.build()).  There is no CPU fallback for the fitting hot path."
    )

lib = C.CDLL(LIB_PATH)

# ---- enums (include/sfhcuda.h) -------------------------------------------------------------
SFH_OK, SFH_ERR_INVALID_ARG, SFH_ERR_SHAPE, SFH_ERR_NO_DEVICE, SFH_ERR_CUDA = 0, 1, 2, 3, 4
SFH_ERR_OOM, SFH_ERR_NCCL, SFH_ERR_UNSUPPORTED, SFH_ERR_NOT_BOUND, SFH_ERR_IO = 5, 6, 7, 8, 9
SFH_F32, SFH_F64, SFH_I64, SFH_U8 = 0, 1, 2, 3
SFH_MH_POWERLAW_MZR, SFH_MH_LINEAR_AMR, SFH_MH_LOG_AMR = 0, 1, 2
SFH_DISP_GAUSSIAN = 0


class sfh_opts(C.Structure):
    _fields_ = [("struct_size", C.c_int32), ("device", C.c_int32), ("row_begin", C.c_int64), ("row_end", C.c_int64),
                ("clamp_eps", C.c_double), ("tile_bins", C.c_int32), ("cluster", C.c_int32),
                ("force_unfused", C.c_int32), ("consumer_warps", C.c_int32), ("variant", C.c_int32), ("reserved", C.c_int32)]


class sfh_info(C.Structure):
    _fields_ = [("nbins_total", C.c_int64), ("ntemplates", C.c_int64), ("row_begin", C.c_int64), ("row_end", C.c_int64),
                ("ld", C.c_int64), ("dtype", C.c_int32), ("device", C.c_int32), ("fused", C.c_int32),
                ("tile_bins", C.c_int32), ("cluster", C.c_int32), ("chunks_per_tile", C.c_int32),
                ("ring_slots", C.c_int32), ("n_clusters", C.c_int32), ("consumer_warps", C.c_int32), ("sm_count", C.c_int32),
                ("cc_major", C.c_int32), ("cc_minor", C.c_int32), ("variant", C.c_int32), ("panel_layout", C.c_int32), ("l2_resident_mb", C.c_int32),
                ("stack_bytes", C.c_int64), ("clamp_eps", C.c_double)]


class sfh_stats(C.Structure):
    _fields_ = [("evals", C.c_int64), ("kernel_launches", C.c_int64), ("last_device_ms", C.c_double)]


class sfh_array_desc(C.Structure):
    _fields_ = [("name", C.c_char * 48), ("dtype", C.c_int32), ("ndim", C.c_int32), ("dims", C.c_int64 * 4),
                ("nbytes", C.c_int64), ("checksum", C.c_uint64)]


class sfh_bfgs_opts(C.Structure):
    _fields_ = [("struct_size", C.c_int32), ("alphaguess", C.c_int32), ("g_abstol", C.c_double), ("maxiter", C.c_int64),
                ("device_hessian", C.c_int32), ("reserved", C.c_int32)]


class sfh_bfgs_report(C.Structure):
    _fields_ = [("f", C.c_double), ("g_norm", C.c_double), ("iterations", C.c_int64), ("f_calls", C.c_int64),
                ("converged", C.c_int32), ("status", C.c_int32)]


class sfh_lbfgsb_opts(C.Structure):
    _fields_ = [("struct_size", C.c_int32), ("m", C.c_int32), ("factr", C.c_double), ("pgtol", C.c_double), ("maxiter", C.c_int64),
                ("maxfun", C.c_int64)]


class sfh_lbfgsb_report(C.Structure):
    _fields_ = [("f", C.c_double), ("pg_norm", C.c_double), ("iterations", C.c_int64), ("f_calls", C.c_int64), ("status", C.c_int32),
                ("reserved", C.c_int32)]


class sfh_nuts_opts(C.Structure):
    _fields_ = [("struct_size", C.c_int32), ("max_depth", C.c_int32), ("nwarmup", C.c_int64), ("delta", C.c_double),
                ("eps0", C.c_double), ("seed", C.c_uint64), ("mass_kind", C.c_int32), ("reserved", C.c_int32)]


sfh_batch_logdensity_fn = C.CFUNCTYPE(C.c_int, C.c_void_p, C.POINTER(C.c_double), C.c_int64, C.c_int64, C.POINTER(C.c_double),
                                      C.POINTER(C.c_double))
sfh_objective_fn = C.CFUNCTYPE(C.c_int, C.c_void_p, C.POINTER(C.c_double), C.c_int64, C.POINTER(C.c_double), C.POINTER(C.c_double))
SFH_FIT_LOG_MAP, SFH_FIT_LOG_MLE, SFH_FIT_SQRT_MLE = 0, 1, 2

_vp, _i64, _int, _dp = C.c_void_p, C.c_int64, C.c_int, C.POINTER(C.c_double)
_u8p = C.POINTER(C.c_uint8)

# every exported symbol of include/sfhcuda.h: name -> (restype, argtypes)
PROTOTYPES = {
    "sfh_version": (_int, []),
    "sfh_last_error": (C.c_char_p, []),
    "sfh_device_count": (_int, [C.POINTER(_int)]),
    "sfh_stack_create": (_int, [C.POINTER(_vp), _vp, _i64, _i64, _int, _vp, _int, C.POINTER(sfh_opts)]),
    "sfh_stack_create_synthetic": (_int, [C.POINTER(_vp), _i64, _i64, _int, C.c_uint64, C.c_double, _dp, C.POINTER(sfh_opts)]),
    "sfh_stack_create_from_points": (_int, [C.POINTER(_vp), _i64, _i64, C.c_double, C.c_double, C.c_double, C.c_double, _i64,
                                           C.POINTER(_i64), _dp, _dp, _dp, _dp, _dp, C.POINTER(C.c_int32), _int, _vp, _int, C.POINTER(sfh_opts)]),
    "sfh_stack_destroy": (_int, [_vp]),
    "sfh_stack_info": (_int, [_vp, C.POINTER(sfh_info)]),
    "sfh_stack_set_data": (_int, [_vp, _vp, _int]),
    "sfh_stack_download": (_int, [_vp, _vp, _dp]),
    "sfh_ctx_create": (_int, [_vp, _vp, C.POINTER(_vp)]),
    "sfh_ctx_destroy": (_int, [_vp]),
    "sfh_ctx_stats": (_int, [_vp, C.POINTER(sfh_stats)]),
    "sfh_eval_fg": (_int, [_vp, _dp, _dp, _dp, _dp]),
    "sfh_composite": (_int, [_vp, _dp, _dp]),
    "sfh_loglikelihood": (_int, [_vp, _dp, _dp]),
    "sfh_loglikelihood_coeffs": (_int, [_vp, _dp, _dp]),
    "sfh_grad_loglikelihood": (_int, [_vp, _dp, _dp]),
    "sfh_column_sums": (_int, [_vp, _dp]),
    "sfh_hier_bind": (_int, [_vp, _dp, _dp, C.POINTER(_i64)]),
    "sfh_calculate_coeffs": (_int, [_vp, _int, _dp, _int, _dp, _dp]),
    "sfh_eval_fg_hier": (_int, [_vp, _int, _dp, _int, _dp, _u8p, _dp, _dp]),
    "sfh_eval_logl_batched": (_int, [_vp, _dp, _i64, _dp]),
    "sfh_eval_fg_batched": (_int, [_vp, _dp, _i64, _dp, _dp]),
    "sfh_eval_fg_hier_batched": (_int, [_vp, _int, _dp, _int, _dp, _i64, _u8p, _dp, _dp]),
    "sfh_mcmc_run": (_int, [_vp, _dp, _i64, _i64, _i64, C.c_double, C.c_uint64, _dp, _dp, _dp, _dp]),
    "sfh_comm_unique_id": (_int, [_vp]),
    "sfh_comm_init": (_int, [_vp, _int, _int, _vp]),
    "sfh_comm_p2p_handle": (_int, [_vp, _int, _vp]),
    "sfh_comm_p2p_init": (_int, [_vp, _int, _int, _vp]),
    "sfh_shard_rows": (_int, [_i64, _int, _int, _i64, C.POINTER(_i64), C.POINTER(_i64)]),
    "sfh_group_create": (_int, [C.POINTER(_vp), _vp, _i64, _i64, _int, _vp, _int, C.POINTER(_int), _int, C.POINTER(sfh_opts)]),
    "sfh_group_create_synthetic": (_int, [C.POINTER(_vp), _i64, _i64, _int, C.c_uint64, C.c_double, _dp, C.POINTER(_int), _int, C.POINTER(sfh_opts)]),
    "sfh_group_destroy": (_int, [_vp]),
    "sfh_group_ctx": (_int, [_vp, C.POINTER(_vp)]),
    "sfh_group_info": (_int, [_vp, C.POINTER(_int), C.POINTER(sfh_info)]),
    "sfh_group_time_fg": (_int, [_vp, _dp, _int, _int, _dp]),
    "sfh_comm_p2p_enable": (_int, [_vp, _int]),
    "sfh_ctx_comm_info": (_int, [_vp, C.POINTER(_int), C.POINTER(_int), C.POINTER(_int)]),
    "sfh_enqueue_fg": (_int, [_vp, _vp, _vp, _int]),
    "sfh_enqueue_logl_batched": (_int, [_vp, _vp, _i64, _vp]),
    "sfh_ctx_synchronize": (_int, [_vp]),
    "sfh_time_fg": (_int, [_vp, _dp, _int, _int, _int, _dp, _dp]),
    "sfh_minimize_bfgs": (_int, [sfh_objective_fn, _vp, _i64, _dp, C.POINTER(sfh_bfgs_opts), C.POINTER(sfh_bfgs_report), _dp]),
    "sfh_fit_templates_bfgs": (_int, [_vp, _int, _dp, C.POINTER(sfh_bfgs_opts), C.POINTER(sfh_bfgs_report), _dp]),
    "sfh_fit_fixed_amr_bfgs": (_int, [_vp, _dp, C.POINTER(C.c_int32), _i64, _int, _dp, C.POINTER(sfh_bfgs_opts),
                                      C.POINTER(sfh_bfgs_report), _dp]),
    "sfh_fit_sfh_bfgs": (_int, [_vp, _int, _dp, _int, _dp, C.POINTER(C.c_int32), _u8p, _int, _dp, C.POINTER(sfh_bfgs_opts),
                                C.POINTER(sfh_bfgs_report), _dp]),
    "sfh_fit_sfh_bfgs_generic": (_int, [sfh_objective_fn, _vp, _i64, C.c_int32, _dp, C.POINTER(C.c_int32), _u8p, _int, _dp,
                                        C.POINTER(sfh_bfgs_opts), C.POINTER(sfh_bfgs_report), _dp]),
    "sfh_nuts_run": (_int, [sfh_batch_logdensity_fn, _vp, _i64, _i64, _dp, C.POINTER(_i64), _dp, C.POINTER(sfh_nuts_opts), _dp, _dp, _dp,
                            C.POINTER(_i64), C.POINTER(_i64)]),
    "sfh_hmc_sample_nuts": (_int, [_vp, _i64, _dp, C.POINTER(_i64), _dp, C.POINTER(sfh_nuts_opts), _dp, _dp, _dp, C.POINTER(_i64),
                                   C.POINTER(_i64)]),
    "sfh_sample_sfh_nuts": (_int, [_vp, _int, _dp, _int, _dp, C.POINTER(C.c_int32), _u8p, _i64, _dp, C.POINTER(_i64), _dp,
                                   C.POINTER(sfh_nuts_opts), _dp, _dp, _dp, C.POINTER(_i64), C.POINTER(_i64)]),
    "sfh_sample_sfh_nuts_generic": (_int, [sfh_batch_logdensity_fn, _vp, _i64, C.c_int32, _dp, C.POINTER(C.c_int32), _u8p, _i64, _dp,
                                           C.POINTER(_i64), _dp, C.POINTER(sfh_nuts_opts), _dp, _dp, _dp, C.POINTER(_i64), C.POINTER(_i64)]),
    "sfh_minimize_lbfgsb": (_int, [sfh_objective_fn, _vp, _i64, _dp, _dp, _dp, C.POINTER(sfh_lbfgsb_opts), C.POINTER(sfh_lbfgsb_report)]),
    "sfh_fit_templates_lbfgsb": (_int, [_vp, _dp, C.POINTER(sfh_lbfgsb_opts), C.POINTER(sfh_lbfgsb_report)]),
    "sfh_checksum64": (_int, [_vp, _i64, C.POINTER(C.c_uint64)]),
    "sfh_file_write": (_int, [C.c_char_p, _int, C.POINTER(_i64), _int, C.POINTER(sfh_array_desc), C.POINTER(_vp)]),
    "sfh_file_open": (_int, [C.c_char_p, C.POINTER(_vp)]),
    "sfh_file_close": (_int, [_vp]),
    "sfh_file_info": (_int, [_vp, C.POINTER(_int), C.POINTER(_int), C.POINTER(_i64)]),
    "sfh_file_find": (_int, [_vp, C.c_char_p, C.POINTER(_int)]),
    "sfh_file_array": (_int, [_vp, _int, C.POINTER(sfh_array_desc), C.POINTER(_vp)]),
    "sfh_file_verify": (_int, [_vp, _int]),
    "sfh_stack_save": (_int, [_vp, C.c_char_p, _i64, _i64, _dp, _dp]),
    "sfh_stack_create_from_file": (_int, [C.POINTER(_vp), C.c_char_p, _int, C.POINTER(sfh_opts)]),
}

for _name, (_res, _args) in PROTOTYPES.items():
    _f = getattr(lib, _name)  # AttributeError here == the .so does not export what the header declares
    _f.restype = _res
    _f.argtypes = _args


class SFHError(RuntimeError):
    def __init__(self, status: int, message: str):
        super().__init__(f"libsfhcuda status {status}: {message}")
        self.status = status


def check(status: int) -> None:
    """Map a non-OK status to the exception the reference would raise (ArgumentError -> ValueError)."""
    if status == SFH_OK:
        return
    msg = (lib.sfh_last_error() or b"").decode("utf-8", "replace")
    if status in (SFH_ERR_SHAPE, SFH_ERR_INVALID_ARG):
        raise ValueError(f"libsfhcuda: {msg}")
    raise SFHError(status, msg)


def device_count() -> int:
    n = C.c_int(0)
    check(lib.sfh_device_count(C.byref(n)))
    return n.value
