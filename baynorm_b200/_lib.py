"""ctypes binding of libbaynorm_b200.so (the C ABI declared in include/baynorm_b200.h).

The shared library is the product; this module only loads it and declares argument
types.  There is no fallback: if the library is missing or no sm_100 GPU is usable,
calls raise -- nothing here ever computes on the CPU.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

_HERE = Path(__file__).resolve().parent
LIB_PATH = _HERE / "libbaynorm_b200.so"

BN_OK = 0
ERR_NAMES = {1: "BN_ERR_INVALID", 2: "BN_ERR_CUDA", 3: "BN_ERR_NO_DEVICE", 4: "BN_ERR_OOM", 5: "BN_ERR_BETA_RANGE",
             6: "BN_ERR_NONCOUNT", 7: "BN_ERR_COLLECTIVE", 8: "BN_ERR_UNSUPPORTED"}


class BayNormError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("%s (%d): %s" % (ERR_NAMES.get(code, "BN_ERR"), code, msg))
        self.code = code


class FitOpts(C.Structure):
    _fields_ = [("method", C.c_int32), ("maxfeval", C.c_int32), ("maxit", C.c_int32), ("M", C.c_int32),
                ("ftol", C.c_double), ("gtol", C.c_double), ("eps", C.c_double), ("xatol", C.c_double),
                ("use_gradient", C.c_int32), ("park_after", C.c_int32)]


class PriorOut(C.Structure):
    _fields_ = [("MME_MU", C.c_void_p), ("MME_SIZE", C.c_void_p), ("BB_SIZE", C.c_void_p),
                ("MME_SIZE_adjust", C.c_void_p), ("conv", C.c_void_p), ("nfeval", C.c_void_p),
                ("size_lower", C.c_double), ("size_upper", C.c_double), ("coef", C.c_double * 3),
                ("seconds", C.c_double * 4)]


ALLREDUCE_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_void_p)
TILE_SINK = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_int64, C.c_int64, C.c_void_p)
GROUP_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_int, C.c_void_p)

# name -> (restype, argtypes); every symbol include/baynorm_b200.h declares
_VP, _I64, _I32, _D, _U64 = C.c_void_p, C.c_int64, C.c_int, C.c_double, C.c_uint64
SIGNATURES = {
    "bn_version": (C.c_int, []),
    "bn_ctx_create": (C.c_int, [C.c_int, C.POINTER(_VP)]),
    "bn_ctx_destroy": (None, [_VP]),
    "bn_last_error": (C.c_char_p, [_VP]),
    "bn_ctx_synchronize": (C.c_int, [_VP]),
    "bn_ctx_stream": (_VP, [_VP]),
    "bn_ctx_launch_count": (_I64, [_VP]),
    "bn_ctx_device_info": (C.c_int, [_VP, C.POINTER(_I64)]),
    "bn_ctx_set_allreduce": (C.c_int, [_VP, ALLREDUCE_FN, _VP, C.c_int, C.c_int]),
    "bn_mat_from_csc": (C.c_int, [_VP, _I64, _I64, _VP, C.c_int, _VP, _VP, C.POINTER(_VP)]),
    "bn_mat_from_dense": (C.c_int, [_VP, _I64, _I64, _VP, C.POINTER(_VP)]),
    "bn_mat_free": (None, [_VP]),
    "bn_mat_info": (C.c_int, [_VP, C.POINTER(_I64)]),
    "bn_mat_set_origin": (C.c_int, [_VP, _I64, _I64]),
    "bn_mat_export_csc": (C.c_int, [_VP, _VP, _VP, _VP, _VP]),
    "bn_mat_select_cols": (C.c_int, [_VP, _VP, _VP, _I64, C.POINTER(_VP)]),
    "bn_mat_synth_nb": (C.c_int, [_VP, _I64, _I64, _VP, _VP, _VP, _U64, _I64, C.POINTER(_VP)]),
    "bn_colsums": (C.c_int, [_VP, _VP, _VP]),
    "bn_beta_default": (C.c_int, [_VP, _VP, _D, _VP]),
    "bn_mme": (C.c_int, [_VP, _VP, _VP, _VP, _VP, _VP]),
    "bn_row_means": (C.c_int, [_VP, _VP, _VP]),
    "bn_row_vars": (C.c_int, [_VP, _VP, _VP, _VP]),
    "bn_mme_dense": (C.c_int, [_VP, _I64, _I64, _VP, _VP, _VP, _VP]),
    "bn_marginal_nb_1d": (C.c_int, [_VP, _D, _D, _VP, _VP, _I64, _VP]),
    "bn_marginal_nb_2d": (C.c_int, [_VP, _VP, _VP, _VP, _I64, _VP]),
    "bn_gradient_nb_1d": (C.c_int, [_VP, _D, _D, _VP, _VP, _I64, _VP]),
    "bn_gradient_nb_2d": (C.c_int, [_VP, _VP, _VP, _VP, _I64, _VP]),
    "bn_fit_opts_default": (None, [C.POINTER(FitOpts)]),
    "bn_fit_size": (C.c_int, [_VP, _VP, _VP, _VP, _VP, _D, _D, C.POINTER(FitOpts), _VP, _VP, _VP]),
    "bn_fit_size_mu": (C.c_int, [_VP, _VP, _VP, _VP, _VP, _D, _D, _D, _D, C.POINTER(FitOpts), _VP, _VP, _VP, _VP]),
    "bn_vec_range": (C.c_int, [_VP, _VP, _I64, _VP]),
    "bn_adjust_size": (C.c_int, [_VP, _I64, _VP, _VP, _VP, _VP]),
    "bn_prior": (C.c_int, [_VP, _VP, _VP, C.c_int, C.POINTER(FitOpts), C.POINTER(PriorOut)]),
    "bn_post_mean": (C.c_int, [_VP, _VP, _VP, _VP, _VP, _I64, _I64, _VP]),
    "bn_post_mode": (C.c_int, [_VP, _VP, _VP, _VP, _VP, _I64, _I64, _VP]),
    "bn_post_sample": (C.c_int, [_VP, _VP, _VP, _VP, _VP, _I64, _I64, C.c_int, _U64, _VP]),
    "bn_downsample": (C.c_int, [_VP, _I64, _I64, _VP, _VP, _U64, _VP]),
    "bn_debug_philox": (C.c_int, [_VP, _VP, _VP, _VP]),
    "bn_debug_log": (C.c_int, [_VP, C.c_int, _VP, _I64, _VP]),
    "bn_mat_drop_gene_major": (C.c_int, [_VP]),
    "bn_synthetic_control": (C.c_int, [_VP, _VP, _VP, _U64, _VP, _VP]),
    "bn_beta_fun": (C.c_int, [_VP, _VP, _D, _VP, _VP, _VP]),
    "bn_debug_fp64_peak": (C.c_int, [_VP, C.POINTER(C.c_double)]),
    "bn_mme_raw": (C.c_int, [_VP, _VP, _VP, _VP, _VP, _VP]),
    "bn_post_mean_csc": (C.c_int, [_VP, _VP, _VP, _VP, _VP, _I64, _I64, C.POINTER(_VP)]),
    "bn_csc_result_info": (C.c_int, [_VP, C.POINTER(_I64)]),
    "bn_csc_result_fetch": (C.c_int, [_VP, _VP, _VP, _VP, _VP, _VP]),
    "bn_csc_result_to_mat": (C.c_int, [_VP, _VP, C.POINTER(_VP)]),
    "bn_csc_result_free": (None, [_VP]),
    "bn_post_stream": (C.c_int, [_VP, _VP, C.c_int, _VP, _VP, _VP, C.c_int, _U64, _I64, TILE_SINK, _VP]),
    "bn_group_create": (C.c_int, [C.c_int, _VP, C.POINTER(_VP)]),
    "bn_group_destroy": (None, [_VP]),
    "bn_group_last_error": (C.c_char_p, [_VP]),
    "bn_group_size": (C.c_int, [_VP]),
    "bn_group_ctx": (_VP, [_VP, C.c_int]),
    "bn_group_run": (C.c_int, [_VP, GROUP_FN, _VP]),
    "bn_group_mat_from_csc": (C.c_int, [_VP, _I64, _I64, _VP, C.c_int, _VP, _VP, C.POINTER(_VP)]),
    "bn_gmat_free": (None, [_VP]),
    "bn_gmat_bounds": (C.c_int, [_VP, _VP]),
    "bn_gmat_shard": (_VP, [_VP, C.c_int]),
    "bn_group_beta_default": (C.c_int, [_VP, _VP, _D, _VP]),
    "bn_group_prior": (C.c_int, [_VP, _VP, _VP, C.c_int, C.POINTER(FitOpts), C.POINTER(PriorOut)]),
    "bn_group_post": (C.c_int, [_VP, _VP, C.c_int, _VP, _VP, _VP, _I64, _I64, C.c_int, _U64, _VP]),
    "bn_debug_last_fit": (C.c_int, [_VP, C.POINTER(_I64)]),
    "bn_debug_dense_table": (C.c_int, [_VP, _VP, C.c_int64, C.c_double, C.c_double, _VP, C.c_int64, _VP]),
    "bn_debug_tune": (C.c_int, [_VP, C.c_int, C.c_int]),
    "bn_debug_write_peak": (C.c_int, [_VP, _I64, C.c_int, C.c_int, C.POINTER(C.c_double)]),
}

_lib = None


def load():
    """Load the CUDA library or raise.  Never substitutes a CPU implementation."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise ImportError("%s is missing: build it with `make -C baynorm_b200/csrc` or __graft_entry__.build(); "
                          "baynorm_b200 has no CPU fallback" % LIB_PATH)
    lib = C.CDLL(str(LIB_PATH), mode=getattr(os, "RTLD_LOCAL", 0))
    for name, (res, args) in SIGNATURES.items():
        f = getattr(lib, name)          # AttributeError if the .so does not export it
        f.restype = res
        f.argtypes = args
    _lib = lib
    return lib
