"""ctypes binding of libjitr_b200.so (the C ABI declared in include/jitr_b200.h).

The product path has NO CPU fallback: if the CUDA library is missing, or no sm_100 device is
present when a compute entry point is called, this module raises.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# JITR_B200_LIB: load another build of the same library (kernel tuning experiments)
LIB_PATH = os.environ.get("JITR_B200_LIB") or os.path.join(_HERE, "_lib", "libjitr_b200.so")

OK, ERR_BAD_ARG, ERR_UNSUPPORTED, ERR_CUDA = 0, 1, 2, 3
STATUS_OK, STATUS_NONFINITE, STATUS_SINGULAR = 0, 1, 2
OPT_SYMMETRIC = 0
OPT_ASSUME_SYMMETRIC_KERNEL = 1
OPT_PACKED_FIRST_PANEL = 2
OPT_OBSERVABLES_DMMA = 3
MODEL_SHAPE, MODEL_KDUQ, MODEL_CHUQ, MODEL_LOCAL, MODEL_WLH, MODEL_SHAPE_PAIR, MODEL_DOM = 0, 1, 2, 3, 4, 5, 6
MODEL_NPARAMS = {MODEL_SHAPE: 17, MODEL_KDUQ: 40, MODEL_CHUQ: 22, MODEL_LOCAL: 15, MODEL_WLH: 46, MODEL_SHAPE_PAIR: 18, MODEL_DOM: 19}
# per-energy table layout (include/jitr_b200.h)
ET_A, ET_ECM, ET_K, ET_ELAB, ET_RCH, ET_AT, ET_ZT, ET_ZP, ET_AP, ET_RSCALE = range(10)
ET_A13, ET_AM13, ET_AM23, ET_AM53 = 10, 11, 12, 13
ET_EF = 14
ET_STRIDE = 16

_p = C.c_void_p
_i = C.c_int
_ll = C.c_longlong
_d = C.c_double

# every symbol include/jitr_b200.h declares: name -> (restype, argtypes)
SIGNATURES = {
    "jitr_b200_last_error": (C.c_char_p, []),
    "jitr_b200_version": (_i, []),
    "jitr_b200_device_ok": (_i, []),
    "jitr_b200_launch_count": (_ll, []),
    "jitr_b200_set_option": (_i, [_i, _i]),
    "jitr_b200_sym_fallback_count": (_ll, []),
    "jitr_b200_debug_set_fallback_mask": (_i, [_p]),
    "jitr_b200_elastic_sweep_params": (_i, [_i, _i, _i, _ll, _i, _p, _p, _p, _p, _p, _p, _i, _d, _p, _p, _p, _p]),
    "jitr_b200_elastic_sweep_tensors": (_i, [_i, _i, _ll, _i, _p, _p, _p, _p, _p, _p, _p, _p, _d, _p, _p, _p, _p]),
    "jitr_b200_eval_potentials": (_i, [_i, _i, _ll, _i, _p, _p, _p, _i, _p, _p, _p, _p]),
    "jitr_b200_elastic_observables": (_i, [_i, _i, _ll, _i, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p]),
    "jitr_b200_elastic_misfit": (_i, [_i, _i, _ll, _i] + [_p] * 14),
    "jitr_b200_percentiles": (_i, [_p, _ll, _ll, _p, _i, _p, _p]),
    "jitr_b200_qe_overlap": (_i, [_i, _ll, _i, _p, _p, _p, _p, _d, _d, _d, _p, _ll, _ll, _p]),
    "jitr_b200_qe_xs": (_i, [_i, _i, _ll, _p, _p, _d, _p, _p]),
    "jitr_b200_perey_buck_kernels": (_i, [_i, _i, _ll, _p, _p, _d, _p, _ll, _ll, _p]),
    "jitr_b200_solve_workspace_bytes": (_ll, [_i, _i, _ll, _i]),
    "jitr_b200_solve_batched": (_i, [_i, _i, _ll, _p, _ll, _p, _p, _p, _p, _p, _i, _p, _p, _ll, _p, _p, _p, _p, _p, _p, _p, _p]),
    "jitr_b200_solve_batched_indexed": (_i, [_i, _i, _ll, _p, _i, _p, _p, _p, _p, _p, _p, _i, _p, _p, _ll, _p, _p, _p, _p, _p, _p, _p, _p]),
    "jitr_b200_fp64_probe": (_i, [_i, _i, _i, C.POINTER(_d), C.POINTER(_ll)]),
}

_lib = None


class JitrB200Error(RuntimeError):
    pass


def lib():
    """Load the shared library (once).  Raises if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise JitrB200Error(
                f"{LIB_PATH} not found: build it with `python -m jitr_b200._build` "
                "(jitr_b200 has no CPU fallback)"
            )
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def check(code, what=""):
    if code == OK:
        return
    msg = lib().jitr_b200_last_error().decode()
    if code == ERR_BAD_ARG:
        raise ValueError(f"{what}: {msg}")
    if code == ERR_UNSUPPORTED:
        raise NotImplementedError(f"{what}: {msg}")
    raise JitrB200Error(f"{what}: {msg}")


def require_device():
    import torch

    if not torch.cuda.is_available() or not lib().jitr_b200_device_ok():
        raise JitrB200Error("jitr_b200 needs an sm_100 (B200) CUDA device; there is no CPU fallback")


def ptr(t):
    """Device pointer of a torch tensor (None -> NULL)."""
    if t is None:
        return None
    return C.c_void_p(t.data_ptr())


def stream_ptr():
    import torch

    return C.c_void_p(torch.cuda.current_stream().cuda_stream)
