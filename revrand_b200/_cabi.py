"""ctypes binding of the C-ABI in ``include/revrand_b200.h``.

The product path has no CPU fallback: if ``librevrand_b200.so`` is missing or
a CUDA device is unavailable, every compute entry point raises.  Loading the
library and resolving its symbols needs no GPU (the CPU test-suite checks
that every declared symbol is exported).
"""

from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "librevrand_b200.so")

RR_ENGINE_AUTO, RR_ENGINE_SIMT, RR_ENGINE_TCGEN05, RR_ENGINE_TCGEN05_FINE = 0, 1, 2, 3
RR_ENGINE_TCGEN05_FUSED16 = 4
RR_OP_SUFFSTATS, RR_OP_GRADPASS, RR_OP_PREDICT = 1, 2, 3
RR_OP_GLM_STEP, RR_OP_GLM_PREDICT, RR_OP_RESIDUAL = 4, 5, 6
RR_OP_GRADPASS_KEPT = 7
RR_GRAD_SPLIT_C = 0x100
(RR_LIK_GAUSSIAN, RR_LIK_BERNOULLI, RR_LIK_BINOMIAL, RR_LIK_POISSON_EXP,
 RR_LIK_POISSON_SOFTPLUS) = range(5)


class RRPlan(C.Structure):
    """Mirror of ``struct rr_plan``."""
    _fields_ = [
        ("d", C.c_int32), ("ktot", C.c_int32), ("next", C.c_int32),
        ("D", C.c_int32),
        ("Wt", C.c_void_p), ("amp", C.c_void_p),
        ("col_cos", C.c_void_p), ("col_sin", C.c_void_p),
        ("ext_src", C.c_void_p), ("ext_val", C.c_void_p),
        ("ext_col", C.c_void_p),
        ("kind", C.c_void_p),
        ("ext_pow", C.c_void_p),
        ("col_scale", C.c_void_p),
    ]


_P = C.c_void_p
_I32, _I64, _F32, _SZ = C.c_int32, C.c_int64, C.c_float, C.c_size_t
_PLAN = C.POINTER(RRPlan)

# name -> (restype, argtypes); must list every symbol of the public header.
SIGNATURES = {
    "rr_version": (C.c_int, []),
    "rr_last_error": (C.c_char_p, []),
    "rr_launch_count": (C.c_uint64, []),
    "rr_device_info": (C.c_int, [C.POINTER(_I32)] * 3),
    "rr_features": (C.c_int, [_PLAN, _P, _I64, _P, _I64, _P]),
    "rr_trig_grad": (C.c_int, [_P, _I64, _I32, _P, _I32, _P, _I32, _I32, _P, _P]),
    "rr_fastfood_features": (C.c_int, [_P, _I64, _I32, _I32, _I32, _P, _P, _P,
                                       _P, _P, _P, _P]),
    "rr_centre_features": (C.c_int, [_P, _I64, _I32, _P, _I32, _P, _I32, _I32, _P, _P, _P]),
    "rr_gm_grad": (C.c_int, [_P, _I64, _I32, _P, _I32, _P, _P, _P, _P, _P]),
    "rr_context_create": (C.c_int, [C.POINTER(_P)]),
    "rr_context_destroy": (C.c_int, [_P]),
    "rr_engine_auto_min_rows": (_I64, []),
    "rr_slm_suffstats": (C.c_int, [_PLAN, _P, _P, _I64, _P, _P, _P, _P, _SZ,
                                   _I32, _P, _P]),
    "rr_slm_residual": (C.c_int, [_PLAN, _P, _P, _I64, _P, _P, _P, _P, _SZ, _P]),
    "rr_slm_gradpass": (C.c_int, [_PLAN, _P, _P, _I64, _P, _P, _P, _P, _P, _SZ,
                                  _I32, _P, _P]),
    "rr_slm_kept_features_bytes": (_SZ, [_PLAN, _I64]),
    "rr_slm_suffstats_keep": (C.c_int, [_PLAN, _P, _P, _I64, _P, _P, _P, _P, _SZ, _P, _SZ,
                                        _P, _P]),
    "rr_slm_gradpass_kept": (C.c_int, [_PLAN, _P, _P, _I64, _P, _P, _P, _P, _P, _SZ, _P, _SZ,
                                       _I32, _P]),
    "rr_slm_predict": (C.c_int, [_PLAN, _P, _I64, _P, _P, _P, _P, _P, _SZ, _P]),
    "rr_glm_step": (C.c_int, [_PLAN, _P, _P, _P, _I64, _P, _P, _I32, _P, _I32,
                              _I32, _F32, _P, _P, _P, _P, _P, _P, _SZ, _P]),
    "rr_glm_predict": (C.c_int, [_PLAN, _P, _I64, _P, _I32, _I32, _F32, _P, _P,
                                 _P, _P, _SZ, _P]),
    "rr_glm_cdf": (C.c_int, [_P, _I64, _I32, _I32, _F32, _P, C.c_double, _P, _P, _P, _P]),
    "rr_glm_quantiles": (C.c_int, [_P, _I64, _I32, _I32, _F32, _P, C.c_double,
                                   C.c_double, _P, _P, _P]),
    "rr_workspace_bytes": (_SZ, [_I32, _I64, _I32, _I32, _I32, _I32, _I32, _I32]),
    "rr_tcgen05_supported": (C.c_int, [_I32, _I32, _I32, _I32]),
    "rr_tcgen05_selftest": (C.c_int, [C.POINTER(C.c_double)]),
    "rr_tcgen05_i8_selftest": (C.c_int, [_I32, C.POINTER(_I64)]),
    "rr_tcgen05_gemm3": (C.c_int, [_I32, _I32, _I32, _F32, _P, _I64, _I32, _P, _I64, _I32, _P,
                                   _I64, _I32, _P, _SZ, _P]),
    "rr_tcgen05_accum_probe": (C.c_int, [_P, _P, _P, _P, _I32, _I32, _P, _P]),
}

_lib = None


class RevrandB200Error(RuntimeError):
    """Raised when the native library is missing or a call fails."""


def load():
    """Load the shared library (no GPU needed) and declare signatures."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RevrandB200Error(
            "native library %s not found: build it with "
            "`python -c 'import __graft_entry__ as g; g.build()'` "
            "(there is no CPU fallback)" % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is missing
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc, what):
    if rc != 0:
        msg = load().rr_last_error()
        raise RevrandB200Error("%s failed (status %d): %s"
                               % (what, rc, msg.decode() if msg else ""))
