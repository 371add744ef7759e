"""ctypes binding of libccgp.so (include/ccgp.h).  No fallback: if the library
is missing or no B200 is usable, the error propagates."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("CCGP_LIB_PATH") or os.path.join(_HERE, "lib", "libccgp.so")   # (override: A/B builds during kernel work)

# every symbol include/ccgp.h declares (tests check that the .so exports them all)
SYMBOLS = [
    "ccgp_num_params", "ccgp_create", "ccgp_destroy", "ccgp_last_error", "ccgp_sync", "ccgp_set_matern_nu", "ccgp_set_stream", "ccgp_use_own_stream", "ccgp_device",
    "ccgp_launch_count", "ccgp_last_nll_config", "ccgp_measure_fp64_peak", "ccgp_set_design",
    "ccgp_nll_batch", "ccgp_nll_batch_dev", "ccgp_argmin_dev", "ccgp_nll_argmin", "ccgp_rinv_batch",
    "ccgp_predict", "ccgp_predict_dev", "ccgp_me_schur_batch", "ccgp_me_schur_batch_dev", "ccgp_me_argmin",
    "ccgp_subset_logdet_batch", "ccgp_subset_logdet_batch_dev", "ccgp_mixed_corr", "ccgp_debug_phase_timing",
    "ccgp_kmedoids_pam", "ccgp_me_schur_paired", "ccgp_me_schur_stencil", "ccgp_rcond_batch",
    "ccgp_create_multi", "ccgp_num_gpus", "ccgp_collective_count", "ccgp_measure_fp64_peak_dmma",
    "ccgp_cgp_objective_batch", "ccgp_cgp_jackknife",
    "ccgp_factors_create", "ccgp_factors_predict", "ccgp_factors_predict_dev", "ccgp_factors_info", "ccgp_factors_destroy",
]

_lib = None


class CcgpError(RuntimeError):
    pass


def load():
    """dlopen libccgp.so once and declare the prototypes."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise CcgpError("libccgp.so not built (%s); run `python -c 'import __graft_entry__ as g; g.build()'`" % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    vp, dp, ip = C.c_void_p, C.c_void_p, C.c_void_p   # raw addresses (host numpy or device pointers)
    i32, i64, f64 = C.c_int, C.c_int64, C.c_double
    P = C.POINTER
    lib.ccgp_num_params.argtypes = [i32, i32]
    lib.ccgp_create.argtypes = [P(vp), i32]
    lib.ccgp_destroy.argtypes = [vp]
    lib.ccgp_create_multi.argtypes = [P(vp), i32]
    lib.ccgp_num_gpus.argtypes = [vp]
    lib.ccgp_collective_count.argtypes = [vp]
    lib.ccgp_collective_count.restype = i64
    lib.ccgp_last_error.argtypes = [vp]
    lib.ccgp_last_error.restype = C.c_char_p
    lib.ccgp_sync.argtypes = [vp]
    lib.ccgp_set_matern_nu.argtypes = [vp, f64]
    lib.ccgp_set_stream.argtypes = [vp, vp]
    lib.ccgp_use_own_stream.argtypes = [vp]
    lib.ccgp_device.argtypes = [vp]
    lib.ccgp_launch_count.argtypes = [vp]
    lib.ccgp_launch_count.restype = i64
    lib.ccgp_last_nll_config.argtypes = [vp, P(i32), P(i32), P(i32), P(i32)]
    lib.ccgp_measure_fp64_peak.argtypes = [vp, P(f64)]
    lib.ccgp_measure_fp64_peak_dmma.argtypes = [vp, P(f64)]
    lib.ccgp_set_design.argtypes = [vp, dp, i32, i32, dp]
    nll = [vp, i32, i32, dp, i64, i64, f64, i32, f64, dp, dp, ip]
    lib.ccgp_nll_batch.argtypes = nll
    lib.ccgp_nll_batch_dev.argtypes = nll
    lib.ccgp_argmin_dev.argtypes = [vp, dp, i64, P(f64), P(i64)]
    lib.ccgp_nll_argmin.argtypes = [vp, i32, i32, dp, i64, i64, f64, i32, f64, P(f64), P(i64)]
    lib.ccgp_rinv_batch.argtypes = [vp, i32, i32, dp, i64, i64, dp, dp, ip]
    lib.ccgp_rcond_batch.argtypes = [vp, i32, i32, dp, i64, i64, dp, dp, ip]
    pred = [vp, i32, dp, i64, i64, i32, dp, i64, dp, i64, f64, dp, dp, ip]
    lib.ccgp_predict.argtypes = pred
    lib.ccgp_predict_dev.argtypes = pred
    lib.ccgp_factors_create.argtypes = [vp, i32, dp, i64, i64, i32, dp, i64, C.POINTER(C.c_void_p)]
    lib.ccgp_factors_predict.argtypes = [vp, vp, dp, i64, f64, dp, dp, ip]
    lib.ccgp_factors_predict_dev.argtypes = [vp, vp, dp, i64, f64, dp, dp, ip]
    lib.ccgp_factors_info.argtypes = [vp, vp, C.POINTER(C.c_int64), C.POINTER(C.c_int), C.POINTER(C.c_int64)]
    lib.ccgp_factors_destroy.argtypes = [vp, vp]
    me = [vp, dp, i32, i32, dp, i32, i64, dp, i64, i64, dp, dp, ip]
    lib.ccgp_me_schur_batch.argtypes = me
    lib.ccgp_me_schur_batch_dev.argtypes = me
    lib.ccgp_me_argmin.argtypes = [vp, dp, i32, i32, dp, i32, i64, dp, i64, i64, dp, dp]
    sub = [vp, dp, i64, i32, ip, i32, i64, i64, i32, dp, dp, ip]
    lib.ccgp_subset_logdet_batch.argtypes = sub
    lib.ccgp_subset_logdet_batch_dev.argtypes = sub
    lib.ccgp_mixed_corr.argtypes = [vp, i32, dp, dp, i32, dp, i32, i32, dp]
    lib.ccgp_debug_phase_timing.argtypes = [vp, i32, vp]
    lib.ccgp_me_schur_paired.argtypes = [vp, dp, i32, i32, dp, i32, i64, dp, i64, i64, dp, ip]
    lib.ccgp_me_schur_stencil.argtypes = [vp, dp, i32, i32, dp, i32, i64, dp, i64, i64, f64, f64, f64, dp, ip]
    lib.ccgp_kmedoids_pam.argtypes = [vp, dp, i64, i32, i32, i32, ip, dp, ip]
    lib.ccgp_cgp_objective_batch.argtypes = [vp, dp, dp, i32, i32, dp, i64, i64, dp, ip]
    lib.ccgp_cgp_jackknife.argtypes = [vp, dp, dp, i32, i32, dp, dp, ip]
    for s in SYMBOLS:
        fn = getattr(lib, s)
        if s not in ("ccgp_last_error", "ccgp_launch_count", "ccgp_collective_count"):
            fn.restype = i32
    _lib = lib
    return lib
