"""ctypes binding of ``liblpgp.so`` (the C-ABI boundary declared in ``include/lpgp.h``).

There is no CPU fallback: if the shared library is missing the import fails loudly, and every compute entry
point requires CUDA tensors.
"""
from __future__ import annotations

import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "liblpgp.so")

MAX_DIM = 4
MAX_COEF = 512
LEAF = 128
MAX_SEG = 64
DIM_MATERN, DIM_EXPQUAD, DIM_RADIAL = 0, 1, 2
RADIAL_NQ = 6
GRAM_FULL, GRAM_LOWER = 0, 1
OPT_DIRECT_EXP = 1
OPT_NO_LOOKAHEAD = 2
OPT_TRSM_REFINE = 3


class KernelDesc(ctypes.Structure):
    _fields_ = [
        ("d", ctypes.c_int32),
        ("dim_type", ctypes.c_int32 * MAX_DIM),
        ("nbasis", ctypes.c_int32 * MAX_DIM),
        ("has_odd", ctypes.c_int32 * MAX_DIM),
        ("reserved", ctypes.c_int32),
        ("scale", ctypes.c_double * MAX_DIM),
        ("diag_value", ctypes.c_double),
        ("coef", ctypes.c_double * MAX_COEF),
    ]


class Factor(ctypes.Structure):
    _fields_ = [
        ("L", ctypes.c_void_p),
        ("n", ctypes.c_int64),
        ("ld", ctypes.c_int64),
        ("dinv", ctypes.c_void_p),
        ("nseg", ctypes.c_int32),
        ("reserved", ctypes.c_int32),
        ("seg_off", ctypes.c_int64 * (MAX_SEG + 1)),
    ]


class ObsBlock(ctypes.Structure):
    _fields_ = [
        ("desc", ctypes.POINTER(KernelDesc)),
        ("X", ctypes.c_void_p),
        ("n", ctypes.c_int64),
        ("col_off", ctypes.c_int64),
    ]


class OzakiPlanes(ctypes.Structure):
    _fields_ = [
        ("planes", ctypes.c_void_p),
        ("exps", ctypes.c_void_p),
        ("rows", ctypes.c_int64),
        ("cols", ctypes.c_int64),
        ("pitch", ctypes.c_int64),
        ("plane_stride", ctypes.c_int64),
        ("lde", ctypes.c_int64),
        ("nslices", ctypes.c_int32),
        ("kblock", ctypes.c_int32),
    ]


OZAKI_MAX_SLICES = 7
MAX_INTEGRAL_COEF = 8


class MaternIntegralDesc(ctypes.Structure):
    _fields_ = [
        ("ncoef", ctypes.c_int32),
        ("reserved", ctypes.c_int32),
        ("scale", ctypes.c_double),
        ("poly1", ctypes.c_double * MAX_INTEGRAL_COEF),
        ("poly2", ctypes.c_double * MAX_INTEGRAL_COEF),
    ]


def _load() -> ctypes.CDLL:
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing. Build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a). linpde_gp_b200 has no CPU fallback."
        )
    lib = ctypes.CDLL(LIB_PATH)
    i64, dbl, vp, ci = ctypes.c_int64, ctypes.c_double, ctypes.c_void_p, ctypes.c_int
    KD, FP, OB = ctypes.POINTER(KernelDesc), ctypes.POINTER(Factor), ctypes.POINTER(ObsBlock)
    OP = ctypes.POINTER(OzakiPlanes)
    sigs = {
        "lpgp_version": (ci, []),
        "lpgp_build_arch": (ctypes.c_char_p, []),
        "lpgp_error_string": (ctypes.c_char_p, [ci]),
        "lpgp_launch_count": (ctypes.c_longlong, [ci]),
        "lpgp_set_option": (ci, [ci, ci]),
        "lpgp_dmma_peak_probe": (ci, [vp, ci, ci, ctypes.POINTER(dbl), vp]),
        "lpgp_gram": (ci, [KD, vp, i64, vp, i64, vp, i64, ci, ci, dbl, vp]),
        "lpgp_gram_pairs": (ci, [KD, vp, vp, i64, vp, dbl, vp]),
        "lpgp_gram_diag": (ci, [KD, i64, vp, dbl, vp]),
        "lpgp_add_diag": (ci, [vp, i64, i64, vp, dbl, vp]),
        "lpgp_symmetrize_lower": (ci, [vp, i64, i64, vp]),
        "lpgp_kron_sum": (ci, [ci, ctypes.POINTER(vp), ctypes.POINTER(i64), ctypes.POINTER(vp), ctypes.POINTER(i64),
                          ctypes.POINTER(dbl), i64, i64, i64, i64, vp, i64, ci, ci, vp]),
        "lpgp_gemm_nt": (ci, [i64, i64, i64, dbl, vp, i64, vp, i64, dbl, vp, i64, ci, vp]),
        "lpgp_gemm_nn": (ci, [i64, i64, i64, dbl, vp, i64, vp, i64, dbl, vp, i64, vp]),
        "lpgp_gemm_nt_limited": (ci, [i64, i64, i64, dbl, vp, i64, vp, i64, dbl, vp, i64, vp, i64, vp]),
        "lpgp_factor_dinv_bytes": (ctypes.c_size_t, [ctypes.POINTER(i64), ci]),
        "lpgp_potrf": (ci, [FP, vp]),
        "lpgp_potrf_async": (ci, [FP, vp]),
        "lpgp_chol_append": (ci, [FP, vp]),
        "lpgp_trsm_rlt": (ci, [FP, i64, vp, i64, i64, vp]),
        "lpgp_trsm_rlt_refined": (ci, [FP, i64, vp, i64, i64, vp]),
        "lpgp_trsm_rln": (ci, [FP, vp, i64, i64, vp]),
        "lpgp_ozaki_split": (ci, [vp, i64, i64, i64, i64, i64, OP, ci, vp]),
        "lpgp_ozaki_gemm_nt": (ci, [i64, i64, i64, dbl, OP, i64, i64, OP, i64, i64, dbl, vp, i64, vp]),
        "lpgp_trsm_rlt_ozaki": (ci, [FP, vp, i64, i64, OP, OP, vp]),
        "lpgp_ozaki_gemm_stats": (ci, [ci, ctypes.POINTER(dbl), ctypes.POINTER(dbl), ctypes.POINTER(dbl), ctypes.POINTER(ctypes.c_longlong)]),
        "lpgp_i8_peak_probe": (ci, [ci, ci, ctypes.POINTER(dbl), vp]),
        "lpgp_potrs": (ci, [FP, vp, i64, i64, vp]),
        "lpgp_trsv": (ci, [FP, ci, vp, vp]),
        "lpgp_gemv": (ci, [ci, i64, i64, dbl, vp, i64, vp, vp, vp]),
        "lpgp_logdet": (ci, [FP, vp, vp]),
        "lpgp_post_mean": (ci, [OB, ci, vp, vp, i64, vp, ci, vp]),
        "lpgp_crosscov": (ci, [OB, ci, i64, vp, i64, vp, i64, vp]),
        "lpgp_post_var": (ci, [OB, ci, FP, vp, i64, dbl, vp, i64, vp, vp]),
        "lpgp_row_sumsq": (ci, [vp, i64, i64, i64, dbl, dbl, vp, vp]),
        "lpgp_matern_integral": (ci, [ctypes.POINTER(MaternIntegralDesc), dbl, dbl, vp, i64, dbl, vp, vp, i64, ci, vp]),
        "lpgp_matern_hat_integral": (ci, [ctypes.POINTER(MaternIntegralDesc), vp, i64, ci, vp, i64, dbl, vp, i64, ci, vp]),
        "lpgp_matern_integral2": (ci, [ctypes.POINTER(MaternIntegralDesc), dbl, dbl, dbl, dbl, dbl, vp, ci, vp]),
    }
    for name, (res, args) in sigs.items():
        fn = getattr(lib, name)  # AttributeError if the library does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    return lib


lib = _load()
EXPORTED = (
    "lpgp_version lpgp_build_arch lpgp_error_string lpgp_launch_count lpgp_set_option lpgp_dmma_peak_probe lpgp_gram lpgp_gram_pairs lpgp_gram_diag lpgp_add_diag lpgp_symmetrize_lower lpgp_kron_sum "
    "lpgp_gemm_nt lpgp_gemm_nn lpgp_gemm_nt_limited lpgp_ozaki_split lpgp_ozaki_gemm_nt lpgp_ozaki_gemm_stats lpgp_i8_peak_probe lpgp_trsm_rlt_ozaki lpgp_factor_dinv_bytes lpgp_potrf lpgp_potrf_async lpgp_chol_append lpgp_trsm_rlt lpgp_trsm_rlt_refined lpgp_trsm_rln lpgp_potrs lpgp_trsv lpgp_gemv lpgp_logdet "
    "lpgp_post_mean lpgp_crosscov lpgp_post_var lpgp_row_sumsq lpgp_matern_integral lpgp_matern_integral2 lpgp_matern_hat_integral"
).split()


def check(rc: int, what: str = "liblpgp call") -> None:
    """Map the C-ABI return convention onto the reference's exception types."""
    if rc == 0:
        return
    if rc > 0:
        raise np.linalg.LinAlgError(f"{what}: {rc}-th leading minor of the array is not positive definite")
    if rc <= -1000:
        raise RuntimeError(f"{what}: CUDA error {-rc - 1000}: {lib.lpgp_error_string(rc).decode()}")
    raise ValueError(f"{what}: invalid argument #{-rc}")
