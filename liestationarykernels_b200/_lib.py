"""ctypes binding of liblsk.so — the C-ABI boundary (include/lsk.h).

There is no CPU fallback: if the shared object is missing, or no CUDA device is present, every compute entry
point raises.  The library is loaded from this directory only (built in-tree by `build.py`).
"""
from __future__ import annotations

import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "liblsk.so")

MAX_FORMS = 5
MAX_VARS = 5
MAX_COEF = 400
MAX_ROWS = 24

EPI_CHEB1, EPI_STAIR2, EPI_ORBIT2, EPI_SPARSE, EPI_LINEAR, EPI_FEAT_CHEB1, EPI_FEAT_SPARSE = 1, 2, 3, 4, 5, 6, 7
EPI_HORNER, EPI_FEAT_HORNER, EPI_SYRK = 8, 9, 10
OUT_ROW_MAJOR, OUT_PANEL_MAJOR, FEAT_COMPLEX = 0, 1, 2


class LskEpilogue(ctypes.Structure):
    """Mirror of `struct LskEpilogue` (include/lsk.h)."""
    _fields_ = [
        ("kind", ctypes.c_int32),
        ("nforms", ctypes.c_int32),
        ("nvars", ctypes.c_int32),
        ("nterms", ctypes.c_int32),
        ("group_begin", ctypes.c_int32 * MAX_FORMS),
        ("group_end", ctypes.c_int32 * MAX_FORMS),
        ("row_len", ctypes.c_int32 * MAX_ROWS),
        ("out_layout", ctypes.c_int32),
        ("out_kg", ctypes.c_int32),
        ("feat_stride", ctypes.c_int32),
        ("reserved", ctypes.c_int32),
        ("scale", ctypes.c_double),
        ("stair_shape", ctypes.c_uint64),
        ("aff", (ctypes.c_double * (1 + MAX_FORMS)) * MAX_VARS),
        ("coef", ctypes.c_double * MAX_COEF),
    ]


class LskError(RuntimeError):
    pass


_lib = None

_PROTOTYPES = {
    "lsk_version": (ctypes.c_int, []),
    "lsk_error_string": (ctypes.c_char_p, [ctypes.c_int]),
    "lsk_snapshot_doubles": (ctypes.c_int, [ctypes.POINTER(ctypes.c_void_p), ctypes.c_int, ctypes.c_void_p, ctypes.c_ulonglong, ctypes.c_void_p]),
    "lsk_embed_points2_snapshot": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_longlong, ctypes.POINTER(ctypes.c_int), ctypes.c_void_p,
                                                  ctypes.c_void_p, ctypes.c_longlong, ctypes.POINTER(ctypes.c_int), ctypes.c_void_p,
                                                  ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.POINTER(ctypes.c_int),
                                                  ctypes.POINTER(ctypes.c_int), ctypes.c_int,
                                                  ctypes.POINTER(ctypes.c_void_p), ctypes.c_int, ctypes.c_void_p, ctypes.c_ulonglong,
                                                  ctypes.c_void_p]),
    "lsk_embed_buffer_doubles": (ctypes.c_longlong, [ctypes.c_longlong, ctypes.c_int]),
    "lsk_embed_points": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_longlong, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                        ctypes.POINTER(ctypes.c_int), ctypes.POINTER(ctypes.c_int),
                                        ctypes.POINTER(ctypes.c_int), ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]),
    "lsk_embed_points2": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_longlong, ctypes.POINTER(ctypes.c_int), ctypes.c_void_p,
                                         ctypes.c_void_p, ctypes.c_longlong, ctypes.POINTER(ctypes.c_int), ctypes.c_void_p,
                                         ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.POINTER(ctypes.c_int),
                                         ctypes.POINTER(ctypes.c_int), ctypes.c_int, ctypes.c_void_p]),
    "lsk_lift_frames": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_longlong, ctypes.c_int, ctypes.c_int, ctypes.c_void_p,
                                       ctypes.c_void_p, ctypes.c_void_p]),
    "lsk_homog_pairs": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_longlong, ctypes.c_longlong,
                                       ctypes.c_int, ctypes.c_int, ctypes.POINTER(LskEpilogue), ctypes.c_void_p,
                                       ctypes.c_longlong, ctypes.c_void_p]),
    "lsk_homog_sample": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_longlong, ctypes.c_longlong,
                                        ctypes.c_int, ctypes.c_int, ctypes.POINTER(LskEpilogue), ctypes.c_void_p, ctypes.c_void_p,
                                        ctypes.c_void_p]),
    "lsk_hyp_features": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_longlong,
                                        ctypes.c_int, ctypes.c_int, ctypes.c_double, ctypes.c_void_p, ctypes.c_int,
                                        ctypes.c_longlong, ctypes.c_int, ctypes.c_void_p]),
    "lsk_spd_features": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_longlong,
                                        ctypes.c_int, ctypes.c_int, ctypes.c_double, ctypes.c_void_p, ctypes.c_int,
                                        ctypes.c_longlong, ctypes.c_int, ctypes.c_void_p]),
    "lsk_hyp_sample": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                      ctypes.c_longlong, ctypes.c_int, ctypes.c_int, ctypes.c_double, ctypes.c_void_p, ctypes.c_void_p]),
    "lsk_spd_sample": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                      ctypes.c_longlong, ctypes.c_int, ctypes.c_int, ctypes.c_double, ctypes.c_void_p, ctypes.c_void_p]),
    "lsk_spd_logdiag": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_longlong, ctypes.c_int, ctypes.c_int,
                                       ctypes.c_void_p, ctypes.c_void_p]),
    "lsk_small_matrix": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_longlong, ctypes.c_int, ctypes.c_int, ctypes.c_void_p,
                                        ctypes.c_void_p, ctypes.c_void_p]),
    "lsk_haar_from_gaussian": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_longlong, ctypes.c_int, ctypes.c_int, ctypes.c_void_p,
                                              ctypes.c_void_p]),
    "lsk_quaternion_to_group": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_longlong, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]),
    "lsk_torus_features": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_longlong, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p,
                                          ctypes.c_int, ctypes.c_double, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p]),
    "lsk_panel_gemv": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_longlong, ctypes.c_int, ctypes.c_void_p, ctypes.c_double,
                                      ctypes.c_void_p, ctypes.c_void_p]),
    "lsk_potrf_work_doubles": (ctypes.c_longlong, [ctypes.c_longlong]),
    "lsk_potrf_upper": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_longlong, ctypes.c_longlong, ctypes.c_void_p, ctypes.c_void_p,
                                       ctypes.c_void_p]),
    "lsk_potrs_upper": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_longlong, ctypes.c_longlong, ctypes.c_void_p, ctypes.c_void_p,
                                       ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p]),
    "lsk_trsm_work_doubles": (ctypes.c_longlong, [ctypes.c_longlong, ctypes.c_longlong]),
    "lsk_trsm_lower": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_longlong, ctypes.c_longlong, ctypes.c_void_p, ctypes.c_void_p,
                                      ctypes.c_longlong, ctypes.c_longlong, ctypes.c_void_p, ctypes.c_void_p]),
    "lsk_potri_work_doubles": (ctypes.c_longlong, [ctypes.c_longlong]),
    "lsk_potri_upper": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_longlong, ctypes.c_longlong, ctypes.c_void_p, ctypes.c_void_p,
                                       ctypes.c_void_p, ctypes.c_longlong, ctypes.c_void_p]),
    "lsk_pairwise_poly": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_longlong, ctypes.c_longlong, ctypes.c_int,
                                         ctypes.POINTER(LskEpilogue), ctypes.c_void_p, ctypes.c_longlong, ctypes.c_void_p]),
}


def exported_symbols():
    return sorted(_PROTOTYPES)


def load():
    """Loads liblsk.so (once). Raises LskError when it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise LskError("liblsk.so not found at %s - run `python -m liestationarykernels_b200.build` "
                           "(there is no CPU fallback)" % LIB_PATH)
        lib = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in _PROTOTYPES.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def require_cuda():
    if not torch.cuda.is_available():
        raise LskError("liestationarykernels_b200 needs a CUDA device (sm_100a); there is no CPU fallback")


def check(code: int, what: str):
    if code == 0:
        return
    text = load().lsk_error_string(int(code))
    raise LskError("%s: %s (code %d)" % (what, text.decode() if text else "error", code))


def stream_ptr():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def ptr(t: torch.Tensor):
    return ctypes.c_void_p(t.data_ptr())


_launch_count = 0


def count_launch(n=1):
    """bench.py reports `gpu_launches`: every kernel launch made through this binding is counted here."""
    global _launch_count
    _launch_count += n


def launches():
    return _launch_count
