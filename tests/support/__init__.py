"""ctypes loader of the test-support library (comparators only; see test_support.cu)."""
import ctypes as C
import os

from evo_b200._lib import AttnParams, GemmParams

LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "libevo_b200_test.so")
_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(f"{LIB_PATH} is missing: run `python -m evo_b200.build`")
        h = C.CDLL(LIB_PATH)
        h.evot_last_error.restype = C.c_char_p
        h.evot_gemm_cublaslt.restype = C.c_int
        h.evot_gemm_cublaslt.argtypes = [C.POINTER(GemmParams), C.c_void_p, C.c_size_t, C.c_void_p]
        h.evot_attn_fwd_simple.restype = C.c_int
        h.evot_attn_fwd_simple.argtypes = [C.POINTER(AttnParams), C.c_void_p]
        h.evot_add.restype = C.c_int
        h.evot_add.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]
        _lib = h
    return _lib


def check(rc, what=""):
    if rc != 0:
        raise RuntimeError(f"{what} failed (rc={rc}): {lib().evot_last_error().decode()}")
