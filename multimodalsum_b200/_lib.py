"""ctypes binding of libmmsum_b200.so (the C-ABI declared in include/mmsum_b200.h).

The library is the product: if it is missing this module raises — there is no torch/CPU fallback.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libmmsum_b200.so")

_lib = None


class MmsumError(RuntimeError):
    pass


class GemmArgs(C.Structure):
    _fields_ = [
        ("A", C.c_void_p), ("B", C.c_void_p), ("D", C.c_void_p),
        ("lda", C.c_int64), ("ldb", C.c_int64), ("ldd", C.c_int64),
        ("M", C.c_int32), ("N", C.c_int32), ("K", C.c_int32),
        ("a_mn_major", C.c_int32), ("b_mn_major", C.c_int32),
        ("out_f32", C.c_int32), ("accumulate", C.c_int32), ("splits", C.c_int32),
        ("block_n", C.c_int32), ("raster_m_fast", C.c_int32),
        ("alpha", C.c_float), ("bias", C.c_void_p),
        ("act", C.c_int32), ("aux_mode", C.c_int32), ("aux", C.c_void_p), ("ld_aux", C.c_int64),
    ]


def lib():
    """Load (once) and return the shared library; raise loudly when it is not built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise MmsumError(
                "libmmsum_b200.so is not built (%s). Run `python -m multimodalsum_b200.build` "
                "(needs nvcc); there is no fallback path." % LIB_PATH)
        _lib = C.CDLL(LIB_PATH)
        for name in EXPORTS:
            fn = getattr(_lib, name)  # AttributeError if the symbol is missing
            fn.restype = C.c_int
    return _lib


# every symbol include/mmsum_b200.h declares (tests/test_abi.py cross-checks this list against the header)
EXPORTS = [
    "mmsum_gemm_bf16",
]


def check(rc, what):
    if rc != 0:
        kind = "invalid argument / driver" if rc < 0 else "cudaError_t"
        raise MmsumError("%s failed: rc=%d (%s)" % (what, rc, kind))
