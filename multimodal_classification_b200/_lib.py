"""ctypes binding of libvilbert_b200.so (the C ABI declared in include/vilbert_b200.h).

The shared library is built in-tree by ``__graft_entry__.build()`` (nvcc, sm_100a).  There is no
fallback: if it is missing, or a call returns a non-zero status, we raise.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libvilbert_b200.so")

ACT_NONE, ACT_GELU, ACT_RELU, ACT_TANH = 0, 1, 2, 3
AUX_NONE, AUX_ADD, AUX_MUL_GELU_GRAD = 0, 1, 2


class VbError(RuntimeError):
    pass


class GemmArgs(C.Structure):
    _fields_ = [
        ("a", C.c_void_p), ("b", C.c_void_p), ("d", C.c_void_p), ("d_preact", C.c_void_p),
        ("scale", C.c_void_p), ("bias", C.c_void_p), ("aux", C.c_void_p),
        ("lda", C.c_int64), ("ldb", C.c_int64), ("ldd", C.c_int64), ("ld_preact", C.c_int64), ("ld_aux", C.c_int64),
        ("m", C.c_int32), ("n", C.c_int32), ("k", C.c_int32),
        ("a_mn_major", C.c_int32), ("b_mn_major", C.c_int32),
        ("d_is_f32", C.c_int32), ("accumulate", C.c_int32), ("act", C.c_int32), ("aux_mode", C.c_int32),
        ("block_n", C.c_int32), ("splits", C.c_int32), ("max_ctas", C.c_int32),
    ]


_lib = None


def lib() -> C.CDLL:
    """Load the kernel library, failing loudly when it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise VbError(
                f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                "(nvcc -gencode arch=compute_100a,code=sm_100a). There is no CPU or PyTorch fallback.")
        _lib = C.CDLL(LIB_PATH)
        _declare(_lib)
    return _lib


def _declare(l: C.CDLL) -> None:
    l.vb_abi_version.restype = C.c_int
    l.vb_last_error.restype = C.c_char_p
    l.vb_build_info.restype = C.c_char_p
    l.vb_gemm_bf16.argtypes = [C.POINTER(GemmArgs), C.c_void_p]
    l.vb_gemm_bf16.restype = C.c_int


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = lib().vb_last_error().decode("utf-8", "replace")
        raise VbError(f"{what} failed with status {rc}: {msg}")


def exported_symbols_in_header() -> list[str]:
    """Names of every function declared in include/vilbert_b200.h (for the ABI test)."""
    import re
    hdr = os.path.join(os.path.dirname(_HERE), "include", "vilbert_b200.h")
    text = open(hdr).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(vb_[a-z0-9_]+)\s*\(", text)))
