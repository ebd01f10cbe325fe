"""CPU-side checks of the drop-in boundary: the C-ABI library loads without a GPU and exports every
symbol include/vilbert_b200.h declares."""
import ctypes
import os

from multimodal_classification_b200 import _lib


def test_library_loads_and_exports_header_symbols():
    assert os.path.exists(_lib.LIB_PATH), "build first: python -c 'import __graft_entry__ as g; g.build()'"
    l = ctypes.CDLL(_lib.LIB_PATH)
    names = _lib.exported_symbols_in_header()
    assert "vb_gemm_bf16" in names and len(names) >= 4
    missing = [n for n in names if not hasattr(l, n)]
    assert not missing, f"declared in the header but not exported: {missing}"
    assert _lib.lib().vb_abi_version() >= 1
    assert b"sm_100a" in _lib.lib().vb_build_info()


def test_bad_args_are_reported_not_thrown():
    l = _lib.lib()
    args = _lib.GemmArgs()
    rc = l.vb_gemm_bf16(ctypes.byref(args), None)
    assert rc == -1
    assert b"non-null" in l.vb_last_error()
