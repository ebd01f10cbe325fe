import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
torch.zeros(1, device="cuda")
from multimodal_classification_b200 import _lib
l = _lib.lib()
for kb in (48, 64, 96, 100, 104, 108, 110, 112, 113):
    b, c = C.c_int(0), C.c_int(0)
    rc = l.vb_gemm_debug_occupancy(kb * 1024, C.byref(b), C.byref(c))
    print(kb, "KB ->", rc, "blocks/SM", b.value, "2-CTA clusters", c.value, flush=True)
