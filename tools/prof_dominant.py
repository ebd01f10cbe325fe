"""The kernel bench.py names as dominant, alone, for `ncu --set full`: text FFN-1 forward GEMM 2048x3072x768 + bias + GELU with
the pre-activation kept for backward (18 launches per forward; the same tile configuration serves its dgrad / wgrad)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from multimodal_classification_b200 import ops
m, n, k = 2048, 3072, 768
a = torch.randn(m, k, device="cuda").to(torch.bfloat16)
w = torch.randn(n, k, device="cuda").to(torch.bfloat16)
bias = torch.randn(n, device="cuda")
out = torch.empty(m, n, device="cuda", dtype=torch.bfloat16)
pre = torch.empty_like(out)
for _ in range(6):
    ops.gemm(a, w, out, bias=bias, act=ops.ACT_GELU, preact=pre, b_streamed=True)
torch.cuda.synchronize()
print("ok")
