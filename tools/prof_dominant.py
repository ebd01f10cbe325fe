"""The GEMM shapes that dominate the step, alone, for `ncu --set full` (+ the tensor-pipe counters): each shape is launched four
times (three warm-ups, the fourth is the one to read).  Order: text FFN-1 forward (bias + GELU + pre-activation kept), text FFN-2
dgrad (gelu' of the saved pre-activation), text FFN-1 wgrad (fp32 out), text attention-output forward (the narrow shape)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from multimodal_classification_b200 import ops
bf = torch.bfloat16
def rnd(*s):
    return (torch.randn(*s, device="cuda") * 0.5).to(bf)
m = 2048
x768, w3072, bias = rnd(m, 768), rnd(3072, 768), torch.randn(3072, device="cuda")
y, pre = torch.empty(m, 3072, device="cuda", dtype=bf), torch.empty(m, 3072, device="cuda", dtype=bf)
dy768, w_ffn2, g_pre = rnd(m, 768), rnd(768, 3072), torch.empty(m, 3072, device="cuda", dtype=bf)
dw = torch.zeros(3072, 768, device="cuda")
w768, b768, y768 = rnd(768, 768), torch.randn(768, device="cuda"), torch.empty(m, 768, device="cuda", dtype=bf)
for _ in range(4):
    ops.gemm(x768, w3072, y, bias=bias, act=ops.ACT_GELU, preact=pre, b_streamed=True, preact_grad=True)
for _ in range(4):
    ops.gemm(dy768, w_ffn2, g_pre, b_mn_major=True, aux=pre, aux_mode=ops.AUX_MUL, b_streamed=True)
for _ in range(4):
    ops.gemm(y, x768, dw, a_mn_major=True, b_mn_major=True, d_streamed=True)
for _ in range(4):
    ops.gemm(x768, w768, y768, bias=b768, b_streamed=True)
torch.cuda.synchronize()
print("ok")
