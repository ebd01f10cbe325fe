import os, sys
sys.path.insert(0, "/root/repo")
import torch
from multimodal_classification_b200 import ops
from tools.bench_kernels import rnd
for (m, n, k) in ((2048,2304,768),(2048,768,768),(2048,3072,768),(2048,768,3072),(1600,3072,1024),(1600,1024,1024),(2048,768,1024),(1600,1024,2048)):
    x, w, dy = rnd(m,k), rnd(n,k), rnd(m,n)
    y, dx, dw = torch.empty(m,n,device="cuda",dtype=torch.bfloat16), torch.empty(m,k,device="cuda",dtype=torch.bfloat16), torch.zeros(n,k,device="cuda")
    ops.gemm(x, w, y); ops.gemm(dy, w, dx, b_mn_major=True); ops.gemm(dy, x, dw, a_mn_major=True, b_mn_major=True)
torch.cuda.synchronize()
