"""GEMM main-loop dissection (library built with -DVB_GEMM_TRACE; VB_GEMM_DEBUG=1 no MMA / 2 no TMA; VB_GEMM_STAGES=n)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from multimodal_classification_b200 import ops
from tools.bench_kernels import graph_time, rnd

cases = [("syn148", 148 * 128, 256, None), ("t.ffn1", 2048, 3072, 768), ("t.attn_out", 2048, 768, 768), ("v.qkv", 1600, 3072, 1024)]
for name, m, n, kk in cases:
    for bn in (64, 128, 256):
        if n < bn: continue
        row = []
        for k in ((64, 768, 3072) if kk is None else (kk,)):
            a, b = rnd(m, k), rnd(n, k)
            out = torch.empty(m, n, device="cuda", dtype=torch.bfloat16)
            row.append(f"K{k} {graph_time(lambda: ops.gemm(a, b, out, block_n=bn)):7.2f}")
        print(f"dbg{os.environ.get('VB_GEMM_DEBUG','0')} st{os.environ.get('VB_GEMM_STAGES','-')} {name:10s} bn{bn}: " + "  ".join(row), flush=True)
