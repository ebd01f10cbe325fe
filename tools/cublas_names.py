"""Which kernels does cuBLAS pick at the ViLBERT shapes?  Run under `ncu --metrics gpu__time_duration.sum`: the kernel
names carry the tile / cluster / 1SM-2SM configuration, the yardstick our GEMM is measured against."""
import torch

BF = torch.bfloat16
shapes = [("t.qkv", 2048, 2304, 768), ("t.attn_out", 2048, 768, 768), ("t.ffn1", 2048, 3072, 768), ("t.ffn2", 2048, 768, 3072),
          ("v.qkv", 1600, 3072, 1024), ("v.1024", 1600, 1024, 1024), ("c.dense2", 2048, 768, 1024), ("img_emb", 1600, 1024, 2048)]
for name, m, n, k in shapes:
    x = torch.randn(m, k, device="cuda").to(BF)
    w = torch.randn(n, k, device="cuda").to(BF)
    dy = torch.randn(m, n, device="cuda").to(BF)
    y = torch.empty(m, n, device="cuda", dtype=BF)
    dx = torch.empty(m, k, device="cuda", dtype=BF)
    dw = torch.empty(n, k, device="cuda", dtype=BF)
    for _ in range(2):
        torch.matmul(x, w.t(), out=y)
        torch.matmul(dy, w, out=dx)
        torch.matmul(dy.t(), x, out=dw)
    torch.cuda.synchronize()
    print(name, m, n, k, flush=True)
