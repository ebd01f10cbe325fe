"""Micro-benchmark of the tcgen05 GEMM on the Appendix-B shapes against torch.matmul (cuBLAS).
Prints one line per shape; CUDA-event timing, L2 flushed between iterations is NOT done here because at
bs16 the operands are L2 resident in the real step as well (documented in DESIGN.md)."""
import sys
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from multimodal_classification_b200 import ops


def timeit(fn, iters=50):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(iters):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / iters * 1e-3


def main():
    shapes = [
        ("text qkv", 2048, 2304, 768), ("text attn-out", 2048, 768, 768), ("text ffn1", 2048, 3072, 768),
        ("text ffn2", 2048, 768, 3072), ("vis qkv", 1600, 3072, 1024), ("vis 1024", 1600, 1024, 1024),
        ("co text qkv", 2048, 3072, 768), ("co dense2", 2048, 768, 1024), ("img emb", 1600, 1024, 2048),
        ("big", 8192, 8192, 8192), ("bs512 ffn1", 65536, 3072, 768),
    ]
    for name, m, n, k in shapes:
        a = torch.randn(m, k, device="cuda").to(torch.bfloat16)
        b = torch.randn(n, k, device="cuda").to(torch.bfloat16)
        out = torch.empty(m, n, device="cuda", dtype=torch.bfloat16)
        flops = 2.0 * m * n * k
        t_ref = timeit(lambda: torch.matmul(a, b.t(), out=out))
        res = []
        for bn in (64, 128, 256):
            t = timeit(lambda: ops.gemm(a, b, out, block_n=bn))
            res.append(f"bn{bn} {flops / t / 1e12:7.1f}")
        t_auto = timeit(lambda: ops.gemm(a, b, out))
        print(f"{name:14s} M{m} N{n} K{k}: cublas {flops / t_ref / 1e12:7.1f} TF | " + " | ".join(res) +
              f" | auto {flops / t_auto / 1e12:7.1f} TF ({t_auto * 1e6:.1f} us)", flush=True)
        # wgrad-shaped: dW[n,k] = dY[m,n]^T X[m,k]
        dy = torch.randn(m, n, device="cuda").to(torch.bfloat16)
        dw = torch.zeros(n, k, device="cuda", dtype=torch.float32)
        if m <= 8192:
            t_w = timeit(lambda: ops.gemm(dy, a, dw, a_mn_major=True, b_mn_major=True, accumulate=True))
            t_w1 = timeit(lambda: ops.gemm(dy, a, dw, a_mn_major=True, b_mn_major=True, accumulate=False))
            t_wref = timeit(lambda: torch.matmul(dy.t(), a))
            print(f"{'':14s} wgrad: cublas {flops / t_wref / 1e12:7.1f} TF | splitk-auto {flops / t_w / 1e12:7.1f} | nosplit {flops / t_w1 / 1e12:7.1f}", flush=True)


if __name__ == "__main__":
    main()
