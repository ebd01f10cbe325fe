"""Decompose GEMM time into fixed overhead + per-k-block slope: one full wave of tiles (148 CTAs), K swept."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from multimodal_classification_b200 import ops
from tools.bench_kernels import graph_time, rnd

def main():
    for bn, n in ((128, 128), (256, 256), (64, 64)):
        for mt in (148, 296, 74):
            m = mt * 128
            prev = None
            for k in (64, 256, 768, 1536, 3072, 6144):
                a, b = rnd(m, k), rnd(n, k)
                out = torch.empty(m, n, device="cuda", dtype=torch.bfloat16)
                t = graph_time(lambda: ops.gemm(a, b, out, block_n=bn))
                tc = graph_time(lambda: torch.matmul(a, b.t(), out=out))
                fl = 2.0 * m * n * k
                print(f"bn{bn} tiles{mt} K{k:5d}: {t:7.2f} us ({fl/t/1e6:6.0f} TF)  cublas {tc:7.2f} us ({fl/tc/1e6:6.0f} TF)", flush=True)
    # square-ish big problem: steady state
    for m, n, k in ((8192, 8192, 8192), (4096, 4096, 4096)):
        a, b = rnd(m, k), rnd(n, k)
        out = torch.empty(m, n, device="cuda", dtype=torch.bfloat16)
        for bn in (128, 256):
            t = graph_time(lambda: ops.gemm(a, b, out, block_n=bn), rep=3, iters=3)
            print(f"big {m} bn{bn}: {t:8.1f} us ({2.0*m*n*k/t/1e6:6.0f} TF)", flush=True)
        t = graph_time(lambda: torch.matmul(a, b.t(), out=out), rep=3, iters=3)
        print(f"big {m} cublas: {t:8.1f} us ({2.0*m*n*k/t/1e6:6.0f} TF)", flush=True)

if __name__ == "__main__":
    main()
