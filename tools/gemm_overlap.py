"""Do consecutive GEMM launches on one stream overlap (PDL + two CTAs per SM)?  %globaltimer at entry / exit of every CTA of
three back-to-back launches captured in one CUDA graph (library built with -DVB_GEMM_TRACE)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from multimodal_classification_b200 import ops, _lib
from tools.bench_kernels import rnd

def run(m, n, k, **kw):
    a, b = rnd(m, k), rnd(n, k)
    out = torch.empty(m, n, device="cuda", dtype=torch.bfloat16)
    bufs = [torch.zeros(320 * 24, dtype=torch.int64, device="cuda") for _ in range(3)]
    for _ in range(3):
        ops.gemm(a, b, out, **kw)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for buf in bufs:
            _lib.lib().vb_gemm_set_trace(buf.data_ptr())
            ops.gemm(a, b, out, **kw)
    _lib.lib().vb_gemm_set_trace(None)
    g.replay(); torch.cuda.synchronize()
    g.replay(); torch.cuda.synchronize()
    t0 = None
    for i, buf in enumerate(bufs):
        t = buf.view(320, 24).cpu()
        t = t[t[:, 22] != 0]
        ent, ext = t[:, 22], t[:, 23]
        if t0 is None:
            t0 = ent.min().item()
        print(f"launch {i}: {t.shape[0]} CTAs  entry {ent.min().item()-t0:6d} .. {ent.max().item()-t0:6d} ns   exit {ext.min().item()-t0:6d} .. {ext.max().item()-t0:6d} ns"
              f"   depwait cycles median {(t[:,2]-t[:,1]).median().item()}")

if __name__ == "__main__":
    print("2048x768x768 (auto)"); run(2048, 768, 768)
    print("2048x768x768 bn=256 (one CTA per SM)"); run(2048, 768, 768, block_n=256)
    print("2048x3072x768 (auto)"); run(2048, 3072, 768)
