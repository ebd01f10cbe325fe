"""In-kernel timeline of one GEMM launch (clock64 stamps written by every CTA, see vb_gemm_set_trace)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from multimodal_classification_b200 import ops, _lib
from tools.bench_kernels import rnd, graph_time

NAMES = ["entry", "prologue", "depwait", "load0", "loads_done", "operands0", "mma_issued", "acc0_ready", "accN_ready",
         "stores_issued", "staging_drained", "exit", "ldtm0_done", "units_done", "proxy_fenced", "epi_synced", "store0_issued",
         "prod_begin", "prod_empty_ok", "bars_inited", "tmem_alloced", "cta_synced"]

def trace(name, m, n, k, wgrad=False, dgrad=False, gelu=False, **kw):
    if wgrad:      # dW[m,n] = dY[k,m]^T X[k,n]: both operands MN-major, fp32 output
        a, b = rnd(k, m), rnd(k, n)
        out = torch.zeros(m, n, device="cuda", dtype=torch.float32)
        kw = dict(kw, a_mn_major=True, b_mn_major=True)
    elif dgrad:    # dX[m,n] = dY[m,k] W[k,n] + aux: B MN-major, residual-gradient add in the epilogue
        a, b = rnd(m, k), rnd(k, n)
        out = torch.empty(m, n, device="cuda", dtype=torch.bfloat16)
        kw = dict(kw, b_mn_major=True, aux=rnd(m, n), aux_mode=ops.AUX_MUL_GELU_GRAD if gelu else ops.AUX_ADD)
    else:
        a, b = rnd(m, k), rnd(n, k)
        out = torch.empty(m, n, device="cuda", dtype=torch.bfloat16)
        kw = dict(kw, bias=torch.randn(n, device="cuda"))
        if gelu:
            kw = dict(kw, act=ops.ACT_GELU, preact=torch.empty(m, n, device="cuda", dtype=torch.bfloat16))
    buf = torch.zeros(160 * 24, dtype=torch.int64, device="cuda")
    for _ in range(3):
        ops.gemm(a, b, out, **kw)
    torch.cuda.synchronize()
    _lib.lib().vb_gemm_set_trace(buf.data_ptr())
    ops.gemm(a, b, out, **kw)
    torch.cuda.synchronize()
    _lib.lib().vb_gemm_set_trace(None)
    t = buf.view(160, 24).cpu()
    used = t[:, 0] != 0
    t = t[used]
    rel = (t - t[:, :1]).float()
    t_us = graph_time(lambda: ops.gemm(a, b, out, **kw))
    kw = {k_: (v if not torch.is_tensor(v) else "T") for k_, v in kw.items()}
    print(f"{name} {m}x{n}x{k} {kw}: {t_us:.2f} us/launch in a graph, {int(used.sum())} CTAs; median cycles since entry (min..max):")
    rows = []
    for i, nm in enumerate(NAMES):
        col = rel[:, i][t[:, i] != 0]
        if col.numel():
            rows.append((col.median().item(), nm, col.min().item(), col.max().item()))
    for med, nm, lo, hi in sorted(rows):
        print(f"   {nm:16s} {med:9.0f}  ({lo:.0f} .. {hi:.0f})")

if __name__ == "__main__":
    which = sys.argv[1] if len(sys.argv) > 1 else "fwd"
    if which == "suite":
        trace("t.attn_out fwd", 2048, 768, 768)
        trace("t.qkv fwd", 2048, 2304, 768)
        trace("t.ffn1 fwd gelu+preact", 2048, 3072, 768, gelu=True)
        trace("c.tqkv fwd plain", 2048, 3072, 768)
        trace("t.ffn2 fwd", 2048, 768, 3072)
        trace("v.1024 fwd", 1600, 1024, 1024)
        trace("t.attn_out dgrad+aux", 2048, 768, 768, dgrad=True)
        trace("t.ffn2 dgrad gelu'", 2048, 3072, 768, dgrad=True, gelu=True)
        trace("t.ffn1 dgrad+aux", 2048, 768, 3072, dgrad=True)
        trace("t.ffn1 wgrad", 3072, 768, 2048, wgrad=True)
        trace("t.ffn1 wgrad 48 ctas", 3072, 768, 2048, wgrad=True, max_ctas=48)
        trace("t.attn_out wgrad", 768, 768, 2048, wgrad=True)
    elif which == "wgrad":
        trace("t.ffn1 wgrad", 3072, 768, 2048, wgrad=True)
        trace("t.attn_out wgrad", 768, 768, 2048, wgrad=True)
        trace("t.ffn1 wgrad d_streamed", 3072, 768, 2048, wgrad=True, d_streamed=True)
    else:
        trace("t.attn_out", 2048, 768, 768)
        trace("t.ffn1", 2048, 3072, 768)
