"""Main-loop rate of the GEMM kernel as a function of ring depth, tile width, cluster shape and operand layout: cycles per
64-deep k-block between "first operands landed" and "accumulator ready" of a long-K problem (trace build of the library,
VB_LIB=...libvilbert_b200_trace.so).  debug_mode 1 = no MMA issue (pure load rate), 2 = no TMA (pure MMA issue rate)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from multimodal_classification_b200 import ops, _lib
from tools.bench_kernels import rnd

L = _lib.lib()
def knob(**kw):
    for k, v in kw.items():
        _lib.check(L.vb_gemm_set_knob(k.encode(), v), "vb_gemm_set_knob"); _lib._launches -= 1

def probe(kind, m, n, k, bn, **kw):
    if kind == "wgrad":
        a, b, out = rnd(k, m), rnd(k, n), torch.zeros(m, n, device="cuda")
        args = dict(a_mn_major=True, b_mn_major=True)
    elif kind == "dgrad":
        a, b, out = rnd(m, k), rnd(k, n), torch.empty(m, n, device="cuda", dtype=torch.bfloat16)
        args = dict(b_mn_major=True)
    else:
        a, b, out = rnd(m, k), rnd(n, k), torch.empty(m, n, device="cuda", dtype=torch.bfloat16)
        args = {}
    buf = torch.zeros(160 * 24, dtype=torch.int64, device="cuda")
    for _ in range(2):
        ops.gemm(a, b, out, block_n=bn, **args)
    torch.cuda.synchronize()
    L.vb_gemm_set_trace(buf.data_ptr())
    ops.gemm(a, b, out, block_n=bn, **args)
    torch.cuda.synchronize()
    L.vb_gemm_set_trace(None)
    t = buf.view(160, 24).cpu()
    t = t[t[:, 0] != 0]
    rel = (t - t[:, :1]).float()
    ctas = t.shape[0]
    kb = (k + 63) // 64
    # 5 = first operands landed (leader CTAs only), 8 = last accumulator ready, 3 = first load issued
    lead = t[:, 5] != 0
    main = (rel[lead, 8] - rel[lead, 5]).median().item()
    tiles = ((m + 255) // 256) * ((n + bn - 1) // bn)
    per_cta_tiles = max(1, -(-tiles // (ctas // 2)))
    return ctas, main / (kb * per_cta_tiles), rel[:, 3].median().item(), rel[lead, 5].median().item()

if __name__ == "__main__":
    K = 3072
    print("kind   MxNxK            bn np stages dbg | ctas  cyc/kblock  B/clk/SM   (load0, operands0)")
    for kind, m, n in (("fwd", 2048, 768), ("fwd", 2048, 3072), ("dgrad", 2048, 768), ("wgrad", 3072, 768)):
        for bn in ((96, 128, 256) if kind == "fwd" else (128, 256)):
            for np_ in (1, 2):
                for stages in (3, 8):
                    for dbg in (0, 1):
                        if dbg == 1 and stages != 8: continue
                        knob(np=np_, stages=stages, debug_mode=dbg, occ1=1)
                        try:
                            ctas, cyc, l0, o0 = probe(kind, m, n, K, bn)
                        except Exception as e:
                            print(kind, m, n, bn, np_, stages, dbg, "ERR", str(e)[:80]); continue
                        bytes_kb = 16384 + bn * 64
                        print(f"{kind:6s} {m}x{n}x{K:<5d} {bn:4d} {np_:2d} {stages:5d} {dbg:4d} | {ctas:4d} {cyc:10.0f} {bytes_kb / cyc:9.1f}   ({l0:.0f}, {o0:.0f})", flush=True)
