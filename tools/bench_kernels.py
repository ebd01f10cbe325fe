"""Warm per-kernel attribution of the ViLBERT training step: every distinct (kernel, shape) of the bs16 step is timed
as a CUDA-graph of REP back-to-back launches (so ctypes/launch overhead is excluded, as in the real graph-replayed
step), multiplied by its count per step.  GEMMs are also timed through torch.matmul (cuBLAS) as a yardstick.

    python tools/bench_kernels.py [--only gemm|ln|attn|misc]
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from multimodal_classification_b200 import ops

REP = 20
BF = torch.bfloat16


def graph_time(fn, rep=REP, iters=10):
    """us per launch of fn, measured over graph replays of `rep` launches."""
    fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(rep):
            fn()
    g.replay()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(iters):
        g.replay()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / (iters * rep) * 1e3


def rnd(*shape, dtype=BF):
    return (torch.randn(*shape, device="cuda") * 0.5).to(dtype)


def bench_gemms():
    Mt, Mv = 2048, 1600
    # (name, M, N, K, count_fwd, act, has_preact)
    lin = [
        ("t.qkv", Mt, 2304, 768, 12, 0), ("t.attn_out", Mt, 768, 768, 12, 0), ("t.ffn1", Mt, 3072, 768, 18, 1),
        ("t.ffn2", Mt, 768, 3072, 18, 0), ("v.qkv", Mv, 3072, 1024, 12, 0), ("v.1024", Mv, 1024, 1024, 18, 0),
        ("v.ffn1", Mv, 1024, 1024, 12, 1), ("c.tqkv", Mt, 3072, 768, 6, 0), ("c.dense2", Mt, 768, 1024, 6, 0),
        ("img_emb", Mv, 1024, 2048, 1, 0),
    ]
    total = {"fwd": 0.0, "dgrad": 0.0, "wgrad": 0.0, "cublas_fwd": 0.0, "cublas_dgrad": 0.0, "cublas_wgrad": 0.0}
    print(f"{'site':10s} {'M':>5s} {'N':>5s} {'K':>5s}  cnt |   fwd us (TF)  cublas |  dgrad us (TF) cublas |  wgrad us (TF) cublas")
    for name, m, n, k, cnt, gelu in lin:
        x, w = rnd(m, k), rnd(n, k)
        bias = torch.randn(n, device="cuda")
        y, pre = torch.empty(m, n, device="cuda", dtype=BF), torch.empty(m, n, device="cuda", dtype=BF)
        dy, dx, aux = rnd(m, n), torch.empty(m, k, device="cuda", dtype=BF), rnd(m, k)
        dw = torch.zeros(n, k, device="cuda")
        fl = 2.0 * m * n * k
        if gelu:
            t_f = graph_time(lambda: ops.gemm(x, w, y, bias=bias, act=ops.ACT_GELU, preact=pre, preact_grad=True))      # as the engine
        else:
            t_f = graph_time(lambda: ops.gemm(x, w, y, bias=bias))
        t_fc = graph_time(lambda: torch.matmul(x, w.t(), out=y))
        # dgrad: dx = dy W (+aux)   [for ffn2 the real step fuses gelu' instead]
        if name in ("t.ffn2",):
            auxk = rnd(m, k)
            t_d = graph_time(lambda: ops.gemm(dy, w, dx, b_mn_major=True, aux=auxk, aux_mode=ops.AUX_MUL))
        else:
            t_d = graph_time(lambda: ops.gemm(dy, w, dx, b_mn_major=True, aux=aux, aux_mode=ops.AUX_ADD))
        t_dc = graph_time(lambda: torch.matmul(dy, w, out=dx))
        t_w = graph_time(lambda: ops.gemm(dy, x, dw, a_mn_major=True, b_mn_major=True, accumulate=os.environ.get("VB_WGRAD_ACC", "0") == "1"))
        dwb = torch.empty(n, k, device="cuda", dtype=BF)
        t_wc = graph_time(lambda: torch.matmul(dy.t(), x, out=dwb))
        tf = lambda t: fl / t / 1e6
        print(f"{name:10s} {m:5d} {n:5d} {k:5d} {cnt:4d} | {t_f:6.1f} ({tf(t_f):5.0f}) {t_fc:6.1f} | {t_d:6.1f} ({tf(t_d):5.0f}) {t_dc:6.1f} |"
              f" {t_w:6.1f} ({tf(t_w):5.0f}) {t_wc:6.1f}", flush=True)
        total["fwd"] += cnt * t_f; total["dgrad"] += cnt * t_d; total["wgrad"] += cnt * t_w
        total["cublas_fwd"] += cnt * t_fc; total["cublas_dgrad"] += cnt * t_dc; total["cublas_wgrad"] += cnt * t_wc
    print("GEMM us/step (serial sum):", {k: round(v, 1) for k, v in total.items()},
          "ours", round(total["fwd"] + total["dgrad"] + total["wgrad"], 1),
          "cublas", round(total["cublas_fwd"] + total["cublas_dgrad"] + total["cublas_wgrad"], 1), flush=True)


def bench_ln():
    seed = torch.tensor([1234], dtype=torch.int64, device="cuda")
    tot = 0.0
    for name, m, h, cnt in (("t.ln", 2048, 768, 36), ("v.ln", 1600, 1024, 25)):
        x, res, dy = rnd(m, h), rnd(m, h), rnd(m, h)
        g, b = torch.ones(h, device="cuda"), torch.zeros(h, device="cuda")
        y, dx, dres = torch.empty_like(x), torch.empty_like(x), torch.empty_like(x)
        mean, rstd = torch.empty(m, device="cuda"), torch.empty(m, device="cuda")
        dg, db, dbias = torch.zeros(h, device="cuda"), torch.zeros(h, device="cuda"), torch.zeros(h, device="cuda")
        for drop in (0.0, 0.1):
            sd = seed if drop else None
            t_f = graph_time(lambda: ops.layernorm_fwd(x, res, g, b, y, mean, rstd, p_in=drop, site_in=2, seed=sd))
            t_b = graph_time(lambda: ops.layernorm_bwd(dy, x, res, g, mean, rstd, dx=dx, dres=dres if drop else None, dgamma=dg,
                                                       dbeta=db, dbias=dbias, p_in=drop, site_in=2, seed=sd))
            byt_f, byt_b = 3 * m * h * 2, (5 if drop else 4) * m * h * 2
            print(f"{name} m{m} h{h} drop{drop}: fwd {t_f:6.2f} us ({byt_f / t_f / 1e3:6.0f} GB/s)  bwd {t_b:6.2f} us ({byt_b / t_b / 1e3:6.0f} GB/s)", flush=True)
        tot += cnt * (t_f + t_b)
        xs = rnd(m, 3 * h)
        out = torch.zeros(3 * h, device="cuda")
        t_c = graph_time(lambda: ops.colsum(xs, out))
        print(f"colsum m{m} n{3 * h}: {t_c:6.2f} us ({m * 3 * h * 2 / t_c / 1e3:6.0f} GB/s)")
    print("LN us/step (dropout on):", round(tot, 1), flush=True)


def bench_attn():
    seed = torch.tensor([1234], dtype=torch.int64, device="cuda")
    B = 16
    tot = 0.0
    for name, heads, d, sq, sk, cnt in (("text self", 12, 64, 128, 128, 12), ("vis self", 8, 128, 100, 100, 6),
                                        ("co t->v", 8, 128, 128, 100, 6), ("co v->t", 8, 128, 100, 128, 6)):
        H = heads * d
        q, k, v = rnd(B * sq, H), rnd(B * sk, H), rnd(B * sk, H)
        out, dout = torch.empty_like(q), rnd(B * sq, H)
        dq, dk, dv = torch.empty_like(q), torch.empty_like(k), torch.empty_like(v)
        lse = torch.empty(B, heads, 128, device="cuda")
        mb = torch.zeros(B, sk, device="cuda")
        for drop in (0.0, 0.1):
            sd = seed if drop else None
            t_f = graph_time(lambda: ops.attention_fwd(q, k, v, out, lse, batch=B, heads=heads, sq=sq, sk=sk, d=d, mask_bias=mb,
                                                       p_drop=drop, site=4, seed=sd))
            t_b = graph_time(lambda: ops.attention_bwd(dout, q, k, v, lse, dq, dk, dv, batch=B, heads=heads, sq=sq, sk=sk, d=d,
                                                       mask_bias=mb, p_drop=drop, site=4, seed=sd))
            fl = 4.0 * B * heads * sq * sk * d
            print(f"attn {name:10s} h{heads} d{d} sq{sq} sk{sk} drop{drop}: fwd {t_f:6.2f} us ({fl / t_f / 1e6:5.0f} TF)  bwd {t_b:6.2f} us"
                  f" ({2.5 * fl / t_b / 1e6:5.0f} TF)", flush=True)
        tot += cnt * (t_f + t_b)
    print("attention us/step (dropout on):", round(tot, 1), flush=True)


def bench_misc():
    n = 224_000_000
    src = torch.randn(n, device="cuda")
    dst = torch.empty(n, device="cuda", dtype=BF)
    t = graph_time(lambda: ops.cast_bf16(src, dst), rep=2, iters=5)
    print(f"cast fp32->bf16 {n / 1e6:.0f}M: {t:7.1f} us ({n * 6 / t / 1e3:6.0f} GB/s)")
    g = torch.empty(25_000_000, device="cuda")
    t = graph_time(lambda: g.zero_(), rep=2, iters=5)
    print(f"zero 25M fp32: {t:7.1f} us")
    x = torch.zeros(8, device="cuda")
    t = graph_time(lambda: ops.seed_advance(torch.zeros(1, dtype=torch.int64, device="cuda")) if False else x.zero_(), rep=50)
    print(f"tiny kernel in graph: {t:5.2f} us per launch (launch floor)")


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default="")
    a = ap.parse_args()
    torch.cuda.set_device(0)
    if a.only in ("", "gemm"):
        bench_gemms()
    if a.only in ("", "ln"):
        bench_ln()
    if a.only in ("", "attn"):
        bench_attn()
    if a.only in ("", "misc"):
        bench_misc()
