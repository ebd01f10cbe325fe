"""tcgen05 fused attention (forward and backward) vs plain PyTorch fp32 on bf16-rounded inputs."""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu


def _bf(shape, seed, scale=1.0):
    g = torch.Generator(device="cpu").manual_seed(seed)
    return (torch.randn(shape, generator=g) * scale).to(torch.bfloat16).cuda()


def _close(out, ref, tol=2e-2, floor=2e-3):
    err = (out.float() - ref.float()).abs().max().item()
    scale = ref.float().abs().max().item()
    assert err <= tol * scale + floor, f"max abs err {err} (scale {scale})"


def _ref(q, k, v, bias, b, heads, sq, sk, d):
    qh = q.float().view(b, sq, heads, d).permute(0, 2, 1, 3)
    kh = k.float().view(b, sk, heads, d).permute(0, 2, 1, 3)
    vh = v.float().view(b, sk, heads, d).permute(0, 2, 1, 3)
    s = qh @ kh.transpose(-1, -2) / math.sqrt(d)
    if bias is not None:
        s = s + bias.view(b, 1, 1, sk)
    p = torch.softmax(s, -1)
    return (p @ vh).permute(0, 2, 1, 3).reshape(b * sq, heads * d)


CASES = [
    # batch, heads, sq, sk, d, masked
    (16, 12, 128, 128, 64, True),    # text self-attention
    (16, 8, 100, 100, 128, False),   # visual self-attention (LMDB: no visual mask)
    (16, 8, 100, 128, 128, True),    # co-attention: regions attend to tokens
    (16, 8, 128, 100, 128, True),    # co-attention: tokens attend to regions
    (3, 8, 36, 40, 128, True),       # ragged
    (2, 12, 40, 40, 64, True),
    (1, 1, 1, 1, 64, False),
    (2, 8, 128, 36, 128, False),
    # above 128 positions: balanced blocks joined by vb_attn_merge / vb_attn_delta / vb_sum_rows_bf16 (BASELINE config 4)
    (2, 8, 257, 257, 128, True),     # 257 DINOv2 tokens as regions: visual self-attention, 3 x 3 blocks
    (2, 8, 128, 257, 128, True),     # tokens attend to 257 regions: key blocks only
    (2, 8, 257, 128, 128, True),     # 257 regions attend to tokens: query blocks only
    (1, 12, 300, 200, 64, False),    # ragged, 64-wide heads
    (1, 2, 512, 129, 64, True),      # the maximum of four blocks
]


@pytest.mark.parametrize("b,heads,sq,sk,d,masked", CASES)
def test_attention_forward_backward(b, heads, sq, sk, d, masked):
    from multimodal_classification_b200 import ops
    H = heads * d
    # q, k, v as column slices of fused projection buffers (as the encoder lays them out)
    qbuf, kvbuf = _bf((b * sq, 3 * H), 1, 0.7), _bf((b * sk, 3 * H), 2, 0.7)
    q, k, v = qbuf[:, :H], kvbuf[:, H:2 * H], kvbuf[:, 2 * H:]
    bias = None
    if masked:
        lens = torch.randint(1, sk + 1, (b,), generator=torch.Generator().manual_seed(5))
        mask = (torch.arange(sk).unsqueeze(0) < lens.unsqueeze(1)).long().cuda()
        bias = torch.empty(b, sk, device="cuda")
        ops.mask_bias(mask, bias)
    out = torch.zeros(b * sq, H, dtype=torch.bfloat16, device="cuda")
    lse = torch.empty(ops.attn_lse_numel(b, heads, sq), device="cuda")
    ops.attention_fwd(q, k, v, out, lse, batch=b, heads=heads, sq=sq, sk=sk, d=d, mask_bias=bias)
    qr, kr, vr = (t.float().requires_grad_(True) for t in (q, k, v))
    ref = _ref(qr, kr, vr, bias, b, heads, sq, sk, d)
    _close(out, ref)
    dout = _bf((b * sq, H), 3)
    ref.backward(dout.float())
    dqkv_q = torch.zeros(b * sq, 3 * H, dtype=torch.bfloat16, device="cuda")
    dqkv_k = torch.zeros(b * sk, 3 * H, dtype=torch.bfloat16, device="cuda")
    dq, dk, dv = dqkv_q[:, :H], dqkv_k[:, H:2 * H], dqkv_k[:, 2 * H:]
    ops.attention_bwd(dout, q, k, v, lse, dq, dk, dv, batch=b, heads=heads, sq=sq, sk=sk, d=d, mask_bias=bias, out=out)
    _close(dq, qr.grad)
    _close(dk, kr.grad)
    _close(dv, vr.grad)
    assert dqkv_q[:, H:].abs().max().item() == 0 and dqkv_k[:, :H].abs().max().item() == 0


def test_attention_dropout_consistency():
    """Dropout on the probabilities: the expected output is unchanged, and backward regenerates the forward mask
    (checked through a directional finite difference of the kernel's own forward)."""
    from multimodal_classification_b200 import ops
    b, heads, s, d, p = 4, 8, 100, 128, 0.1
    H = heads * d
    q, k, v = _bf((b * s, H), 1, 0.5), _bf((b * s, H), 2, 0.5), _bf((b * s, H), 3)
    lse = torch.empty(b, heads, 128, device="cuda")
    seed = torch.tensor([4242], dtype=torch.int64, device="cuda")
    base = torch.empty(b * s, H, dtype=torch.bfloat16, device="cuda")
    ops.attention_fwd(q, k, v, base, lse, batch=b, heads=heads, sq=s, sk=s, d=d)
    acc = torch.zeros(b * s, H, device="cuda")
    n = 64
    o = torch.empty_like(base)
    for i in range(n):
        seed.fill_(1000 + i)
        ops.attention_fwd(q, k, v, o, lse, batch=b, heads=heads, sq=s, sk=s, d=d, p_drop=p, site=11, seed=seed)
        acc += o.float()
    _close(acc / n, base, tol=0.08, floor=0.02)
    # determinism for a fixed seed
    o2 = torch.empty_like(base)
    ops.attention_fwd(q, k, v, o2, lse, batch=b, heads=heads, sq=s, sk=s, d=d, p_drop=p, site=11, seed=seed)
    assert torch.equal(o, o2)
    # backward mask == forward mask: d/deps <out(v + eps*dv), dout> = <dV_kernel, dv>  (out is linear in v)
    dout = _bf((b * s, H), 4)
    dq, dk, dvk = (torch.empty(b * s, H, dtype=torch.bfloat16, device="cuda") for _ in range(3))
    ops.attention_bwd(dout, q, k, v, lse, dq, dk, dvk, batch=b, heads=heads, sq=s, sk=s, d=d, p_drop=p, site=11, seed=seed)
    dirv = _bf((b * s, H), 5)
    o3 = torch.empty_like(base)
    ops.attention_fwd(q, k, dirv, o3, lse, batch=b, heads=heads, sq=s, sk=s, d=d, p_drop=p, site=11, seed=seed)
    lhs = (o3.float() * dout.float()).sum().item()
    rhs = (dvk.float() * dirv.float()).sum().item()
    assert abs(lhs - rhs) <= 2e-2 * abs(lhs) + 1.0, (lhs, rhs)


def test_blocked_attention_dropout_is_consistent_between_forward_and_backward():
    """Sequences above 128 with dropout: the blocks' masks are regenerated by backward (directional finite difference of the
    kernel's own forward, as above) and the kept fraction is 1 - p."""
    from multimodal_classification_b200 import ops
    b, heads, sq, sk, d, p = 2, 8, 257, 200, 128, 0.1
    H = heads * d
    q, k, v = _bf((b * sq, H), 11, 0.5), _bf((b * sk, H), 12, 0.5), _bf((b * sk, H), 13, 1.0)
    seed = torch.tensor([1234], dtype=torch.int64, device="cuda")
    lse = torch.empty(ops.attn_lse_numel(b, heads, sq), device="cuda")

    def fwd(vv, pd):
        out = torch.empty(b * sq, H, dtype=torch.bfloat16, device="cuda")
        ops.attention_fwd(q, k, vv, out, lse, batch=b, heads=heads, sq=sq, sk=sk, d=d, p_drop=pd, site=7, seed=seed)
        return out
    out_d, out_0 = fwd(v, p), fwd(v, 0.0)
    assert torch.equal(out_d, fwd(v, p))                              # same seed and site: same masks
    rel = ((out_d.float() - out_0.float()).norm() / out_0.float().norm()).item()
    assert 0.02 < rel < 0.6, rel                                       # dropout changed the output, but not its scale
    # with v = ones the output is the kept probability mass / (1 - p): its mean is 1
    ones = torch.ones_like(v)
    assert abs(fwd(ones, p).float().mean().item() - 1.0) < 0.02
    # dv is linear in dout through the dropped probabilities: <dout, fwd(dv_dir)> == <dv, dv_dir>
    dout = _bf((b * sq, H), 14)
    dq, dk, dv = (torch.empty_like(t) for t in (q, k, v))
    out_d = fwd(v, p)
    ops.attention_bwd(dout, q, k, v, lse, dq, dk, dv, batch=b, heads=heads, sq=sq, sk=sk, d=d, p_drop=p, site=7, seed=seed,
                      out=out_d)
    direction = _bf((b * sk, H), 15)
    lhs = (dout.float() * fwd(direction, p).float()).sum().item()
    rhs = (dv.float() * direction.float()).sum().item()
    assert abs(lhs - rhs) <= 3e-2 * max(abs(lhs), abs(rhs)) + 1.0, (lhs, rhs)
