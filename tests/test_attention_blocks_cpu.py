"""Host-side block algebra of attention over sequences above 128 (ops.attention_fwd / attention_bwd: block views, pointer
offsets, per-sample row strides, log-sum-exp merge, external delta, partial-gradient sums) on the CPU.  The five C-ABI entry
points are replaced by plain numpy restatements of their documented semantics that work on HOST memory through the very
pointers and strides ops.py passes, so every address computation of the host loop is exercised; the result is compared with
fp32 softmax attention and its autograd gradients.  The CUDA kernels themselves are covered by tests/test_attention_gpu.py."""
import ctypes as C
import math

import numpy as np
import pytest
import torch


def _bf16_view(ptr, rows, cols, ld):
    """float32 copy of a bf16 [rows, cols] view with row stride ld at host address ptr."""
    n = (rows - 1) * ld + cols
    raw = np.ctypeslib.as_array((C.c_uint16 * n).from_address(ptr))
    idx = (np.arange(rows)[:, None] * ld + np.arange(cols)[None, :])
    return (raw[idx].astype(np.uint32) << 16).view(np.float32)


def _bf16_store(ptr, ld, values, rows_ok=None):
    rows, cols = values.shape
    n = (rows - 1) * ld + cols
    raw = np.ctypeslib.as_array((C.c_uint16 * n).from_address(ptr))
    bits = torch.from_numpy(np.ascontiguousarray(values, np.float32)).to(torch.bfloat16).view(torch.int16).numpy().view(np.uint16)
    for r in range(rows if rows_ok is None else rows_ok):
        raw[r * ld:r * ld + cols] = bits[r]


def _f32(ptr, n):
    return np.ctypeslib.as_array((C.c_float * n).from_address(ptr))


class FakeLib:
    """numpy restatements of vb_attention_fwd / _bwd (one block of <= 128 x <= 128) and the three joining kernels."""

    def __init__(self):
        self.calls = []

    @staticmethod
    def _geometry(a):
        return (a.q_batch_rows or a.sq), (a.k_batch_rows or a.sk), (a.bias_ld or a.sk)

    def _scores(self, a, b, h):
        qr, kr, bl = self._geometry(a)
        d = a.d
        q = _bf16_view(a.q + (b * qr * a.ldq + h * d) * 2, a.sq, d, a.ldq)
        k = _bf16_view(a.k + (b * kr * a.ldk + h * d) * 2, a.sk, d, a.ldk)
        v = _bf16_view(a.v + (b * kr * a.ldv + h * d) * 2, a.sk, d, a.ldv)
        s = q @ k.T * a.scale
        if a.mask_bias:
            s = s + _f32(a.mask_bias + b * bl * 4, a.sk)[None, :]
        return q, k, v, s

    def vb_attention_fwd(self, ref, stream):
        a = ref._obj
        assert 1 <= a.sq <= 128 and 1 <= a.sk <= 128 and a.p_drop == 0.0
        self.calls.append(("fwd", a.sq, a.sk))
        qr, _, _ = self._geometry(a)
        for b in range(a.batch):
            for h in range(a.heads):
                q, k, v, s = self._scores(a, b, h)
                m = s.max(1, keepdims=True)
                e = np.exp(s - m)
                lse = (m[:, 0] + np.log(e.sum(1))).astype(np.float32)
                _f32(a.lse + ((b * a.heads + h) * 128) * 4, a.sq)[:] = lse
                _bf16_store(a.out + (b * qr * a.ldo + h * a.d) * 2, a.ldo, (e / e.sum(1, keepdims=True)) @ v)
        return 0

    def vb_attention_bwd(self, ref, stream):
        a = ref._obj
        assert 1 <= a.sq <= 128 and 1 <= a.sk <= 128
        self.calls.append(("bwd", a.sq, a.sk))
        qr, kr, _ = self._geometry(a)
        for b in range(a.batch):
            for h in range(a.heads):
                q, k, v, s = self._scores(a, b, h)
                lse = _f32(a.lse + ((b * a.heads + h) * 128) * 4, a.sq)
                p = np.exp(s - lse[:, None])
                do = _bf16_view(a.dout + (b * qr * a.lddo + h * a.d) * 2, a.sq, a.d, a.lddo)
                dp = do @ v.T
                delta = _f32(a.delta + ((b * a.heads + h) * 128) * 4, a.sq) if a.delta else (p * dp).sum(1)
                ds = p * (dp - delta[:, None]) * a.scale
                _bf16_store(a.dq + (b * qr * a.lddq + h * a.d) * 2, a.lddq, ds @ k)
                _bf16_store(a.dk + (b * kr * a.lddk + h * a.d) * 2, a.lddk, ds.T @ q)
                _bf16_store(a.dv + (b * kr * a.lddv + h * a.d) * 2, a.lddv, p.T @ do)
        return 0

    def vb_attn_merge(self, o_parts, lse_parts, n, ldp, out, ldo, lse_out, batch, heads, sq, batch_rows, d, stream):
        self.calls.append(("merge", n, sq))
        for b in range(batch):
            for h in range(heads):
                ls = np.stack([_f32(lse_parts[j] + ((b * heads + h) * 128) * 4, sq) for j in range(n)])
                m = ls.max(0)
                w = np.exp(ls - m)
                tot = w.sum(0)
                _f32(lse_out + ((b * heads + h) * 128) * 4, sq)[:] = m + np.log(tot)
                acc = sum((w[j] / tot)[:, None] * _bf16_view(o_parts[j] + (b * batch_rows * ldp + h * d) * 2, sq, d, ldp)
                          for j in range(n))
                _bf16_store(out + (b * batch_rows * ldo + h * d) * 2, ldo, acc)
        return 0

    def vb_attn_delta(self, out, ldo, dout, lddo, delta, batch, heads, sq, batch_rows, d, stream):
        self.calls.append(("delta", sq))
        for b in range(batch):
            for h in range(heads):
                o = _bf16_view(out + (b * batch_rows * ldo + h * d) * 2, sq, d, ldo)
                g = _bf16_view(dout + (b * batch_rows * lddo + h * d) * 2, sq, d, lddo)
                _f32(delta + ((b * heads + h) * 128) * 4, sq)[:] = (o * g).sum(1)
        return 0

    def vb_sum_rows_bf16(self, parts, n, ldp, dst, ldd, rows, width, stream):
        self.calls.append(("sum", n))
        _bf16_store(dst, ldd, sum(_bf16_view(parts[j], rows, width, ldp) for j in range(n)))
        return 0

    def vb_last_error(self):
        return b""


@pytest.fixture
def ops_on_host(monkeypatch):
    from multimodal_classification_b200 import _lib, ops
    fake = FakeLib()
    monkeypatch.setattr(_lib, "lib", lambda: fake)
    monkeypatch.setattr(ops, "_need_cuda", lambda *ts: None)
    monkeypatch.setattr(ops, "_stream", lambda: 0)
    monkeypatch.setattr(torch.cuda, "is_current_stream_capturing", lambda: False)
    ops._attn_scratch.clear()
    yield ops, fake
    ops._attn_scratch.clear()


def _reference(q, k, v, bias, b, heads, sq, sk, d):
    qh = q.float().view(b, sq, heads, d).permute(0, 2, 1, 3)
    kh = k.float().view(b, sk, heads, d).permute(0, 2, 1, 3)
    vh = v.float().view(b, sk, heads, d).permute(0, 2, 1, 3)
    s = qh @ kh.transpose(-1, -2) / math.sqrt(d)
    if bias is not None:
        s = s + bias.view(b, 1, 1, sk)
    return (torch.softmax(s, -1) @ vh).permute(0, 2, 1, 3).reshape(b * sq, heads * d)


def test_block_partition():
    from multimodal_classification_b200 import ops
    assert ops.attn_blocks(128) == [(0, 128)] and ops.attn_blocks(100) == [(0, 100)]
    assert ops.attn_blocks(257) == [(0, 86), (86, 86), (172, 85)] and ops.attn_blocks(129) == [(0, 65), (65, 64)]
    assert ops.attn_blocks(512) == [(0, 128), (128, 128), (256, 128), (384, 128)]
    assert ops.attn_lse_numel(2, 8, 257) == 3 * 2 * 8 * 128 and ops.attn_lse_numel(16, 12, 128) == 16 * 12 * 128


@pytest.mark.parametrize("b,heads,sq,sk,d,masked", [(2, 2, 257, 257, 64, True), (1, 2, 100, 257, 64, True),
                                                      (2, 1, 257, 100, 128, False), (1, 1, 300, 150, 64, True),
                                                      (2, 2, 100, 128, 64, True)])
def test_blocked_attention_matches_softmax_attention(ops_on_host, b, heads, sq, sk, d, masked):
    ops, fake = ops_on_host
    H = heads * d
    g = torch.Generator().manual_seed(sq * 1000 + sk)
    qbuf = (torch.randn(b * sq, 3 * H, generator=g) * 0.7).to(torch.bfloat16)      # q / k / v as column slices, as the encoder
    kvbuf = (torch.randn(b * sk, 3 * H, generator=g) * 0.7).to(torch.bfloat16)     # lays them out (row stride 3H)
    q, k, v = qbuf[:, :H], kvbuf[:, H:2 * H], kvbuf[:, 2 * H:]
    bias = None
    if masked:
        lens = torch.randint(1, sk + 1, (b,), generator=g)
        bias = ((torch.arange(sk).unsqueeze(0) >= lens.unsqueeze(1)).float() * -10000.0).contiguous()
    out = torch.zeros(b * sq, H, dtype=torch.bfloat16)
    lse = torch.zeros(ops.attn_lse_numel(b, heads, sq))
    ops.attention_fwd(q, k, v, out, lse, batch=b, heads=heads, sq=sq, sk=sk, d=d, mask_bias=bias)
    qr, kr, vr = (t.float().requires_grad_(True) for t in (q, k, v))
    ref = _reference(qr, kr, vr, bias, b, heads, sq, sk, d)
    assert (out.float() - ref).abs().max().item() <= 1.5e-2 * ref.abs().max().item()
    dout = torch.randn(b * sq, H, generator=g).to(torch.bfloat16)
    ref.backward(dout.float())
    dq_buf = torch.zeros(b * sq, 3 * H, dtype=torch.bfloat16)
    dkv_buf = torch.zeros(b * sk, 3 * H, dtype=torch.bfloat16)
    dq, dk, dv = dq_buf[:, :H], dkv_buf[:, H:2 * H], dkv_buf[:, 2 * H:]
    ops.attention_bwd(dout, q, k, v, lse, dq, dk, dv, batch=b, heads=heads, sq=sq, sk=sk, d=d, mask_bias=bias, out=out)
    for got, want in ((dq, qr.grad), (dk, kr.grad), (dv, vr.grad)):
        assert (got.float() - want).abs().max().item() <= 2e-2 * want.abs().max().item() + 2e-3
    assert dq_buf[:, H:].abs().max().item() == 0 and dkv_buf[:, :H].abs().max().item() == 0     # neighbours untouched
    nq, nk = len(ops.attn_blocks(sq)), len(ops.attn_blocks(sk))
    kinds = [c[0] for c in fake.calls]
    assert kinds.count("fwd") == kinds.count("bwd") == nq * nk
    assert kinds.count("merge") == (nq if nk > 1 else 0) and kinds.count("delta") == (nq if nk > 1 else 0)
    assert kinds.count("sum") == (1 if nk > 1 else 0) + (2 if nq > 1 else 0)
    # a second call finds its scratch again (what a captured graph relies on)
    n_scratch = len(ops._attn_scratch)
    ops.attention_fwd(q, k, v, out, lse, batch=b, heads=heads, sq=sq, sk=sk, d=d, mask_bias=bias)
    assert len(ops._attn_scratch) == n_scratch


def test_key_blocked_backward_needs_forward_output(ops_on_host):
    ops, _ = ops_on_host
    from multimodal_classification_b200._lib import VbError
    t = torch.zeros(200, 64, dtype=torch.bfloat16)
    lse = torch.zeros(ops.attn_lse_numel(1, 1, 200))
    with pytest.raises(VbError, match="forward output"):
        ops.attention_bwd(t, t, t, t, lse, t.clone(), t.clone(), t.clone(), batch=1, heads=1, sq=200, sk=200, d=64)
    big = torch.zeros(600, 64, dtype=torch.bfloat16)
    with pytest.raises(VbError, match="not supported"):
        ops.attention_fwd(big, big, big, big.clone(), torch.zeros(ops.attn_lse_numel(1, 1, 600)), batch=1, heads=1, sq=600,
                          sk=600, d=64)
