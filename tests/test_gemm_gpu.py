"""tcgen05 GEMM vs torch fp32 matmul on bf16-rounded inputs (per-kernel bar of SURVEY.md §8c:
max rel err <= 2e-2 with an absolute floor)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _rand(shape, seed):
    g = torch.Generator(device="cpu").manual_seed(seed)
    return (torch.randn(shape, generator=g) * 0.5).to(torch.bfloat16).cuda()


def _check(out, ref, tol=2e-2):
    out = out.float()
    err = (out - ref).abs().max().item()
    scale = ref.abs().max().item()
    assert err <= tol * scale + 1e-3, f"max abs err {err} vs scale {scale}"


SHAPES = [
    (2048, 768, 768), (2048, 2304, 768), (1600, 1024, 1024), (1600, 3072, 1024), (2048, 768, 3072),
    (576, 1024, 2048), (4112, 1024, 1024), (100, 64, 64), (128, 128, 64), (257, 192, 320), (300, 200, 136), (1600, 72, 24),
]


@pytest.mark.parametrize("m,n,k", SHAPES)
@pytest.mark.parametrize("block_n", [0, 64, 96, 128, 160, 192, 224, 256])
def test_forward_layout(m, n, k, block_n):
    from multimodal_classification_b200 import ops
    a, b = _rand((m, k), 1), _rand((n, k), 2)
    out = torch.empty(m, n, dtype=torch.bfloat16, device="cuda")
    ops.gemm(a, b, out, block_n=block_n)
    _check(out, a.float() @ b.float().t())


@pytest.mark.parametrize("m,n,k", SHAPES[:6])
def test_dgrad_layout(m, n, k):
    """dX[m,n] = dY[m,k] W[k,n]: B operand is MN-major."""
    from multimodal_classification_b200 import ops
    dy, w = _rand((m, k), 3), _rand((k, n), 4)
    out = torch.empty(m, n, dtype=torch.bfloat16, device="cuda")
    ops.gemm(dy, w, out, b_mn_major=True)
    _check(out, dy.float() @ w.float())


@pytest.mark.parametrize("m,n,k", [(768, 768, 2048), (3072, 768, 2048), (1024, 1024, 1600), (1024, 2048, 1600), (768, 3072, 4112), (64, 128, 100)])
@pytest.mark.parametrize("splits", [0, 1, 4])
def test_wgrad_layout(m, n, k, splits):
    """dW[m,n] = dY[k,m]^T X[k,n]: both operands MN-major, fp32 accumulate output."""
    from multimodal_classification_b200 import ops
    dy, x = _rand((k, m), 5), _rand((k, n), 6)
    out = torch.full((m, n), 1.0, dtype=torch.float32, device="cuda")
    ops.gemm(dy, x, out, a_mn_major=True, b_mn_major=True, accumulate=True, splits=splits)
    _check(out, dy.float().t() @ x.float() + 1.0)
    out2 = torch.full((m, n), 7.0, dtype=torch.float32, device="cuda")
    ops.gemm(dy, x, out2, a_mn_major=True, b_mn_major=True, accumulate=False)
    _check(out2, dy.float().t() @ x.float())


def test_a_mn_only():
    from multimodal_classification_b200 import ops
    a, b = _rand((320, 256), 7), _rand((192, 320), 8)  # a stored [K,M]
    out = torch.empty(256, 192, dtype=torch.bfloat16, device="cuda")
    ops.gemm(a, b, out, a_mn_major=True)
    _check(out, a.float().t() @ b.float().t())


def test_epilogues():
    from multimodal_classification_b200 import ops
    m, n, k = 1600, 1024, 1024
    a, b = _rand((m, k), 9), _rand((n, k), 10)
    bias = torch.randn(n, device="cuda")
    scale = torch.rand(n, device="cuda") + 0.5
    aux = _rand((m, n), 11)
    base = a.float() @ b.float().t()
    out = torch.empty(m, n, dtype=torch.bfloat16, device="cuda")
    pre = torch.empty_like(out)
    ops.gemm(a, b, out, bias=bias, act=ops.ACT_GELU, preact=pre)
    _check(pre, base + bias)
    _check(out, torch.nn.functional.gelu(base + bias))
    ops.gemm(a, b, out, bias=bias, scale=scale, aux=aux, aux_mode=ops.AUX_ADD, act=ops.ACT_RELU)
    _check(out, torch.relu(base * scale + bias + aux.float()))
    ops.gemm(a, b, out, bias=bias, act=ops.ACT_TANH)
    _check(out, torch.tanh(base + bias))
    x = aux.float().requires_grad_(True)
    torch.nn.functional.gelu(x).sum().backward()
    ops.gemm(a, b, out, aux=aux, aux_mode=ops.AUX_MUL_GELU_GRAD)
    _check(out, base * x.grad)
    # the forward can store GELU'(pre-activation) itself (one tanh serves the activation and its derivative); the backward GEMM
    # then only multiplies: both against exact F.gelu and its autograd derivative, for the compiled pair-CTA epilogues and for
    # the generic one (a single 128-row block runs as a lone CTA)
    for rows in (m, 96):
        z = (base[:rows] + bias).detach().requires_grad_(True)
        torch.nn.functional.gelu(z).sum().backward()
        gp = torch.empty(rows, n, dtype=torch.bfloat16, device="cuda")
        o2 = torch.empty(rows, n, dtype=torch.bfloat16, device="cuda")
        ops.gemm(a[:rows], b, o2, bias=bias, act=ops.ACT_GELU, preact=gp, preact_grad=True)
        _check(o2, torch.nn.functional.gelu(z.detach()))
        assert float((gp.float() - z.grad).abs().max()) <= 1e-2            # GELU' is O(1): bf16 storage + the f16x2 tanh
        ops.gemm(a[:rows], b, o2, aux=gp, aux_mode=ops.AUX_MUL)
        _check(o2, base[:rows] * gp.float())


def test_strided_views():
    """Operands that are column slices of a wider buffer (fused QKV output) must work through ld."""
    from multimodal_classification_b200 import ops
    m, k, n = 512, 256, 384
    big = _rand((m, 3 * k), 12)
    a = big[:, k:2 * k]
    b = _rand((n, k), 13)
    wide = torch.zeros(m, 2 * n, dtype=torch.bfloat16, device="cuda")
    ops.gemm(a, b, wide[:, n:])
    _check(wide[:, n:], a.float() @ b.float().t())
    assert wide[:, :n].abs().max().item() == 0


@pytest.mark.parametrize("max_ctas", [4, 8, 24])
@pytest.mark.parametrize("kind", ["fwd", "dgrad", "wgrad"])
def test_persistent_many_tiles_per_cta(max_ctas, kind):
    """A capped grid walks many tiles per CTA: both TMEM accumulator stages, the operand ring and every epilogue warp's
    staging box are re-used dozens of times (the warp-autonomous epilogue has no CTA-wide barrier to hide a missed wait)."""
    from multimodal_classification_b200 import ops
    m, n, k = 2048, 768, 768
    if kind == "fwd":
        a, b = _rand((m, k), 21), _rand((n, k), 22)
        bias = torch.randn(n, device="cuda")
        out, pre = torch.empty(m, n, dtype=torch.bfloat16, device="cuda"), torch.empty(m, n, dtype=torch.bfloat16, device="cuda")
        ops.gemm(a, b, out, bias=bias, act=ops.ACT_GELU, preact=pre, max_ctas=max_ctas)
        base = a.float() @ b.float().t() + bias
        _check(pre, base)
        _check(out, torch.nn.functional.gelu(base))
    elif kind == "dgrad":
        dy, w, aux = _rand((m, k), 23), _rand((k, n), 24), _rand((m, n), 25)
        out = torch.empty(m, n, dtype=torch.bfloat16, device="cuda")
        ops.gemm(dy, w, out, b_mn_major=True, aux=aux, aux_mode=ops.AUX_ADD, max_ctas=max_ctas)
        _check(out, dy.float() @ w.float() + aux.float())
    else:
        dy, x = _rand((k, m), 26), _rand((k, n), 27)
        out = torch.zeros(m, n, dtype=torch.float32, device="cuda")
        ops.gemm(dy, x, out, a_mn_major=True, b_mn_major=True, max_ctas=max_ctas)
        _check(out, dy.float().t() @ x.float())


def test_preact_with_aux_and_ragged_edges():
    """Pre-activation output and the aux operand together (they no longer share a staging panel), on a shape whose last
    row block and last column chunk are partial."""
    from multimodal_classification_b200 import ops
    m, n, k = 1000, 328, 200
    a, b, aux = _rand((m, k), 31), _rand((n, k), 32), _rand((m, n), 33)
    bias = torch.randn(n, device="cuda")
    out = torch.full((m + 8, n), 3.0, dtype=torch.bfloat16, device="cuda")      # guard rows: nothing may be written past m
    pre = torch.full((m + 8, n), 3.0, dtype=torch.bfloat16, device="cuda")
    for bn in (0, 96, 160, 256):
        ops.gemm(a, b, out[:m], bias=bias, aux=aux, aux_mode=ops.AUX_ADD, act=ops.ACT_RELU, preact=pre[:m], block_n=bn)
        base = a.float() @ b.float().t() + bias
        _check(pre[:m], base)
        _check(out[:m], torch.relu(base + aux.float()))
        assert (out[m:] == 3.0).all() and (pre[m:] == 3.0).all()
