"""Row-wise kernels vs plain PyTorch fp32 on bf16-rounded inputs (per-kernel bar of SURVEY.md §8c)."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _bf(shape, seed, scale=1.0):
    g = torch.Generator(device="cpu").manual_seed(seed)
    return (torch.randn(shape, generator=g) * scale).to(torch.bfloat16).cuda()


def _close(out, ref, tol=2e-2, floor=1e-3):
    err = (out.float() - ref.float()).abs().max().item()
    scale = ref.float().abs().max().item()
    assert err <= tol * scale + floor, f"max abs err {err} (scale {scale})"


@pytest.mark.parametrize("m,h", [(2048, 768), (1600, 1024), (37, 768), (576, 2048), (5, 256)])
@pytest.mark.parametrize("with_res", [True, False])
def test_layernorm_fwd_bwd(m, h, with_res):
    from multimodal_classification_b200 import ops
    x, res = _bf((m, h), 1), (_bf((m, h), 2) if with_res else None)
    gamma = (1 + 0.1 * torch.randn(h)).cuda()
    beta = (0.1 * torch.randn(h)).cuda()
    y = torch.empty_like(x)
    mean, rstd = torch.empty(m, device="cuda"), torch.empty(m, device="cuda")
    ops.layernorm_fwd(x, res, gamma, beta, y, mean, rstd)
    xr = x.float().requires_grad_(True)
    rr = res.float().requires_grad_(True) if with_res else None
    gr, br = gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
    s = xr + rr if with_res else xr
    u = s.mean(-1, keepdim=True)
    var = (s - u).pow(2).mean(-1, keepdim=True)
    ref = gr * ((s - u) / torch.sqrt(var + 1e-12)) + br
    _close(y, ref)
    dy = _bf((m, h), 3)
    ref.backward(dy.float())
    dx = torch.empty_like(x)
    dres = torch.empty_like(x) if with_res else None
    dg, db, dbias = (torch.zeros(h, device="cuda") for _ in range(3))
    ops.layernorm_bwd(dy, x, res, gamma, mean, rstd, dx=dx, dres=dres, dgamma=dg, dbeta=db, dbias=dbias)
    _close(dx, xr.grad)
    if with_res:
        _close(dres, rr.grad)
    _close(dg, gr.grad, tol=1e-2, floor=1e-2)
    _close(db, br.grad, tol=1e-2, floor=1e-2)
    _close(dbias, xr.grad.sum(0), tol=1e-2, floor=5e-2)


def test_layernorm_dropout_statistics_and_mask_reuse():
    from multimodal_classification_b200 import ops
    m, h, p = 1024, 768, 0.1
    x, res = _bf((m, h), 4), _bf((m, h), 5)
    gamma, beta = torch.ones(h, device="cuda"), torch.zeros(h, device="cuda")
    seed = torch.tensor([12345], dtype=torch.int64, device="cuda")
    mean, rstd = torch.empty(m, device="cuda"), torch.empty(m, device="cuda")
    # output dropout: zero fraction ~ p, survivors scaled by 1/(1-p)
    y0, y1 = torch.empty_like(x), torch.empty_like(x)
    ops.layernorm_fwd(x, res, gamma, beta, y0, mean, rstd)
    ops.layernorm_fwd(x, res, gamma, beta, y1, mean, rstd, p_out=p, site_out=7, seed=seed)
    zero = (y1 == 0)
    frac = zero.float().mean().item()
    assert abs(frac - p) < 0.01, frac
    keep = ~zero
    _close(y1[keep], y0.float()[keep] / (1 - p), tol=1e-2)
    # the same (seed, site) reproduces the mask; another site or seed does not
    y2 = torch.empty_like(x)
    ops.layernorm_fwd(x, res, gamma, beta, y2, mean, rstd, p_out=p, site_out=7, seed=seed)
    assert torch.equal(y1, y2)
    ops.layernorm_fwd(x, res, gamma, beta, y2, mean, rstd, p_out=p, site_out=8, seed=seed)
    assert not torch.equal(y1, y2)
    # backward through output dropout uses the same mask: dx is zero-gradient where... check dres against autograd
    dy = _bf((m, h), 6)
    mask = keep.float() / (1 - p)
    s = (x.float() + res.float()).requires_grad_(True)
    ref = F.layer_norm(s, (h,), gamma, beta, 1e-12) * mask
    ref.backward(dy.float())
    dres = torch.empty_like(x)
    ops.layernorm_bwd(dy, x, res, gamma, mean, rstd, dres=dres, p_out=p, site_out=7, seed=seed)
    _close(dres, s.grad)
    # input dropout: dx is masked, dres is not
    y3 = torch.empty_like(x)
    ops.layernorm_fwd(x, res, gamma, beta, y3, mean, rstd, p_in=p, site_in=9, seed=seed)
    dx, dres = torch.empty_like(x), torch.empty_like(x)
    ops.layernorm_bwd(dy, x, res, gamma, mean, rstd, dx=dx, dres=dres, p_in=p, site_in=9, seed=seed)
    zin = (dx == 0) & (dres != 0)
    assert abs(zin.float().mean().item() - p) < 0.01
    km = ~zin
    _close(dx[km], dres.float()[km] / (1 - p), tol=2e-2)
    # forward consistent with that mask
    s3 = (x.float() * km.float() / (1 - p) + res.float())
    _close(y3, F.layer_norm(s3, (h,), gamma, beta, 1e-12))


@pytest.mark.parametrize("b,t", [(16, 128), (3, 40)])
def test_embed_text_fwd_bwd(b, t):
    from multimodal_classification_b200 import ops
    h, vocab = 768, 1000
    g = torch.Generator(device="cpu").manual_seed(0)
    word = (torch.randn(vocab, h, generator=g) * 0.02).cuda()
    pos = (torch.randn(512, h, generator=g) * 0.02).cuda()
    typ = (torch.randn(2, h, generator=g) * 0.02).cuda()
    gamma, beta = (1 + 0.1 * torch.randn(h, generator=g)).cuda(), (0.1 * torch.randn(h, generator=g)).cuda()
    ids = torch.randint(0, vocab, (b, t), generator=g)
    ids[:, t // 2:] = 0  # PAD tail (padding_idx)
    ids[0, :4] = 5       # duplicates
    tt = torch.randint(0, 2, (b, t), generator=g)
    ids_d, tt_d = ids.int().cuda(), tt.int().cuda()
    y = torch.empty(b * t, h, dtype=torch.bfloat16, device="cuda")
    mean, rstd = torch.empty(b * t, device="cuda"), torch.empty(b * t, device="cuda")
    ops.embed_text_fwd(ids_d.view(-1), tt_d.view(-1), word, pos, typ, gamma, beta, y, mean, rstd, b, t)
    wr, pr, tr = (x.clone().requires_grad_(True) for x in (word, pos, typ))
    gr, br = gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
    e = F.embedding(ids.cuda(), wr, padding_idx=0) + tr[tt.cuda()] + pr[:t].unsqueeze(0)
    ref = F.layer_norm(e, (h,), gr, br, 1e-12)
    _close(y.view(b, t, h), ref)
    dy = _bf((b * t, h), 9)
    ref.backward(dy.float().view(b, t, h))
    dword, dpos, dtyp = torch.zeros_like(word), torch.zeros_like(pos), torch.zeros_like(typ)
    dg, db = torch.zeros(h, device="cuda"), torch.zeros(h, device="cuda")
    ops.embed_text_bwd(dy, ids_d.view(-1), tt_d.view(-1), word, pos, typ, gamma, mean, rstd, b, t, dword=dword,
                       dpos=dpos, dtype=dtyp, dgamma=dg, dbeta=db)
    _close(dword, wr.grad, tol=1e-3, floor=1e-3)
    assert dword[0].abs().max().item() == 0.0
    _close(dpos, pr.grad, tol=1e-3, floor=1e-3)
    _close(dtyp, tr.grad, tol=1e-3, floor=1e-2)
    _close(dg, gr.grad, tol=1e-3, floor=1e-2)
    _close(db, br.grad, tol=1e-3, floor=1e-2)


def test_colsum_cast_mask_misc():
    from multimodal_classification_b200 import ops
    x = _bf((1600, 3072), 11)
    out = torch.zeros(3072, device="cuda")
    ops.colsum(x, out)
    _close(out, x.float().sum(0), tol=1e-3, floor=1e-2)
    view = x[:, 1024:2048]
    out2 = torch.ones(1024, device="cuda")
    ops.colsum(view, out2)
    _close(out2, view.float().sum(0) + 1, tol=1e-3, floor=1e-2)
    src = torch.randn(100003, device="cuda")
    dst = torch.empty(100003, dtype=torch.bfloat16, device="cuda")
    ops.cast_bf16(src, dst)
    assert torch.equal(dst, src.to(torch.bfloat16))
    # multi-tensor cast
    srcs = [torch.randn(n, device="cuda") for n in (8, 8192, 8200, 768 * 768, 24)]
    dsts = [torch.zeros(n, dtype=torch.bfloat16, device="cuda") for n in (8, 8192, 8200, 768 * 768, 24)]
    mc = ops.MultiCast(list(zip(srcs, dsts)), "cuda")
    mc.run()
    for s, d in zip(srcs, dsts):
        assert torch.equal(d, s.to(torch.bfloat16))
    # masks: bit-exact with the reference expression for every accepted dtype
    m64 = torch.randint(0, 2, (16, 128)).cuda()
    for m in (m64, m64.int(), m64.float()):
        o = torch.empty(16, 128, device="cuda")
        ops.mask_bias(m, o)
        assert torch.equal(o, (1.0 - m) * -10000.0)
    ids = torch.tensor([0, 5, 30521, 7], dtype=torch.int64).cuda()
    o32 = torch.empty(4, dtype=torch.int32, device="cuda")
    flag = torch.zeros(1, dtype=torch.int32, device="cuda")
    ops.i64_to_i32(ids, o32, 0, 30522, flag)
    assert o32.tolist() == [0, 5, 30521, 7] and flag.item() == 0
    ops.i64_to_i32(ids, o32, 0, 30521, flag)
    assert flag.item() == 1


def test_dropout_and_activation_backward():
    from multimodal_classification_b200 import ops
    x = _bf((16, 2048), 21)
    y = torch.empty_like(x)
    seed = torch.tensor([99], dtype=torch.int64, device="cuda")
    ops.dropout(x, y, 0.1, 3, seed)
    z = (y == 0) & (x != 0)
    assert abs(z.float().mean().item() - 0.1) < 0.02
    _close(y[~z], x.float()[~z] / 0.9, tol=1e-2)
    dy = _bf((16, 2048), 22)
    t = torch.tanh(x.float()).to(torch.bfloat16)
    dx = torch.empty_like(x)
    ops.act_bwd(dy, t, dx, ops.ACT_TANH)
    _close(dx, dy.float() * (1 - t.float() ** 2))
    r = torch.relu(x)
    ops.act_bwd(dy, r, dx, ops.ACT_RELU)
    _close(dx, dy.float() * (r > 0).float())


def test_loc_embed_and_classifier_tail():
    from multimodal_classification_b200 import ops
    m, n = 1600, 1024
    loc = torch.rand(m, 5, device="cuda")
    w, b = torch.randn(n, 5, device="cuda") * 0.4, torch.randn(n, device="cuda") * 0.1
    out = torch.empty(m, n, dtype=torch.bfloat16, device="cuda")
    ops.loc_embed_fwd(loc, w, b, out)
    _close(out, F.linear(loc, w, b), tol=1e-2)
    ds = _bf((m, n), 31)
    dw, db = torch.zeros_like(w), torch.zeros_like(b)
    ops.loc_embed_bwd(ds, loc, dw, db)
    _close(dw, ds.float().t() @ loc, tol=1e-3, floor=1e-2)
    _close(db, ds.float().sum(0), tol=1e-3, floor=1e-2)
    # classifier tail + CE
    bsz, k, c = 16, 1024, 2
    h = _bf((bsz, k), 32)
    w4 = (torch.randn(c, k, device="cuda") * 0.03).requires_grad_(True)
    b4 = (torch.randn(c, device="cuda") * 0.1).requires_grad_(True)
    labels = torch.randint(0, c, (bsz,), device="cuda")
    logits, probs, loss = torch.empty(bsz, c, device="cuda"), torch.empty(bsz, c, device="cuda"), torch.empty(1, device="cuda")
    ops.cls_ce_fwd(h, w4.detach(), b4.detach(), labels.int(), logits, probs, loss)
    hr = h.float().requires_grad_(True)
    ref_logits = F.linear(hr, w4, b4)
    ref_loss = F.cross_entropy(ref_logits, labels)
    _close(logits, ref_logits, tol=1e-4, floor=1e-5)
    assert abs(loss.item() - ref_loss.item()) < 1e-5
    ext = torch.randn(bsz, c, device="cuda") * 0.1
    (ref_loss * 0.7 + (ref_logits * ext).sum()).backward()
    dloss = torch.tensor([0.7], device="cuda")
    dw, db, dh = torch.empty_like(w4), torch.empty_like(b4), torch.empty_like(h)
    ops.cls_ce_bwd(h, w4.detach(), labels.int(), probs, dloss, ext, dw, db, dh)
    _close(dw, w4.grad, tol=1e-4, floor=1e-5)
    _close(db, b4.grad, tol=1e-4, floor=1e-5)
    _close(dh, hr.grad, tol=1e-2, floor=1e-5)
