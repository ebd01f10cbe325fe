"""CPU functional stand-ins for the kernel wrappers of multimodal_classification_b200/ops.py that the ViLBERT engine calls
(dropout off), written from the semantics documented in include/vilbert_b200.h with plain torch fp32 arithmetic on the
tensors the engine hands over (bf16 storage is kept: results are rounded into the engine's own buffers).  They let the
engine's host schedule — which buffer feeds which kernel on which stream, the flat parameter / gradient layout, the fp32
residual ring, the backward order — run against the oracle in the GPU-less container.  TEST INFRASTRUCTURE ONLY: nothing
here is importable from the package, and the product path still refuses to run without CUDA."""
import contextlib
import math

import torch
import torch.nn.functional as F

ACT_NONE, ACT_GELU, ACT_RELU, ACT_TANH = 0, 1, 2, 3
AUX_NONE, AUX_ADD, AUX_MUL_GELU_GRAD, AUX_MUL = 0, 1, 2, 3


IDENTITY_DROPOUT = False     # smoke mode: let the training-mode code paths run with dropout acting as the identity


def _no_dropout(*ps):
    assert IDENTITY_DROPOUT or all(p == 0.0 for p in ps), "the simulator covers the dropout-off configuration"


def gemm(a, b, out, *, a_mn_major=False, b_mn_major=False, bias=None, scale=None, aux=None, aux_mode=AUX_NONE, act=ACT_NONE,
         preact=None, accumulate=False, block_n=0, splits=0, max_ctas=0, b_streamed=False, d_streamed=False, conv=None,
         preact_grad=False):
    if conv is not None:                                          # implicit-GEMM convolution: a is the NHWC activation
        kh, kw, stride, pad = conv
        n, h, w, c = a.shape
        ho, wo = (h + 2 * pad - kh) // stride + 1, (w + 2 * pad - kw) // stride + 1
        cols = F.unfold(a.float().permute(0, 3, 1, 2), (kh, kw), padding=pad, stride=stride)
        a = cols.view(n, c, kh * kw, ho * wo).permute(0, 3, 2, 1).reshape(n * ho * wo, kh * kw * c)
    am = a.float().t() if a_mn_major else a.float()               # [M, K]
    bm = b.float().t() if b_mn_major else b.float()               # [N, K]
    v = am @ bm.t()
    if scale is not None:
        v = v * scale
    if bias is not None:
        v = v + bias
    if preact is not None:
        if preact_grad:       # GELU'(v) = Phi(v) + v phi(v)
            preact.copy_(0.5 * (1.0 + torch.erf(v / math.sqrt(2.0))) + v * torch.exp(-0.5 * v * v) / math.sqrt(2.0 * math.pi))
        else:
            preact.copy_(v)
    if aux_mode == AUX_ADD:
        v = v + aux.float()
    elif aux_mode == AUX_MUL:
        v = v * aux.float()
    elif aux_mode == AUX_MUL_GELU_GRAD:
        x = aux.float()
        v = v * (0.5 * (1.0 + torch.erf(x / math.sqrt(2.0))) + x * torch.exp(-0.5 * x * x) / math.sqrt(2.0 * math.pi))
    if act == ACT_GELU:
        v = F.gelu(v)
    elif act == ACT_RELU:
        v = F.relu(v)
    elif act == ACT_TANH:
        v = torch.tanh(v)
    if accumulate:
        assert out.dtype == torch.float32
        out.add_(v)
    else:
        out.copy_(v)
    return out


def _ln_input(x, res, res32):
    s = x.float()
    if res32 is not None:
        s = s + res32
    elif res is not None:
        s = s + res.float()
    return s


def layernorm_fwd(x, res, gamma, beta, y, mean, rstd, *, eps=1e-12, p_in=0.0, site_in=0, p_out=0.0, site_out=0, seed=None,
                  res32=None, y32=None):
    _no_dropout(p_in, p_out)
    s = _ln_input(x, res, res32)
    mu = s.mean(-1)
    rs = torch.rsqrt(s.var(-1, unbiased=False) + eps)
    v = (s - mu[:, None]) * rs[:, None] * gamma + beta
    y.copy_(v)
    if y32 is not None:
        y32.copy_(v)
    mean.copy_(mu)
    rstd.copy_(rs)
    return y


def layernorm_bwd(dy, x, res, gamma, mean, rstd, *, dx=None, dres=None, dgamma=None, dbeta=None, dbias=None, eps=1e-12,
                  p_in=0.0, site_in=0, p_out=0.0, site_out=0, seed=None, res32=None):
    _no_dropout(p_in, p_out)
    s = _ln_input(x, res, res32)
    xhat = (s - mean[:, None]) * rstd[:, None]
    g = dy.float()
    gg = g * gamma
    d = rstd[:, None] * (gg - gg.mean(-1, keepdim=True) - xhat * (gg * xhat).mean(-1, keepdim=True))
    if dx is not None:
        dx.copy_(d)
    if dres is not None:
        dres.copy_(d)
    if dgamma is not None:
        dgamma.add_((g * xhat).sum(0))
    if dbeta is not None:
        dbeta.add_(g.sum(0))
    if dbias is not None:
        dbias.add_(d.sum(0))


def _emb_sum(ids, type_ids, word, pos, typ, b, t):
    idx = ids.long()
    e = word[idx] + (typ[type_ids.long()] if type_ids is not None else typ[0])
    return e + pos[:t].repeat(b, 1)


def embed_text_fwd(ids, type_ids, word, pos, typ, gamma, beta, y, mean, rstd, b, t, *, eps=1e-12, p_out=0.0, site_out=0,
                   seed=None, y32=None):
    _no_dropout(p_out)
    s = _emb_sum(ids, type_ids, word, pos, typ, b, t)
    mu = s.mean(-1)
    rs = torch.rsqrt(s.var(-1, unbiased=False) + eps)
    v = (s - mu[:, None]) * rs[:, None] * gamma + beta
    y.copy_(v)
    if y32 is not None:
        y32.copy_(v)
    mean.copy_(mu)
    rstd.copy_(rs)
    return y


def embed_text_bwd(dy, ids, type_ids, word, pos, typ, gamma, mean, rstd, b, t, *, dword=None, dpos=None, dtype=None,
                   dgamma=None, dbeta=None, eps=1e-12, p_out=0.0, site_out=0, seed=None):
    _no_dropout(p_out)
    s = _emb_sum(ids, type_ids, word, pos, typ, b, t)
    xhat = (s - mean[:, None]) * rstd[:, None]
    g = dy.float()
    gg = g * gamma
    d = rstd[:, None] * (gg - gg.mean(-1, keepdim=True) - xhat * (gg * xhat).mean(-1, keepdim=True))
    if dword is not None:
        keep = (ids != 0).float()[:, None]                       # padding_idx = 0 receives no gradient
        dword.index_add_(0, ids.long(), d * keep)
    if dpos is not None:
        dpos[:t].add_(d.view(b, t, -1).sum(0))
    if dtype is not None:
        tt = type_ids.long() if type_ids is not None else torch.zeros_like(ids, dtype=torch.long)
        dtype.index_add_(0, tt, d)
    if dgamma is not None:
        dgamma.add_((g * xhat).sum(0))
    if dbeta is not None:
        dbeta.add_(g.sum(0))


def colsum(x, out):
    out.add_(x.float().sum(0))
    return out


def cast_bf16(src, dst):
    dst.copy_(src)
    return dst


def cast_f32(src, dst):
    dst.copy_(src)
    return dst


def mask_bias(mask, out):
    out.view(-1).copy_(((1.0 - mask.float()) * -10000.0).view(-1))
    return out


def i64_to_i32(src, dst, lo, hi, err_flag=None):
    v = src.reshape(-1)
    dst.copy_(torch.where((v < lo) | (v >= hi), torch.full_like(v, lo), v))
    return dst


def dropout(x, y, p, site, seed):
    assert IDENTITY_DROPOUT, "dropout kernel reached with dropout off"
    y.copy_(x)
    return y


def seed_advance(seed, snapshot=None):
    seed.mul_(6364136223846793005).add_(1442695040888963407)     # int64 wrap-around = the kernel's uint64 LCG
    if snapshot is not None:
        snapshot.copy_(seed)


def stage_batch(segs, err_flag=None):
    """vb_stage_batch: range-checked index conversion (ignore_index -100 legal for labels), mask -> additive bias, feature
    rounding, box copy; violations OR their bit into the flag and store `lo`."""
    from multimodal_classification_b200 import _lib as L
    for kind, src, dst, lo, hi, bit in segs:
        v = src.reshape(-1)
        if kind == L.STAGE_INDEX:
            ok = (v >= lo) & (v < hi)
            if bit == L.STAGE_ERR_LABEL:
                ok |= v == L.IGNORE_INDEX
            if err_flag is not None and not bool(ok.all()):
                err_flag._host[0] |= bit
            dst.view(-1).copy_(torch.where(ok, v, torch.full_like(v, lo)))
        elif kind == L.STAGE_MASK:
            dst.view(-1).copy_((1.0 - v.float()) * -10000.0)
        else:
            dst.view(-1).copy_(v)


def act_bwd(dy, y, dx, act):
    yv = y.float()
    dx.copy_(dy.float() * ((1.0 - yv * yv) if act == ACT_TANH else (yv > 0).float()))
    return dx


def loc_embed_fwd(loc, w, b, out):
    out.copy_(loc @ w.t() + b)
    return out


def loc_embed_bwd(ds, loc, dw, db):
    g = ds.float()
    dw.add_(g.t() @ loc)
    db.add_(g.sum(0))


def cls_ce_fwd(h, w, bias, labels, logits, probs, loss):
    z = h.float() @ w.t() + bias
    logits.copy_(z)
    probs.copy_(torch.softmax(z, -1))
    loss.fill_(0.0 if labels is None else F.cross_entropy(z, labels.long()).item())     # ignore_index = -100, as the kernel


def cls_ce_bwd(h, w, labels, probs, dloss, dlogits_ext, dw, db, dh):
    bsz = probs.shape[0]
    dz = dlogits_ext.clone() if dlogits_ext is not None else torch.zeros_like(probs)
    if labels is not None:
        valid = (labels != -100)
        onehot = F.one_hot(labels.long().clamp(min=0), probs.shape[1]).float()
        dz = dz + dloss * (probs - onehot) * valid[:, None].float() / valid.sum()
    dw.copy_(dz.t() @ h.float())
    db.copy_(dz.sum(0))
    dh.copy_(dz @ w)


def _heads(x, batch, s, heads, d):
    return x.float().reshape(batch, s, heads, d).permute(0, 2, 1, 3)


def _probs(q, k, batch, heads, sq, sk, d, mask_bias_t, scale):
    sc = _heads(q, batch, sq, heads, d) @ _heads(k, batch, sk, heads, d).transpose(-1, -2) * scale
    if mask_bias_t is not None:
        sc = sc + mask_bias_t.view(batch, 1, 1, sk)
    return torch.softmax(sc, -1)


def attention_fwd(q, k, v, out, lse, *, batch, heads, sq, sk, d, mask_bias=None, scale=None, p_drop=0.0, site=0, seed=None):
    _no_dropout(p_drop)
    scale = 1.0 / math.sqrt(d) if scale is None else scale
    p = _probs(q, k, batch, heads, sq, sk, d, mask_bias, scale)
    out.copy_((p @ _heads(v, batch, sk, heads, d)).permute(0, 2, 1, 3).reshape(batch * sq, heads * d))
    return out


def attention_bwd(dout, q, k, v, lse, dq, dk, dv, *, batch, heads, sq, sk, d, mask_bias=None, scale=None, p_drop=0.0, site=0,
                  seed=None, out=None):
    _no_dropout(p_drop)
    scale = 1.0 / math.sqrt(d) if scale is None else scale
    p = _probs(q, k, batch, heads, sq, sk, d, mask_bias, scale)
    qh, kh, vh = _heads(q, batch, sq, heads, d), _heads(k, batch, sk, heads, d), _heads(v, batch, sk, heads, d)
    do = _heads(dout, batch, sq, heads, d)
    dp = do @ vh.transpose(-1, -2)
    ds = p * (dp - (p * dp).sum(-1, keepdim=True)) * scale

    def flat(t, s):
        return t.permute(0, 2, 1, 3).reshape(batch * s, heads * d)
    dq.copy_(flat(ds @ kh, sq))
    dk.copy_(flat(ds.transpose(-1, -2) @ qh, sk))
    dv.copy_(flat(p.transpose(-1, -2) @ do, sk))


# ------------------------------------------------------------------------------------------------ RoI / grid feature stage
def stem_im2col(img, out, kh=7, kw=7, stride=2, pad=3):
    n, c, h, w = img.shape
    ho, wo = (h + 2 * pad - kh) // stride + 1, (w + 2 * pad - kw) // stride + 1
    cols = F.unfold(img, (kh, kw), padding=pad, stride=stride).view(n, c, kh * kw, ho * wo).permute(0, 3, 2, 1)
    out.zero_()
    out[:, : kh * kw * c].copy_(cols.reshape(n * ho * wo, kh * kw * c))          # column (ky*kw + kx)*3 + ci, zero padded
    return out


def im2col_nhwc(x, out, kh, kw, stride, pad):
    n, h, w, c = x.shape
    ho, wo = (h + 2 * pad - kh) // stride + 1, (w + 2 * pad - kw) // stride + 1
    cols = F.unfold(x.float().permute(0, 3, 1, 2), (kh, kw), padding=pad, stride=stride)
    out.copy_(cols.view(n, c, kh * kw, ho * wo).permute(0, 3, 2, 1).reshape(n * ho * wo, kh * kw * c))
    return out


def maxpool_nhwc(x, out, k=3, stride=2, pad=1):
    out.copy_(F.max_pool2d(x.float().permute(0, 3, 1, 2), k, stride, pad).permute(0, 2, 3, 1))
    return out


def roi_pool_nhwc(x, rois, out, spatial_scale, argmax=None):
    import torchvision
    r, ph, pw, c = out.shape
    out.copy_(torchvision.ops.roi_pool(x.float().permute(0, 3, 1, 2).contiguous(), rois, (ph, pw), spatial_scale).permute(0, 2, 3, 1))
    return out


def roi_align_nhwc(x, rois, out, spatial_scale, sampling_ratio=2, aligned=False):
    import torchvision
    r, ph, pw, c = out.shape
    out.copy_(torchvision.ops.roi_align(x.float().permute(0, 3, 1, 2).contiguous(), rois, (ph, pw), spatial_scale, sampling_ratio,
                                        aligned).permute(0, 2, 3, 1))
    return out


def avgpool_nhwc(x, out):
    out.copy_(x.float().mean(1))
    return out


def box_area_score(boxes, img_w, img_h, scores, target=0.15):
    w = (boxes[:, 2] - boxes[:, 0]) / img_w
    h = (boxes[:, 3] - boxes[:, 1]) / img_h
    scores.copy_(1.0 - torch.abs(w * h - target))
    return scores


def nms(boxes, scores, iou_threshold):
    import torchvision
    return torchvision.ops.nms(boxes, scores, iou_threshold)


def nms_device(boxes, scores, iou_threshold, workspace, keep, num_keep):
    import torchvision
    k = torchvision.ops.nms(boxes, scores, iou_threshold)
    keep[: k.numel()] = k.to(torch.int32)
    num_keep.fill_(k.numel())
    return keep, num_keep


def rowmax(x, out, col_begin=0, col_end=None):
    out.copy_(x[:, col_begin:col_end].max(dim=1)[0])
    return out


def select_regions(candidates, keep, num_keep, regions, img_w, img_h, *, boxes=None, spatial=None, index=None, feat_src=None,
                   feat_dst=None, rois=None, batch_index=0, box_div=1.0):
    nk = int(num_keep.reshape(-1)[0])
    if nk <= 0:
        return
    idx = keep[torch.arange(regions).clamp(max=nk - 1)].long()
    b = candidates[idx]
    if boxes is not None:
        boxes.copy_(b)
    if index is not None:
        index.copy_(idx.to(torch.int32))
    if rois is not None:
        rois[:, 0] = float(batch_index)
        rois[:, 1:] = b
    if spatial is not None:
        nb = b / box_div
        nb[:, [0, 2]] /= img_w
        nb[:, [1, 3]] /= img_h
        nb = nb.clamp(0, 1)
        spatial.copy_(torch.cat([nb, ((nb[:, 2] - nb[:, 0]) * (nb[:, 3] - nb[:, 1])).unsqueeze(1)], dim=1))
    if feat_src is not None:
        feat_dst.copy_(feat_src[idx])


def rpn_decode(heads, fh, fw, base_anchors, stride, img_h, img_w, min_size, boxes, scores, num_valid):
    """fasterrcnn_vg_rpn.py:78-174 + the min-size test of :444-450 in torch."""
    a = base_anchors.shape[0]
    cls = heads[:, : 2 * a].reshape(-1, 2)
    deltas = heads[:, 2 * a: 6 * a].reshape(-1, 4)
    fg = F.softmax(cls, dim=-1)[:, 1]
    sx = torch.arange(0, fw) * stride + stride // 2
    sy = torch.arange(0, fh) * stride + stride // 2
    sy, sx = torch.meshgrid(sy, sx, indexing="ij")
    shifts = torch.stack([sx, sy, sx, sy], dim=-1).reshape(-1, 4)
    anchors = (torch.from_numpy(base_anchors).unsqueeze(0) + shifts.unsqueeze(1)).reshape(-1, 4)
    w, h = anchors[:, 2] - anchors[:, 0], anchors[:, 3] - anchors[:, 1]
    cx, cy = anchors[:, 0] + 0.5 * w, anchors[:, 1] + 0.5 * h
    pcx, pcy = deltas[:, 0] * w + cx, deltas[:, 1] * h + cy
    pw, ph = torch.exp(deltas[:, 2].clamp(max=4.0)) * w, torch.exp(deltas[:, 3].clamp(max=4.0)) * h
    p = torch.stack([pcx - 0.5 * pw, pcy - 0.5 * ph, pcx + 0.5 * pw, pcy + 0.5 * ph], dim=-1)
    p[:, 0::2] = p[:, 0::2].clamp(0, img_w)
    p[:, 1::2] = p[:, 1::2].clamp(0, img_h)
    ok = ((p[:, 2] - p[:, 0]) >= min_size) & ((p[:, 3] - p[:, 1]) >= min_size)
    boxes.copy_(p)
    scores.copy_(torch.where(ok, fg, torch.full_like(fg, float("-inf"))))
    num_valid.fill_(int(ok.sum()))


def rank_sort_desc(scores, order, limit=None):
    s = scores.clone()
    if limit is not None:
        s[int(limit.reshape(-1)[0]):] = float("-inf")
    order[: s.numel()] = torch.sort(s, descending=True, stable=True)[1].to(torch.int32)
    return order


def gather_sorted(boxes, scores, order, num_valid, out_boxes, out_scores, count):
    cap = out_boxes.shape[0]
    n = min(cap, int(num_valid.reshape(-1)[0]))
    idx = order[:n].long()
    out_boxes.zero_()
    out_scores.fill_(float("-inf"))
    out_boxes[:n] = boxes[idx]
    out_scores[:n] = scores[idx]
    count.fill_(n)


def nms_sorted(boxes, count, iou_threshold, keep, num_keep):
    import torchvision
    n = int(count.reshape(-1)[0])
    k = torchvision.ops.nms(boxes[:n], torch.arange(n, 0, -1, dtype=torch.float32), iou_threshold)[: keep.numel()]
    keep[: k.numel()] = k.to(torch.int32)
    num_keep.fill_(k.numel())


ROI_SIMULATED = ["stem_im2col", "im2col_nhwc", "maxpool_nhwc", "roi_pool_nhwc", "roi_align_nhwc", "avgpool_nhwc", "box_area_score",
                 "nms", "nms_device", "rowmax", "select_regions", "rpn_decode", "rank_sort_desc", "gather_sorted", "nms_sorted"]
SIMULATED = ["gemm", "layernorm_fwd", "layernorm_bwd", "embed_text_fwd", "embed_text_bwd", "colsum", "cast_bf16", "cast_f32", "mask_bias",
             "i64_to_i32", "stage_batch", "dropout", "seed_advance", "act_bwd", "loc_embed_fwd", "loc_embed_bwd", "cls_ce_fwd", "cls_ce_bwd",
             "attention_fwd", "attention_bwd"]


class _Stream:
    cuda_stream = 0

    def __init__(self, *a, **k):
        pass

    def wait_stream(self, other):
        pass

    def wait_event(self, event):
        pass

    def record_event(self, event=None):
        return event


class _Event:
    def __init__(self, *a, **k):
        pass

    def record(self, stream=None):
        pass

    def synchronize(self):
        pass

    def wait(self, stream=None):
        pass


def install(monkeypatch):
    """Route the engine's kernel wrappers to the stand-ins above and make the CUDA runtime objects it touches inert."""
    from multimodal_classification_b200 import ops
    here = globals()
    for name in SIMULATED + ROI_SIMULATED:
        monkeypatch.setattr(ops, name, here[name])
    main = _Stream()
    for name, value in [("Stream", _Stream), ("Event", _Event), ("current_stream", lambda d=None: main),
                        ("stream", lambda s: contextlib.nullcontext()), ("device", lambda d: contextlib.nullcontext()),
                        ("synchronize", lambda d=None: None), ("is_current_stream_capturing", lambda: False)]:
        monkeypatch.setattr(torch.cuda, name, value)
    monkeypatch.setattr(torch.Tensor, "is_cuda", property(lambda self: True))
    monkeypatch.setenv("VB_NO_GRAPH", "1")


class PlainPatch:
    """monkeypatch-shaped setter for spawned worker processes (no undo needed there)."""

    @staticmethod
    def setattr(obj, name, value):
        setattr(obj, name, value)

    @staticmethod
    def setenv(name, value):
        import os
        os.environ[name] = value


def install_device_shims(monkeypatch):
    """For modules that move themselves or allocate with an explicit "cuda" device: keep everything on the CPU."""
    orig_mto, orig_tto = torch.nn.Module.to, torch.Tensor.to

    def is_cuda_arg(a):
        return bool(a) and isinstance(a[0], (str, torch.device)) and str(a[0]).startswith("cuda")
    monkeypatch.setattr(torch.nn.Module, "to", lambda self, *a, **k: self if is_cuda_arg(a) else orig_mto(self, *a, **k))
    monkeypatch.setattr(torch.Tensor, "to", lambda self, *a, **k: self if is_cuda_arg(a) else orig_tto(self, *a, **k))

    def on_cpu(fn):
        def wrapped(*a, **k):
            if "device" in k and str(k["device"]).startswith("cuda"):
                k["device"] = "cpu"
            return fn(*a, **k)
        return wrapped
    for name in ("zeros", "empty", "arange", "tensor", "full"):
        monkeypatch.setattr(torch, name, on_cpu(getattr(torch, name)))
