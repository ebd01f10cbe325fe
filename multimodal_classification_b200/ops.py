"""Thin tensor-level wrappers over the C ABI (one Python function per exported kernel).

Tensors are only used for their device pointers, shapes and the current CUDA stream; no torch
compute op is issued here.  Every wrapper raises on a non-zero status.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

import torch

from . import _lib
from ._lib import ACT_GELU, ACT_NONE, ACT_RELU, ACT_TANH, AUX_ADD, AUX_MUL, AUX_MUL_GELU_GRAD, AUX_NONE, GemmArgs  # noqa: F401


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _ptr(t):
    return None if t is None else t.data_ptr()


def _need_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise _lib.VbError("libvilbert_b200 kernels need CUDA tensors; there is no CPU fallback")


def gemm(a: torch.Tensor, b: torch.Tensor, out: torch.Tensor, *, a_mn_major=False, b_mn_major=False,
         bias=None, scale=None, aux=None, aux_mode=AUX_NONE, act=ACT_NONE, preact=None, accumulate=False,
         block_n=0, splits=0, max_ctas=0, b_streamed=False, d_streamed=False, conv=None, preact_grad=False) -> torch.Tensor:
    """out[M,N] = epilogue(sum_k A(m,k) B(n,k)).

    conv=(kh, kw, stride, pad): `a` is a contiguous NHWC bf16 activation [n, h, w, c] (c % 64 == 0) and the product is the
    convolution with the weight b [N, kh*kw*c] (column (ky*kw + kx)*c + ci): out[n*ho*wo, N].  The A tiles are gathered by TMA in
    im2col mode -- no [pixels, kh*kw*c] matrix is materialised (nn.Conv2d of the ResNet trunk, resnet152_roi.py:49-74).

    a: [M,K] (K-major) or [K,M] (a_mn_major); b: [N,K] or [K,N] (b_mn_major); bf16, last dim contiguous.
    out: bf16 or fp32 [M,N].  Mirrors nn.Linear forward / dgrad / wgrad
    (reference models/vilbert_facebook_arch.py:127-129 and autograd thereof).
    """
    _need_cuda(a, b, out, bias, scale, aux, preact)
    assert a.dtype == torch.bfloat16 and b.dtype == torch.bfloat16
    assert a.stride(-1) == 1 and b.stride(-1) == 1 and out.stride(-1) == 1
    cn = ch = cw = cc = ckh = ckw = cstride = cpad = 0
    if conv is not None:
        assert not a_mn_major and not b_mn_major and a.dim() == 4 and a.is_contiguous()
        ckh, ckw, cstride, cpad = conv
        cn, ch, cw, cc = a.shape
        m = cn * ((ch + 2 * cpad - ckh) // cstride + 1) * ((cw + 2 * cpad - ckw) // cstride + 1)
        k = ckh * ckw * cc
    elif a_mn_major:
        k, m = a.shape
    else:
        m, k = a.shape
    if b_mn_major:
        kb, n = b.shape
    else:
        n, kb = b.shape
    assert kb == k, (a.shape, b.shape)
    assert tuple(out.shape) == (m, n), (out.shape, m, n)
    args = GemmArgs()
    args.a, args.b, args.d = a.data_ptr(), b.data_ptr(), out.data_ptr()
    args.d_preact = _ptr(preact)
    args.scale, args.bias, args.aux = _ptr(scale), _ptr(bias), _ptr(aux)
    if bias is not None:
        assert bias.dtype == torch.float32 and bias.numel() == n
    if scale is not None:
        assert scale.dtype == torch.float32 and scale.numel() == n
    args.lda, args.ldb, args.ldd = (0 if conv is not None else a.stride(0)), b.stride(0), out.stride(0)
    args.conv_n, args.conv_h, args.conv_w, args.conv_c = cn, ch, cw, cc
    args.conv_kh, args.conv_kw, args.conv_stride, args.conv_pad = ckh, ckw, cstride, cpad
    args.preact_grad = int(preact_grad)      # `preact` receives GELU'(pre-activation) instead of the pre-activation
    args.ld_preact = preact.stride(0) if preact is not None else 0
    args.ld_aux = aux.stride(0) if aux is not None else 0
    args.m, args.n, args.k = m, n, k
    args.a_mn_major, args.b_mn_major = int(a_mn_major), int(b_mn_major)
    args.d_is_f32 = int(out.dtype == torch.float32)
    assert out.dtype in (torch.float32, torch.bfloat16)
    args.accumulate = int(accumulate)
    args.act, args.aux_mode = act, aux_mode
    args.block_n, args.splits, args.max_ctas = block_n, splits, max_ctas
    args.b_streamed, args.d_streamed = int(b_streamed), int(d_streamed)
    _lib.check(_lib.lib().vb_gemm_bf16(C.byref(args), _stream()), "vb_gemm_bf16")
    return out


# ------------------------------------------------------------------------------------------------ row-wise kernels
def _ld(t):
    return 0 if t is None else t.stride(0)


def _ln_args(x, res, gamma, beta, mean, rstd, eps, p_in, site_in, p_out, site_out, seed, res32=None):
    _need_cuda(x, res, gamma, beta, mean, rstd, seed, res32)
    m, h = x.shape
    assert x.dtype == torch.bfloat16 and x.stride(1) == 1
    a = _lib.LayerNormArgs()
    a.x, a.res, a.gamma, a.beta = x.data_ptr(), _ptr(res), gamma.data_ptr(), _ptr(beta)
    a.mean, a.rstd = mean.data_ptr(), rstd.data_ptr()
    a.ldx, a.ldres = x.stride(0), _ld(res)
    a.m, a.h, a.eps = m, h, eps
    a.p_in, a.p_out, a.site_in, a.site_out = p_in, p_out, site_in, site_out
    a.seed = _ptr(seed)
    if res32 is not None:
        assert res32.dtype == torch.float32 and res32.shape == x.shape and res32.stride(1) == 1
        a.res_f32, a.ldres_f32 = res32.data_ptr(), res32.stride(0)
    return a


def layernorm_fwd(x, res, gamma, beta, y, mean, rstd, *, eps=1e-12, p_in=0.0, site_in=0, p_out=0.0, site_out=0,
                  seed=None, res32=None, y32=None):
    """y = dropout_out(LN(dropout_in(x) + res)); BertLayerNorm + the dropout / residual before it
    (reference models/vilbert_facebook_arch.py:63-76, 156-160, 197-201, 329-336)."""
    a = _ln_args(x, res, gamma, beta, mean, rstd, eps, p_in, site_in, p_out, site_out, seed, res32)
    _need_cuda(y, y32)
    a.y, a.ldy = y.data_ptr(), y.stride(0)
    if y32 is not None:
        assert y32.dtype == torch.float32 and y32.shape == x.shape and y32.stride(1) == 1
        a.y_f32, a.ldy_f32 = y32.data_ptr(), y32.stride(0)
    _lib.check(_lib.lib().vb_layernorm_fwd(C.byref(a), _stream()), "vb_layernorm_fwd")
    return y


def layernorm_bwd(dy, x, res, gamma, mean, rstd, *, dx=None, dres=None, dgamma=None, dbeta=None, dbias=None,
                  eps=1e-12, p_in=0.0, site_in=0, p_out=0.0, site_out=0, seed=None, res32=None):
    """Gradients of layernorm_fwd; dgamma / dbeta / dbias are accumulated atomically (zero them first)."""
    a = _ln_args(x, res, gamma, None, mean, rstd, eps, p_in, site_in, p_out, site_out, seed, res32)
    _need_cuda(dy, dx, dres, dgamma, dbeta, dbias)
    a.dy, a.lddy = dy.data_ptr(), dy.stride(0)
    a.dx, a.lddx, a.dres, a.lddres = _ptr(dx), _ld(dx), _ptr(dres), _ld(dres)
    a.dgamma, a.dbeta, a.dbias = _ptr(dgamma), _ptr(dbeta), _ptr(dbias)
    _lib.check(_lib.lib().vb_layernorm_bwd(C.byref(a), _stream()), "vb_layernorm_bwd")


def _emb_args(ids, type_ids, word, pos, typ, gamma, beta, mean, rstd, b, t, eps, p_out, site_out, seed):
    _need_cuda(ids, type_ids, word, pos, typ, gamma, beta, mean, rstd, seed)
    assert ids.dtype == torch.int32 and (type_ids is None or type_ids.dtype == torch.int32)
    assert word.dtype == torch.float32 and word.is_contiguous() and pos.is_contiguous() and typ.is_contiguous()
    a = _lib.EmbedArgs()
    a.ids, a.type_ids, a.word, a.pos, a.type = ids.data_ptr(), _ptr(type_ids), word.data_ptr(), pos.data_ptr(), typ.data_ptr()
    a.gamma, a.beta, a.mean, a.rstd = gamma.data_ptr(), _ptr(beta), mean.data_ptr(), rstd.data_ptr()
    a.b, a.t, a.h, a.vocab = b, t, word.shape[1], word.shape[0]
    a.eps, a.p_out, a.site_out, a.seed = eps, p_out, site_out, _ptr(seed)
    assert pos.shape[0] >= t
    return a


def embed_text_fwd(ids, type_ids, word, pos, typ, gamma, beta, y, mean, rstd, b, t, *, eps=1e-12, p_out=0.0,
                   site_out=0, seed=None, y32=None):
    """transformers BertEmbeddings.forward as called at models/vilbert_facebook_arch.py:524."""
    a = _emb_args(ids, type_ids, word, pos, typ, gamma, beta, mean, rstd, b, t, eps, p_out, site_out, seed)
    a.y = y.data_ptr()
    if y32 is not None:
        assert y32.dtype == torch.float32 and y32.is_contiguous() and y32.shape == y.shape
        a.y_f32 = y32.data_ptr()
    _lib.check(_lib.lib().vb_embed_text_fwd(C.byref(a), _stream()), "vb_embed_text_fwd")
    return y


def embed_text_bwd(dy, ids, type_ids, word, pos, typ, gamma, mean, rstd, b, t, *, dword=None, dpos=None, dtype=None,
                   dgamma=None, dbeta=None, eps=1e-12, p_out=0.0, site_out=0, seed=None):
    a = _emb_args(ids, type_ids, word, pos, typ, gamma, None, mean, rstd, b, t, eps, p_out, site_out, seed)
    _need_cuda(dy, dword, dpos, dtype, dgamma, dbeta)
    a.dy = dy.data_ptr()
    a.dword, a.dpos, a.dtype, a.dgamma, a.dbeta = _ptr(dword), _ptr(dpos), _ptr(dtype), _ptr(dgamma), _ptr(dbeta)
    _lib.check(_lib.lib().vb_embed_text_bwd(C.byref(a), _stream()), "vb_embed_text_bwd")


def colsum(x, out):
    """out[n] += sum_m x[m, n] (bias gradient)."""
    _need_cuda(x, out)
    assert x.dtype == torch.bfloat16 and out.dtype == torch.float32 and x.stride(1) == 1
    _lib.check(_lib.lib().vb_colsum_bf16(x.data_ptr(), x.stride(0), x.shape[0], x.shape[1], out.data_ptr(), _stream()),
               "vb_colsum_bf16")
    return out


def cast_bf16(src, dst):
    _need_cuda(src, dst)
    assert src.dtype == torch.float32 and dst.dtype == torch.bfloat16 and src.is_contiguous() and dst.is_contiguous()
    assert src.numel() == dst.numel()
    _lib.check(_lib.lib().vb_cast_f32_bf16(src.data_ptr(), dst.data_ptr(), src.numel(), _stream()), "vb_cast_f32_bf16")
    return dst


def cast_f32(src, dst):
    _need_cuda(src, dst)
    assert src.dtype == torch.bfloat16 and dst.dtype == torch.float32 and src.is_contiguous() and dst.is_contiguous()
    assert src.numel() == dst.numel()
    _lib.check(_lib.lib().vb_cast_bf16_f32(src.data_ptr(), dst.data_ptr(), src.numel(), _stream()), "vb_cast_bf16_f32")
    return dst


class MultiCast:
    """One launch that refreshes every bf16 weight shadow from its fp32 master (pointer table on the device)."""
    CHUNK = 256 * 8 * 4

    def __init__(self, pairs, device):
        import numpy as np
        segs = np.zeros((len(pairs), 3), dtype=np.int64)
        block_seg, block_off = [], []
        self._keep = pairs
        for i, (src, dst) in enumerate(pairs):
            assert src.dtype == torch.float32 and dst.dtype == torch.bfloat16 and src.is_contiguous() and dst.is_contiguous()
            assert src.numel() == dst.numel() and src.is_cuda and dst.is_cuda
            assert src.data_ptr() % 16 == 0 and dst.data_ptr() % 16 == 0
            segs[i] = (src.data_ptr(), dst.data_ptr(), src.numel())
            for off in range(0, src.numel(), self.CHUNK):
                block_seg.append(i)
                block_off.append(off)
        self.src_ptrs = [s.data_ptr() for s, _ in pairs]
        self.segs = torch.from_numpy(segs).to(device)
        self.block_seg = torch.tensor(block_seg, dtype=torch.int32, device=device)
        self.block_off = torch.tensor(block_off, dtype=torch.int64, device=device)

    def valid(self):
        return all(s.data_ptr() == p for (s, _), p in zip(self._keep, self.src_ptrs))

    def run(self):
        _lib.check(_lib.lib().vb_cast_f32_bf16_multi(self.segs.data_ptr(), self.block_seg.data_ptr(),
                                                     self.block_off.data_ptr(), self.block_seg.numel(), _stream()),
                   "vb_cast_f32_bf16_multi")


_MASK_DT = {torch.float32: _lib.DT_F32, torch.int32: _lib.DT_I32, torch.int64: _lib.DT_I64}


def mask_bias(mask, out):
    """(1.0 - mask) * -10000.0, reference models/vilbert_facebook_arch.py:530-540."""
    _need_cuda(mask, out)
    if mask.dtype not in _MASK_DT:
        raise _lib.VbError(f"attention mask dtype {mask.dtype} is not supported (int64, int32, float32)")
    assert mask.is_contiguous() and out.dtype == torch.float32 and out.numel() == mask.numel()
    _lib.check(_lib.lib().vb_mask_bias(mask.data_ptr(), _MASK_DT[mask.dtype], out.data_ptr(), mask.numel(), _stream()),
               "vb_mask_bias")
    return out


def i64_to_i32(src, dst, lo, hi, err_flag=None):
    _need_cuda(src, dst, err_flag)
    assert src.dtype == torch.int64 and dst.dtype == torch.int32 and src.is_contiguous()
    _lib.check(_lib.lib().vb_i64_to_i32(src.data_ptr(), dst.data_ptr(), src.numel(), lo, hi, _ptr(err_flag), _stream()),
               "vb_i64_to_i32")
    return dst


def dropout(x, y, p, site, seed):
    _need_cuda(x, y, seed)
    assert x.is_contiguous() and y.is_contiguous() and x.dtype == torch.bfloat16
    _lib.check(_lib.lib().vb_dropout_bf16(x.data_ptr(), y.data_ptr(), x.numel(), p, site, seed.data_ptr(), _stream()),
               "vb_dropout_bf16")
    return y


def seed_advance(seed, snapshot=None):
    """Advance the engine's dropout seed; with `snapshot`, also record the new value in a plan-owned buffer."""
    _need_cuda(seed, snapshot)
    if snapshot is None:
        _lib.check(_lib.lib().vb_seed_advance(seed.data_ptr(), _stream()), "vb_seed_advance")
    else:
        _lib.check(_lib.lib().vb_seed_advance_to(seed.data_ptr(), snapshot.data_ptr(), _stream()), "vb_seed_advance_to")


_STAGE_DT = {torch.float32: _lib.DT_F32, torch.int32: _lib.DT_I32, torch.int64: _lib.DT_I64, torch.bfloat16: _lib.DT_BF16}


def stage_batch(segs, err_flag=None):
    """One launch for the whole batch hand-over (vb_stage_batch).  `segs`: [(kind, src, dst, lo, hi, err_bit)]; sources must be
    contiguous CUDA tensors of a dtype their kind accepts."""
    arr = (_lib.StageSeg * len(segs))()
    for i, (kind, src, dst, lo, hi, bit) in enumerate(segs):
        _need_cuda(src, dst)
        if src.dtype not in _STAGE_DT:
            raise _lib.VbError(f"input dtype {src.dtype} is not supported by the staging kernel")
        if kind == _lib.STAGE_INDEX and src.dtype not in (torch.int64, torch.int32):
            raise _lib.VbError(f"ids / token types / labels must be int64 or int32, got {src.dtype}")
        if kind == _lib.STAGE_MASK and src.dtype == torch.bfloat16:
            raise _lib.VbError("attention mask dtype torch.bfloat16 is not supported (int64, int32, float32)")
        if kind in (_lib.STAGE_FEAT,) and src.dtype not in (torch.float32, torch.bfloat16):
            raise _lib.VbError(f"region features must be float32 or bfloat16, got {src.dtype}")
        if kind == _lib.STAGE_COPY_F32 and src.dtype != torch.float32:
            raise _lib.VbError(f"spatial locations must be float32, got {src.dtype}")
        assert src.is_contiguous() and dst.is_contiguous() and src.numel() == dst.numel(), (kind, src.shape, dst.shape)
        a = arr[i]
        a.src, a.dst, a.n, a.kind, a.dtype = src.data_ptr(), dst.data_ptr(), src.numel(), kind, _STAGE_DT[src.dtype]
        a.lo, a.hi, a.err_bit = lo, hi, bit
    _lib.check(_lib.lib().vb_stage_batch(arr, len(segs), _ptr(err_flag), _stream()), "vb_stage_batch")


def act_bwd(dy, y, dx, act):
    _need_cuda(dy, y, dx)
    assert dy.is_contiguous() and y.is_contiguous() and dx.is_contiguous()
    _lib.check(_lib.lib().vb_act_bwd_bf16(dy.data_ptr(), y.data_ptr(), dx.data_ptr(), dy.numel(), act, _stream()),
               "vb_act_bwd_bf16")
    return dx


def loc_embed_fwd(loc, w, b, out):
    _need_cuda(loc, w, b, out)
    assert loc.dtype == torch.float32 and loc.is_contiguous() and w.is_contiguous() and out.is_contiguous()
    m, k = loc.shape
    _lib.check(_lib.lib().vb_loc_embed_fwd(loc.data_ptr(), w.data_ptr(), b.data_ptr(), out.data_ptr(), m, w.shape[0], k,
                                           _stream()), "vb_loc_embed_fwd")
    return out


def loc_embed_bwd(ds, loc, dw, db):
    _need_cuda(ds, loc, dw, db)
    assert ds.is_contiguous() and loc.is_contiguous()
    m, k = loc.shape
    _lib.check(_lib.lib().vb_loc_embed_bwd(ds.data_ptr(), loc.data_ptr(), dw.data_ptr(), _ptr(db), m, ds.shape[1], k,
                                           _stream()), "vb_loc_embed_bwd")


def cls_ce_fwd(h, w, bias, labels, logits, probs, loss):
    _need_cuda(h, w, bias, labels, logits, probs, loss)
    assert h.is_contiguous() and w.is_contiguous() and h.dtype == torch.bfloat16 and w.dtype == torch.float32
    _lib.check(_lib.lib().vb_cls_ce_fwd(h.data_ptr(), w.data_ptr(), bias.data_ptr(), _ptr(labels), logits.data_ptr(),
                                        probs.data_ptr(), _ptr(loss), h.shape[0], h.shape[1], w.shape[0], _stream()),
               "vb_cls_ce_fwd")


def cls_ce_bwd(h, w, labels, probs, dloss, dlogits_ext, dw, db, dh):
    _need_cuda(h, w, labels, probs, dloss, dlogits_ext, dw, db, dh)
    _lib.check(_lib.lib().vb_cls_ce_bwd(h.data_ptr(), w.data_ptr(), _ptr(labels), probs.data_ptr(), _ptr(dloss),
                                        _ptr(dlogits_ext), _ptr(dw), _ptr(db), dh.data_ptr(), h.shape[0], h.shape[1],
                                        w.shape[0], _stream()), "vb_cls_ce_bwd")


# ------------------------------------------------------------------------------------------------ attention
ATTN_BLOCK = 128            # queries / keys one fused-attention CTA handles
ATTN_MAX_SEQ = 4 * ATTN_BLOCK   # vb_attn_merge / vb_sum_rows_bf16 join at most four blocks
_BLOCK_SITE_STRIDE = 1 << 16  # dropout sites of the blocks of one attention call (engine sites are far below this)
_attn_scratch = {}            # (kind, output data_ptr, geometry) -> buffers that outlive the call (CUDA-graph replays)


def attn_blocks(s: int):
    """[(start, length)] of the balanced blocks of <= 128 a sequence of s positions is processed in."""
    n = (s + ATTN_BLOCK - 1) // ATTN_BLOCK
    base, rem = divmod(s, n)
    out, start = [], 0
    for i in range(n):
        ln = base + (1 if i < rem else 0)
        out.append((start, ln))
        start += ln
    return out


def attn_lse_numel(batch: int, heads: int, sq: int) -> int:
    """Elements of the fp32 log-sum-exp buffer forward writes and backward reads: [query blocks, batch, heads, 128]."""
    return len(attn_blocks(sq)) * batch * heads * 128


def _attn_args(q, k, v, lse, mask_bias_t, batch, heads, sq, sk, d, scale, p_drop, site, seed):
    _need_cuda(q, k, v, lse, mask_bias_t, seed)
    for t in (q, k, v):
        assert t.dtype == torch.bfloat16 and t.stride(-1) == 1 and t.dim() == 2
    assert q.shape[0] == batch * sq and k.shape[0] == batch * sk and v.shape[0] == batch * sk
    assert q.shape[1] == heads * d and k.shape[1] == heads * d and v.shape[1] == heads * d
    assert lse.dtype == torch.float32 and lse.numel() == attn_lse_numel(batch, heads, sq)
    if sq > ATTN_MAX_SEQ or sk > ATTN_MAX_SEQ:
        raise _lib.VbError(f"attention over more than {ATTN_MAX_SEQ} positions is not supported (sq={sq}, sk={sk})")
    a = _lib.AttnArgs()
    a.q, a.k, a.v = q.data_ptr(), k.data_ptr(), v.data_ptr()
    a.ldq, a.ldk, a.ldv = q.stride(0), k.stride(0), v.stride(0)
    a.lse, a.mask_bias = lse.data_ptr(), _ptr(mask_bias_t)
    if mask_bias_t is not None:
        assert mask_bias_t.dtype == torch.float32 and mask_bias_t.numel() == batch * sk and mask_bias_t.is_contiguous()
    a.batch, a.heads, a.sq, a.sk, a.d = batch, heads, sq, sk, d
    a.scale, a.p_drop, a.site, a.seed = scale, p_drop, site, _ptr(seed)
    return a


def _block_view(a, full, i, j, q0, ql, k0, kl, sq, sk, nkb):
    """Copy of the whole-sequence argument block `full`, narrowed to query block i (rows q0 .. q0+ql) x key block j."""
    C.memmove(C.byref(a), C.byref(full), C.sizeof(_lib.AttnArgs))
    a.q = full.q + q0 * full.ldq * 2
    a.k = full.k + k0 * full.ldk * 2
    a.v = full.v + k0 * full.ldv * 2
    a.sq, a.sk, a.q_batch_rows, a.k_batch_rows, a.bias_ld = ql, kl, sq, sk, sk
    if full.mask_bias:
        a.mask_bias = full.mask_bias + k0 * 4
    a.site = full.site + (i * nkb + j + 1) * _BLOCK_SITE_STRIDE          # an independent dropout stream per block
    return a


def _ptr_array(ptrs):
    return (C.c_void_p * len(ptrs))(*ptrs)


def _scratch(kind, anchor, shapes):
    """Persistent scratch of one attention call site (keyed by its output tensor): allocated on the eager warm-up step,
    found again on every later call, so that captured graphs keep valid addresses."""
    key = (kind, anchor.data_ptr(), tuple(shapes))
    got = _attn_scratch.get(key)
    if got is None:
        if torch.cuda.is_current_stream_capturing():
            raise _lib.VbError("attention scratch requested for the first time during graph capture (warm-up step missing)")
        got = _attn_scratch[key] = [torch.empty(shape, dtype=dt, device=anchor.device) for shape, dt in shapes]
    return got


def attention_fwd(q, k, v, out, lse, *, batch, heads, sq, sk, d, mask_bias=None, scale=None, p_drop=0.0, site=0,
                  seed=None):
    """out = dropout(softmax(q k^T * scale + mask_bias)) v per (sample, head); q/k/v/out are 2-D strided views
    [batch*seq, heads*d] (reference models/vilbert_facebook_arch.py:126-144 and :253-294).  Sequences above 128 run as
    balanced blocks of <= 128 queries x <= 128 keys joined by vb_attn_merge; lse is [query blocks, batch, heads, 128]."""
    scale = (1.0 / d ** 0.5) if scale is None else scale
    a = _attn_args(q, k, v, lse, mask_bias, batch, heads, sq, sk, d, scale, p_drop, site, seed)
    _need_cuda(out)
    assert out.shape == q.shape and out.stride(1) == 1 and out.dtype == torch.bfloat16
    a.out, a.ldo = out.data_ptr(), out.stride(0)
    if sq <= ATTN_BLOCK and sk <= ATTN_BLOCK:
        _lib.check(_lib.lib().vb_attention_fwd(C.byref(a), _stream()), "vb_attention_fwd")
        return out
    qb, kb = attn_blocks(sq), attn_blocks(sk)
    nkb, width, blk = len(kb), heads * d, batch * heads * 128
    parts = lses = None
    if nkb > 1:      # per key block: a full-size partial output and its log-sum-exp for every query block
        parts, lses = _scratch("fwd", out, [((nkb, batch * sq, width), torch.bfloat16), ((nkb, len(qb), blk), torch.float32)])
    b = _lib.AttnArgs()
    for i, (q0, ql) in enumerate(qb):
        for j, (k0, kl) in enumerate(kb):
            _block_view(b, a, i, j, q0, ql, k0, kl, sq, sk, nkb)
            if nkb > 1:
                b.out, b.ldo = parts[j].data_ptr() + q0 * width * 2, width
                b.lse = lses[j, i].data_ptr()
            else:
                b.out = a.out + q0 * a.ldo * 2
                b.lse = a.lse + i * blk * 4
            _lib.check(_lib.lib().vb_attention_fwd(C.byref(b), _stream()), "vb_attention_fwd")
        if nkb > 1:
            _lib.check(_lib.lib().vb_attn_merge(_ptr_array([parts[j].data_ptr() + q0 * width * 2 for j in range(nkb)]),
                                                _ptr_array([lses[j, i].data_ptr() for j in range(nkb)]), nkb, width,
                                                a.out + q0 * a.ldo * 2, a.ldo, a.lse + i * blk * 4, batch, heads, ql, sq, d,
                                                _stream()), "vb_attn_merge")
    return out


def attention_bwd(dout, q, k, v, lse, dq, dk, dv, *, batch, heads, sq, sk, d, mask_bias=None, scale=None, p_drop=0.0,
                  site=0, seed=None, out=None):
    """Backward of attention_fwd.  ``out`` (the forward output) is needed when the keys span more than one block."""
    scale = (1.0 / d ** 0.5) if scale is None else scale
    a = _attn_args(q, k, v, lse, mask_bias, batch, heads, sq, sk, d, scale, p_drop, site, seed)
    _need_cuda(dout, dq, dk, dv, out)
    for t in (dout, dq, dk, dv):
        assert t.dtype == torch.bfloat16 and t.stride(1) == 1
    a.dout, a.lddo = dout.data_ptr(), dout.stride(0)
    a.dq, a.dk, a.dv = dq.data_ptr(), dk.data_ptr(), dv.data_ptr()
    a.lddq, a.lddk, a.lddv = dq.stride(0), dk.stride(0), dv.stride(0)
    if sq <= ATTN_BLOCK and sk <= ATTN_BLOCK:
        _lib.check(_lib.lib().vb_attention_bwd(C.byref(a), _stream()), "vb_attention_bwd")
        return
    qb, kb = attn_blocks(sq), attn_blocks(sk)
    nqb, nkb, width, blk = len(qb), len(kb), heads * d, batch * heads * 128
    shapes = []
    if nkb > 1:
        if out is None:
            raise _lib.VbError("attention_bwd over more than 128 keys needs the forward output (out=...)")
        assert out.dtype == torch.bfloat16 and out.stride(1) == 1 and out.shape == q.shape
        shapes += [((nkb, batch * sq, width), torch.bfloat16), ((nqb, blk), torch.float32)]      # dq partials, delta
    if nqb > 1:
        shapes += [((nqb, batch * sk, width), torch.bfloat16), ((nqb, batch * sk, width), torch.bfloat16)]   # dk, dv partials
    bufs = _scratch("bwd", dq, shapes)
    dq_parts, delta = (bufs[0], bufs[1]) if nkb > 1 else (None, None)
    dk_parts, dv_parts = (bufs[-2], bufs[-1]) if nqb > 1 else (None, None)
    b = _lib.AttnArgs()
    for i, (q0, ql) in enumerate(qb):
        if nkb > 1:
            _lib.check(_lib.lib().vb_attn_delta(out.data_ptr() + q0 * out.stride(0) * 2, out.stride(0),
                                                a.dout + q0 * a.lddo * 2, a.lddo, delta[i].data_ptr(), batch, heads, ql, sq, d,
                                                _stream()), "vb_attn_delta")
        for j, (k0, kl) in enumerate(kb):
            _block_view(b, a, i, j, q0, ql, k0, kl, sq, sk, nkb)
            b.lse = a.lse + i * blk * 4
            b.dout = a.dout + q0 * a.lddo * 2
            if nkb > 1:
                b.delta = delta[i].data_ptr()
                b.dq, b.lddq = dq_parts[j].data_ptr() + q0 * width * 2, width
            else:
                b.dq = a.dq + q0 * a.lddq * 2
            if nqb > 1:
                b.dk, b.lddk = dk_parts[i].data_ptr() + k0 * width * 2, width
                b.dv, b.lddv = dv_parts[i].data_ptr() + k0 * width * 2, width
            else:
                b.dk, b.dv = a.dk + k0 * a.lddk * 2, a.dv + k0 * a.lddv * 2
            _lib.check(_lib.lib().vb_attention_bwd(C.byref(b), _stream()), "vb_attention_bwd")
    if nkb > 1:
        _lib.check(_lib.lib().vb_sum_rows_bf16(_ptr_array([dq_parts[j].data_ptr() for j in range(nkb)]), nkb, width, a.dq,
                                               a.lddq, batch * sq, width, _stream()), "vb_sum_rows_bf16")
    if nqb > 1:
        for parts, dst, ld in ((dk_parts, a.dk, a.lddk), (dv_parts, a.dv, a.lddv)):
            _lib.check(_lib.lib().vb_sum_rows_bf16(_ptr_array([parts[i].data_ptr() for i in range(nqb)]), nqb, width, dst, ld,
                                                   batch * sk, width, _stream()), "vb_sum_rows_bf16")


# ------------------------------------------------------------------------------------------------ RoI feature stage
def stem_im2col(img, out, kh=7, kw=7, stride=2, pad=3):
    """conv1 operand of torchvision resnet152 (reference resnet152_roi.py:49): fp32 NCHW image -> bf16 [pixels, kpad]."""
    _need_cuda(img, out)
    n, c, h, w = img.shape
    assert c == 3 and img.dtype == torch.float32 and img.is_contiguous() and out.dtype == torch.bfloat16 and out.is_contiguous()
    _lib.check(_lib.lib().vb_stem_im2col(img.data_ptr(), out.data_ptr(), n, h, w, kh, kw, stride, pad, out.shape[1], _stream()),
               "vb_stem_im2col")
    return out


def im2col_nhwc(x, out, kh, kw, stride, pad):
    """bf16 NHWC [n,h,w,c] -> bf16 [n*ho*wo, kh*kw*c] (3x3 and strided 1x1 convolutions of the bottlenecks)."""
    _need_cuda(x, out)
    n, h, w, c = x.shape
    assert x.dtype == torch.bfloat16 and x.is_contiguous() and out.is_contiguous()
    _lib.check(_lib.lib().vb_im2col_nhwc(x.data_ptr(), out.data_ptr(), n, h, w, c, kh, kw, stride, pad, _stream()), "vb_im2col_nhwc")
    return out


def maxpool_nhwc(x, out, k=3, stride=2, pad=1):
    _need_cuda(x, out)
    n, h, w, c = x.shape
    _lib.check(_lib.lib().vb_maxpool_nhwc(x.data_ptr(), out.data_ptr(), n, h, w, c, k, stride, pad, _stream()), "vb_maxpool_nhwc")
    return out


def roi_pool_nhwc(x, rois, out, spatial_scale, argmax=None):
    """torchvision.ops.RoIPool on an NHWC bf16 map (reference resnet152_roi.py:126, 167-170); rois fp32 [r,5]."""
    _need_cuda(x, rois, out, argmax)
    n, h, w, c = x.shape
    r, ph, pw, c2 = out.shape
    assert c2 == c and rois.dtype == torch.float32 and rois.shape == (r, 5) and rois.is_contiguous()
    _lib.check(_lib.lib().vb_roi_pool_nhwc(x.data_ptr(), rois.data_ptr(), out.data_ptr(), _ptr(argmax), r, n, h, w, c, ph, pw,
                                           float(spatial_scale), _stream()), "vb_roi_pool_nhwc")
    return out


def roi_align_nhwc(x, rois, out, spatial_scale, sampling_ratio=2, aligned=False):
    _need_cuda(x, rois, out)
    n, h, w, c = x.shape
    r, ph, pw, c2 = out.shape
    assert c2 == c and rois.dtype == torch.float32 and rois.shape == (r, 5) and rois.is_contiguous()
    _lib.check(_lib.lib().vb_roi_align_nhwc(x.data_ptr(), rois.data_ptr(), out.data_ptr(), r, n, h, w, c, ph, pw,
                                            float(spatial_scale), int(sampling_ratio), int(aligned), _stream()), "vb_roi_align_nhwc")
    return out


def avgpool_nhwc(x, out):
    """AdaptiveAvgPool2d((1,1)) + flatten: bf16 [r,s,c] -> fp32 [r,c]."""
    _need_cuda(x, out)
    r, s, c = x.shape
    assert out.dtype == torch.float32 and out.shape == (r, c)
    _lib.check(_lib.lib().vb_avgpool_nhwc(x.data_ptr(), out.data_ptr(), r, s, c, _stream()), "vb_avgpool_nhwc")
    return out


def box_area_score(boxes, img_w, img_h, scores, target=0.15):
    _need_cuda(boxes, scores)
    assert boxes.dtype == torch.float32 and boxes.is_contiguous() and boxes.shape[1] == 4
    _lib.check(_lib.lib().vb_box_area_score(boxes.data_ptr(), boxes.shape[0], float(img_w), float(img_h), float(target),
                                            scores.data_ptr(), _stream()), "vb_box_area_score")
    return scores


def nms(boxes, scores, iou_threshold):
    """torchvision.ops.nms with the CPU kernel's tie order; returns the kept indices (int64, score order)."""
    _need_cuda(boxes, scores)
    n = boxes.shape[0]
    assert boxes.dtype == torch.float32 and boxes.is_contiguous() and scores.dtype == torch.float32
    if n == 0:
        return torch.empty(0, dtype=torch.long, device=boxes.device)
    ws = torch.empty(2 * max(n, 1), dtype=torch.int32, device=boxes.device)
    keep = torch.empty(max(n, 1), dtype=torch.int32, device=boxes.device)
    cnt = torch.zeros(1, dtype=torch.int32, device=boxes.device)
    _lib.check(_lib.lib().vb_nms(boxes.data_ptr(), scores.data_ptr(), n, float(iou_threshold), ws.data_ptr(), keep.data_ptr(),
                                 cnt.data_ptr(), _stream()), "vb_nms")
    return keep[:int(cnt.item())].long()


def nms_device(boxes, scores, iou_threshold, workspace, keep, num_keep):
    """vb_nms without the host read of the survivor count (graph-capturable): keep int32 [n] / num_keep int32 [1] stay on the
    device for ``select_regions``."""
    _need_cuda(boxes, scores, workspace, keep, num_keep)
    n = boxes.shape[0]
    assert boxes.dtype == torch.float32 and boxes.is_contiguous() and scores.dtype == torch.float32 and scores.is_contiguous()
    assert workspace.dtype == torch.int32 and workspace.numel() >= 2 * n and keep.dtype == torch.int32 and keep.numel() >= n
    assert num_keep.dtype == torch.int32 and scores.numel() == n
    _lib.check(_lib.lib().vb_nms(boxes.data_ptr(), scores.data_ptr(), n, float(iou_threshold), workspace.data_ptr(),
                                 keep.data_ptr(), num_keep.data_ptr(), _stream()), "vb_nms")
    return keep, num_keep


def rowmax(x, out, col_begin=0, col_end=None):
    """out[r] = max(x[r, col_begin:col_end]) over fp32 rows (fasterrcnn_vg.py:360-363)."""
    _need_cuda(x, out)
    assert x.dtype == torch.float32 and out.dtype == torch.float32 and x.stride(1) == 1 and out.numel() == x.shape[0]
    col_end = x.shape[1] if col_end is None else col_end
    _lib.check(_lib.lib().vb_rowmax_f32(x.data_ptr(), x.shape[0], x.stride(0), col_begin, col_end, out.data_ptr(), _stream()),
               "vb_rowmax_f32")
    return out


def select_regions(candidates, keep, num_keep, regions, img_w, img_h, *, boxes=None, spatial=None, index=None, feat_src=None,
                   feat_dst=None, rois=None, batch_index=0, box_div=1.0):
    """Region r <- candidate keep[min(r, num_keep - 1)]: box, normalised spatial row, feature row (fasterrcnn_vg.py:367-469)."""
    _need_cuda(candidates, keep, num_keep, boxes, spatial, index, feat_src, feat_dst, rois)
    assert candidates.dtype == torch.float32 and candidates.is_contiguous() and keep.dtype == torch.int32
    for t, width in ((boxes, 4), (spatial, 5), (rois, 5)):
        assert t is None or (t.dtype == torch.float32 and t.is_contiguous() and tuple(t.shape) == (regions, width))
    assert index is None or (index.dtype == torch.int32 and index.numel() == regions)
    dim = 0
    if feat_src is not None:
        dim = feat_src.shape[1]
        assert feat_src.dtype == torch.float32 and feat_dst.dtype == torch.float32 and feat_src.is_contiguous()
        assert feat_dst.is_contiguous() and tuple(feat_dst.shape) == (regions, dim)
    _lib.check(_lib.lib().vb_select_regions(candidates.data_ptr(), keep.data_ptr(), num_keep.data_ptr(), regions, float(img_w),
                                            float(img_h), _ptr(feat_src), dim, _ptr(boxes), _ptr(spatial), _ptr(feat_dst),
                                            _ptr(index), _ptr(rois), float(batch_index), float(box_div), _stream()),
               "vb_select_regions")


def rpn_decode(heads, fh, fw, base_anchors, stride, img_h, img_w, min_size, boxes, scores, num_valid):
    """RPN head outputs -> clipped proposals + foreground scores (-inf below min_size) + their count (fasterrcnn_vg_rpn.py:78-174,
    444-450).  heads fp32 [fh*fw, >= 6A]; base_anchors: HOST float32 array [A, 4]."""
    _need_cuda(heads, boxes, scores, num_valid)
    a = base_anchors.shape[0]
    assert heads.dtype == torch.float32 and heads.stride(1) == 1 and heads.shape[0] == fh * fw
    assert boxes.dtype == torch.float32 and boxes.is_contiguous() and tuple(boxes.shape) == (fh * fw * a, 4)
    assert scores.dtype == torch.float32 and scores.numel() == fh * fw * a and num_valid.dtype == torch.int32
    host = np.ascontiguousarray(base_anchors, dtype=np.float32)
    _lib.check(_lib.lib().vb_rpn_decode(heads.data_ptr(), heads.stride(0), fh, fw, a, host.ctypes.data, float(stride), float(img_h),
                                        float(img_w), float(min_size), boxes.data_ptr(), scores.data_ptr(), num_valid.data_ptr(),
                                        _stream()), "vb_rpn_decode")


def rank_sort_desc(scores, order, limit=None):
    """order[rank] = index, stable descending; elements at or beyond *limit (device int32) count as -inf."""
    _need_cuda(scores, order, limit)
    assert scores.dtype == torch.float32 and scores.is_contiguous() and order.dtype == torch.int32 and order.numel() >= scores.numel()
    _lib.check(_lib.lib().vb_rank_sort_desc(scores.data_ptr(), scores.numel(), _ptr(limit), order.data_ptr(), _stream()),
               "vb_rank_sort_desc")
    return order


def gather_sorted(boxes, scores, order, num_valid, out_boxes, out_scores, count):
    """The first min(cap, *num_valid) boxes / scores in `order` (cap = rows of out_boxes)."""
    _need_cuda(boxes, scores, order, num_valid, out_boxes, out_scores, count)
    cap = out_boxes.shape[0]
    assert out_boxes.dtype == torch.float32 and out_boxes.is_contiguous() and out_scores.numel() == cap and count.dtype == torch.int32
    _lib.check(_lib.lib().vb_gather_sorted(boxes.data_ptr(), scores.data_ptr(), order.data_ptr(), num_valid.data_ptr(), cap,
                                           out_boxes.data_ptr(), out_scores.data_ptr(), count.data_ptr(), _stream()), "vb_gather_sorted")


def nms_sorted(boxes, count, iou_threshold, keep, num_keep):
    """torchvision.ops.nms over boxes already in descending score order, first keep.numel() survivors (device counts)."""
    _need_cuda(boxes, count, keep, num_keep)
    assert boxes.dtype == torch.float32 and boxes.is_contiguous() and boxes.shape[0] <= 8192 and keep.dtype == torch.int32
    _lib.check(_lib.lib().vb_nms_sorted(boxes.data_ptr(), count.data_ptr(), float(iou_threshold), keep.numel(), keep.data_ptr(),
                                        num_keep.data_ptr(), _stream()), "vb_nms_sorted")


def lmdb_regions(features=None, features_bf16=None, boxes=None, spatial=None, box_div=1000.0, area_div=1000000.0, stream=None):
    """Raw LMDB batch -> encoder inputs in one launch: fp32 features -> bf16; raw boxes [rows, >=4] -> spatial [rows, 5]
    (lmdb_dataset.py:189-208, bit-exact).  Either half may be omitted."""
    n = rows = stride = 0
    if features is not None:
        _need_cuda(features, features_bf16)
        assert features.dtype == torch.float32 and features_bf16.dtype == torch.bfloat16
        assert features.is_contiguous() and features_bf16.is_contiguous() and features.numel() == features_bf16.numel()
        n = features.numel()
    if boxes is not None:
        _need_cuda(boxes, spatial)
        assert boxes.dtype == torch.float32 and spatial.dtype == torch.float32 and boxes.is_contiguous() and spatial.is_contiguous()
        stride = boxes.shape[-1]
        rows = boxes.numel() // stride
        assert spatial.numel() == rows * 5
    _lib.check(_lib.lib().vb_lmdb_regions(_ptr(features), _ptr(features_bf16), n, _ptr(boxes), _ptr(spatial), rows, stride,
                                          float(box_div), float(area_div),
                                          _stream() if stream is None else stream.cuda_stream), "vb_lmdb_regions")
    return features_bf16, spatial
