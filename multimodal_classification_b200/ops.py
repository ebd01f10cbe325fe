"""Thin tensor-level wrappers over the C ABI (one Python function per exported kernel).

Tensors are only used for their device pointers, shapes and the current CUDA stream; no torch
compute op is issued here.  Every wrapper raises on a non-zero status.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib
from ._lib import ACT_GELU, ACT_NONE, ACT_RELU, ACT_TANH, AUX_ADD, AUX_MUL_GELU_GRAD, AUX_NONE, GemmArgs  # noqa: F401


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
         block_n=0, splits=0, max_ctas=0) -> torch.Tensor:
    """out[M,N] = epilogue(sum_k A(m,k) B(n,k)).

    a: [M,K] (K-major) or [K,M] (a_mn_major); b: [N,K] or [K,N] (b_mn_major); bf16, last dim contiguous.
    out: bf16 or fp32 [M,N].  Mirrors nn.Linear forward / dgrad / wgrad
    (reference models/vilbert_facebook_arch.py:127-129 and autograd thereof).
    """
    _need_cuda(a, b, out, bias, scale, aux, preact)
    assert a.dtype == torch.bfloat16 and b.dtype == torch.bfloat16
    assert a.stride(-1) == 1 and b.stride(-1) == 1 and out.stride(-1) == 1
    if a_mn_major:
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
    args.lda, args.ldb, args.ldd = a.stride(0), b.stride(0), out.stride(0)
    args.ld_preact = preact.stride(0) if preact is not None else 0
    args.ld_aux = aux.stride(0) if aux is not None else 0
    args.m, args.n, args.k = m, n, k
    args.a_mn_major, args.b_mn_major = int(a_mn_major), int(b_mn_major)
    args.d_is_f32 = int(out.dtype == torch.float32)
    assert out.dtype in (torch.float32, torch.bfloat16)
    args.accumulate = int(accumulate)
    args.act, args.aux_mode = act, aux_mode
    args.block_n, args.splits, args.max_ctas = block_n, splits, max_ctas
    _lib.check(_lib.lib().vb_gemm_bf16(C.byref(args), _stream()), "vb_gemm_bf16")
    return out
