"""Fused gradient clipping + AdamW + bf16 weight-shadow refresh for ``ViLBERTForClassification``.

Drop-in for what the reference's training loop does per step (pipelines/model_training/nodes.py:757-760, 795-799):

    torch.nn.utils.clip_grad_norm_(model.parameters(), 1.0)
    optimizer.step()                      # torch.optim.AdamW(lr=1e-5, weight_decay=0.01)

as two kernels over the model's flat fp32 buffers (``vb_grad_sumsq`` + ``vb_adamw_step``) instead of several hundred
foreach launches over 523 tensors.  It is a ``torch.optim.Optimizer``, so the reference's ``LambdaLR`` warm-up schedule
drives ``param_groups[0]["lr"]`` unchanged.  Parameters without a gradient (frozen layers, the unused ``q_dense*``) are
skipped exactly as torch skips ``grad is None``.  CUDA only; there is no fallback."""
from __future__ import annotations

import ctypes as C
from typing import List, Optional, Tuple

import torch

from . import _lib
from ._lib import VbError


class FusedAdamW(torch.optim.Optimizer):
    def __init__(self, model, lr: float = 1e-3, betas: Tuple[float, float] = (0.9, 0.999), eps: float = 1e-8,
                 weight_decay: float = 1e-2, max_grad_norm: Optional[float] = None):
        self.model = model
        self.max_grad_norm = max_grad_norm
        super().__init__(list(model.parameters()), dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))
        self._flat = None
        self._step = 0
        self.last_grad_norm: Optional[torch.Tensor] = None

    def _bind(self):
        eng = self.model._engine
        if eng is None:
            raise VbError("FusedAdamW.step() before the first forward/backward: the model's flat buffers do not exist yet")
        flat = eng.flat
        if self._flat is not flat:
            if self._flat is not None:
                raise VbError("the model was re-flattened (moved / re-created) after optimizer state was built")
            self._flat = flat
            self.exp_avg = torch.zeros(flat.s_end, dtype=torch.float32, device=flat.device)
            self.exp_avg_sq = torch.zeros(flat.s_end, dtype=torch.float32, device=flat.device)
            self._sumsq = torch.zeros(1, dtype=torch.float64, device=flat.device)
        return flat

    def _ranges(self, flat) -> List[Tuple[int, int]]:
        """Contiguous [lo, hi) ranges of the flat buffer holding parameters that received a gradient (merged, 4-aligned)."""
        spans = []
        for k, p in flat.named.items():
            if k in flat.used and p.requires_grad and p.grad is not None:
                o = flat.offsets[k]
                spans.append((o, o + (p.numel() + 63) // 64 * 64 if o >= flat.w_end else o + p.numel()))
        spans.sort()
        out: List[List[int]] = []
        for lo, hi in spans:
            if out and lo <= out[-1][1]:
                out[-1][1] = max(out[-1][1], hi)
            else:
                out.append([lo, hi])
        return [(lo, min(hi, flat.s_end)) for lo, hi in out]

    @torch.no_grad()
    def step(self, closure=None):
        loss = closure() if closure is not None else None
        flat = self._bind()
        g = self.param_groups[0]
        self._step += 1
        stream = torch.cuda.current_stream(flat.device).cuda_stream
        lib = _lib.lib()
        ranges = self._ranges(flat)
        clip = self.max_grad_norm is not None and self.max_grad_norm > 0
        if clip:
            self._sumsq.zero_()
            for lo, hi in ranges:
                _lib.check(lib.vb_grad_sumsq(flat.grad[lo:hi].data_ptr(), hi - lo, self._sumsq.data_ptr(), stream), "vb_grad_sumsq")
            self.last_grad_norm = self._sumsq
        for lo, hi in ranges:
            a = _lib.AdamWArgs()
            a.param, a.grad = flat.master[lo:hi].data_ptr(), flat.grad[lo:hi].data_ptr()
            a.exp_avg, a.exp_avg_sq = self.exp_avg[lo:hi].data_ptr(), self.exp_avg_sq[lo:hi].data_ptr()
            sn = max(0, min(hi, flat.w_end) - lo)
            a.shadow = flat.shadow[lo:lo + sn].data_ptr() if sn > 0 else None
            a.n, a.shadow_n = hi - lo, sn
            a.grad_sumsq = self._sumsq.data_ptr() if clip else None
            a.max_norm = float(self.max_grad_norm) if clip else 0.0
            a.lr, a.beta1, a.beta2, a.eps, a.weight_decay = g["lr"], g["betas"][0], g["betas"][1], g["eps"], g["weight_decay"]
            a.step = self._step
            _lib.check(lib.vb_adamw_step(C.byref(a), stream), "vb_adamw_step")
        # the master buffer changed under the parameters' feet (raw pointers): shadows are already current, so tell the
        # engine not to recast them on the next forward
        flat._version = flat.versions()
        return loss

    def grad_norm(self) -> float:
        """Total gradient norm seen by the last clipped step (what ``clip_grad_norm_`` returns)."""
        return float(self._sumsq.sqrt().item()) if self.last_grad_norm is not None else float("nan")
