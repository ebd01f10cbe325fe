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
        self._flat = None
        self._step = 0
        self._pending_state = None          # moments loaded before the model's flat buffers exist
        self.last_grad_norm: Optional[torch.Tensor] = None
        super().__init__(list(model.parameters()), dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))

    def add_param_group(self, param_group):
        """One group only: the update is two launches over the model's flat buffers with one set of hyper-parameters.  A second
        group (per-group lr / weight decay) would be silently ignored by that kernel, so it is refused."""
        if len(self.param_groups) >= 1:
            raise VbError("FusedAdamW applies ONE set of hyper-parameters to the model's flat parameter buffer; per-group lr / "
                          "weight_decay is not supported (use torch.optim.AdamW for that)")
        super().add_param_group(param_group)

    # the moments are two flat fp32 buffers, not per-parameter entries of self.state: carry them (and the step count, which
    # drives the bias correction) through state_dict() / load_state_dict() explicitly
    def state_dict(self):
        sd = super().state_dict()
        fused = {"step": self._step}
        if self._flat is not None:
            fused["exp_avg"], fused["exp_avg_sq"] = self.exp_avg.clone(), self.exp_avg_sq.clone()
        elif self._pending_state is not None:
            fused.update({k: v for k, v in self._pending_state.items() if k != "step"})
        sd["fused_adamw"] = fused
        return sd

    def load_state_dict(self, state_dict):
        state_dict = dict(state_dict)
        fused = state_dict.pop("fused_adamw", None)
        super().load_state_dict(state_dict)
        if fused is None:
            raise VbError("not a FusedAdamW state_dict (no 'fused_adamw' entry): moments and step count would be lost")
        self._step = int(fused["step"])
        if "exp_avg" in fused:
            if self._flat is not None:
                self.exp_avg.copy_(fused["exp_avg"])
                self.exp_avg_sq.copy_(fused["exp_avg_sq"])
            else:
                self._pending_state = {"exp_avg": fused["exp_avg"], "exp_avg_sq": fused["exp_avg_sq"]}

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
            if self._pending_state is not None:
                if self._pending_state["exp_avg"].numel() != flat.s_end:
                    raise VbError("loaded FusedAdamW moments do not match this model's flat parameter layout")
                self.exp_avg.copy_(self._pending_state["exp_avg"])
                self.exp_avg_sq.copy_(self._pending_state["exp_avg_sq"])
                self._pending_state = None
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
        # The kernel writes the master buffer through raw pointers, so no parameter's _version moves and the engine does not
        # recast the (already refreshed) shadows on the next forward.  flat._version is deliberately left alone: an in-place
        # edit made since the last forward (load_state_dict, a manual change of a frozen tensor) must still trigger a recast.
        return loss

    def grad_norm(self) -> float:
        """Total gradient norm seen by the last clipped step (what ``clip_grad_norm_`` returns)."""
        return float(self._sumsq.sqrt().item()) if self.last_grad_norm is not None else float("nan")
