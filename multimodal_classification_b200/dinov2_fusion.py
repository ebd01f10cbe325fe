"""B200 kernels for the multi-layer fusion tail of the reference's ``DINOv2MultiLayerExtractor``
(models/feature_extractors/dinov2_multilayer.py:342-403, fusion_strategy="concat"): everything after the third-party ViT.

    layer features (L x [B, 1 + g*g, 1024] fp32, CLS first)
      -> concat along features, g x g grid -> bilinear resize to t x t regions           vb_bilinear_concat
      -> Linear(L*1024, 2048)                                                             vb_gemm_bf16
      -> LayerNorm(2048, eps 1e-5) -> GELU                                                 vb_layernorm_fwd, vb_gelu_bf16
      -> Linear(2048, 2048)                                                               vb_gemm_bf16 (fp32 out)
    + the uniform grid boxes of ``_generate_grid_spatial`` (:383-403)

``projection`` keeps the reference's parameter names (``projection.0.weight`` ... ``projection.3.bias``), so its state_dict
loads from / into the reference extractor's.  CUDA only; there is no fallback."""
from __future__ import annotations

import ctypes as C
from typing import Sequence, Tuple

import torch
import torch.nn as nn

from . import _lib, ops
from ._lib import VbError


def grid_spatial(num_regions: int) -> torch.Tensor:
    """Reference ``_generate_grid_spatial`` (:383-403): [num_regions, 5] = (x1, y1, x2, y2, area), Python-float arithmetic."""
    g = int(num_regions ** 0.5)
    out = torch.zeros(num_regions, 5)
    for i in range(g):
        for j in range(g):
            x1, y1, x2, y2 = j / g, i / g, (j + 1) / g, (i + 1) / g
            out[i * g + j] = torch.tensor([x1, y1, x2, y2, (x2 - x1) * (y2 - y1)])
    return out


class DINOv2FusionTail(nn.Module):
    def __init__(self, num_layers: int = 4, hidden_size: int = 1024, output_dim: int = 2048, num_regions: int = 36,
                 device: str = "cuda"):
        super().__init__()
        if not str(device).startswith("cuda"):
            raise VbError("DINOv2FusionTail (B200) runs on CUDA only; there is no CPU fallback")
        self.num_layers, self.hidden_size, self.output_dim, self.num_regions = num_layers, hidden_size, output_dim, num_regions
        self.projection = nn.Sequential(nn.Linear(num_layers * hidden_size, output_dim), nn.LayerNorm(output_dim), nn.GELU(),
                                        nn.Linear(output_dim, output_dim))
        for m in self.projection.modules():          # reference _init_projection_weights (:260-266)
            if isinstance(m, nn.Linear):
                nn.init.xavier_uniform_(m.weight)
                nn.init.zeros_(m.bias)
        self.to(device)
        self._shadow = None

    def _weights(self):
        ver = sum(p._version for p in self.parameters())
        if self._shadow is None or self._shadow[0] != ver:
            w1 = self.projection[0].weight.detach().to(torch.bfloat16).contiguous()
            w2 = self.projection[3].weight.detach().to(torch.bfloat16).contiguous()
            self._shadow = (ver, w1, w2)
        return self._shadow[1], self._shadow[2]

    @torch.no_grad()
    def fuse(self, layer_features: Sequence[torch.Tensor], has_cls: bool = False) -> Tuple[torch.Tensor, torch.Tensor]:
        """L tensors [B, P(+1), hidden] fp32 on the GPU -> ([B, num_regions, output_dim] fp32, [B, num_regions, 5] fp32)."""
        if len(layer_features) != self.num_layers:
            raise VbError(f"expected {self.num_layers} layer tensors")
        f0 = layer_features[0]
        if not f0.is_cuda:
            raise VbError("DINOv2FusionTail needs CUDA tensors; there is no CPU fallback")
        b, tokens, h = f0.shape
        patches = tokens - (1 if has_cls else 0)
        g, t = int(patches ** 0.5), int(self.num_regions ** 0.5)
        if g * g != patches or h != self.hidden_size:
            raise VbError("layer features must be [B, g*g (+CLS), hidden]")
        feats = [f.float().contiguous() for f in layer_features]
        for f in feats:
            assert f.shape == f0.shape
        with torch.cuda.device(f0.device):
            m, k = b * t * t, self.num_layers * h
            ptrs = (C.c_void_p * len(feats))(*[f.data_ptr() for f in feats])
            x = torch.empty(m, k, dtype=torch.bfloat16, device=f0.device)
            _lib.check(_lib.lib().vb_bilinear_concat(ptrs, len(feats), x.data_ptr(), b, g, t, h, tokens * h, h, 1 if has_cls else 0,
                                                     torch.cuda.current_stream().cuda_stream), "vb_bilinear_concat")
            w1, w2 = self._weights()
            p = self.projection
            y1 = torch.empty(m, self.output_dim, dtype=torch.bfloat16, device=f0.device)
            ops.gemm(x, w1, y1, bias=p[0].bias.detach())
            mean, rstd = torch.empty(m, device=f0.device), torch.empty(m, device=f0.device)
            y2 = torch.empty_like(y1)
            ops.layernorm_fwd(y1, None, p[1].weight.detach(), p[1].bias.detach(), y2, mean, rstd, eps=p[1].eps)
            _lib.check(_lib.lib().vb_gelu_bf16(y2.data_ptr(), y2.data_ptr(), y2.numel(), torch.cuda.current_stream().cuda_stream),
                       "vb_gelu_bf16")
            out = torch.empty(m, self.output_dim, dtype=torch.float32, device=f0.device)
            ops.gemm(y2, w2, out, bias=p[3].bias.detach())
        spatial = grid_spatial(self.num_regions).to(f0.device)
        return out.view(b, t * t, self.output_dim), spatial.unsqueeze(0).expand(b, -1, -1).contiguous()
