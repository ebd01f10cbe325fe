"""Grid feature extractor on the RoI stage's kernels: drop-in for the reference's
``ResNetFeatureExtractor`` (/root/reference/src/multimodalclassification/models/feature_extractors/resnet.py:17-85) — the
whole frozen ResNet-152 trunk on a 224 x 224 picture, adaptive average pooling of the 7 x 7 x 2048 map to a
sqrt(num_regions) grid, grid boxes as spatial locations.

Same constructor, ``backbone`` state_dict keys ("0." conv1, "1." bn1, "4.".."7." layer1..4), ``extract_features`` and
``forward`` as the reference.  The arithmetic is ``resnet152_roi._Trunk`` (NHWC bf16, every convolution = [im2col +]
``vb_gemm_bf16`` with folded BatchNorm / residual / ReLU epilogue); the adaptive pool is the im2col gather of each grid
cell's window followed by ``vb_avgpool_nhwc`` (fp32 mean).  CUDA only, no fall-back.
"""
from __future__ import annotations

from types import SimpleNamespace
from typing import Dict, Optional, Tuple

import torch
import torch.nn as nn

from . import ops
from ._lib import VbError
from .dinov2_fusion import grid_spatial
from .resnet152_roi import _Trunk


def adaptive_windows(size: int, grid: int) -> Tuple[int, int]:
    """(kernel, stride) such that ``adaptive_avg_pool`` of ``size`` cells to ``grid`` cells averages the windows
    [i*stride, i*stride + kernel); raises when the windows of this pair are not uniform (torch: start = floor(i*size/grid),
    end = ceil((i+1)*size/grid))."""
    starts = [(i * size) // grid for i in range(grid)]
    ends = [-((-(i + 1) * size) // grid) for i in range(grid)]
    kernel = ends[0] - starts[0]
    stride = starts[1] - starts[0] if grid > 1 else 1
    if any(e - s != kernel for s, e in zip(starts, ends)) or any(s != i * stride for i, s in enumerate(starts)):
        raise VbError(f"adaptive pooling of a {size}-wide map to {grid} cells has non-uniform windows; supported grids for the "
                      "7 x 7 map of a 224 x 224 picture are 1, 2, 3, 6 and 7 (num_regions 1, 4, 9, 36, 49)")
    return kernel, max(stride, 1)


class ResNetFeatureExtractor(nn.Module):
    """Reference ``ResNetFeatureExtractor`` (resnet.py:17-85).  Extra keyword-only arguments: ``weights`` (torchvision weight
    name, or None for random init; the reference hard-codes IMAGENET1K_V2) and ``image_size`` (the reference resizes to 224)."""

    def __init__(self, output_dim: int = 2048, num_regions: int = 36, device: Optional[str] = None, *,
                 weights: Optional[str] = "IMAGENET1K_V2", image_size: int = 224):
        super().__init__()
        device = "cuda" if device is None else device
        if not str(device).startswith("cuda"):
            raise VbError("ResNetFeatureExtractor (B200) runs on CUDA only; there is no CPU fallback")
        from torchvision import transforms
        self.output_dim, self.num_regions, self.device, self.image_size = output_dim, num_regions, device, image_size
        self.backbone = self._make_backbone(weights)                           # parameter container only
        self.backbone.eval().to(device)
        for p in self.backbone.parameters():
            p.requires_grad = False
        self.transform = transforms.Compose([
            transforms.Resize((image_size, image_size)), transforms.ToTensor(),
            transforms.Normalize(mean=[0.485, 0.456, 0.406], std=[0.229, 0.224, 0.225])])
        self._to_pil = transforms.ToPILImage()
        self._trunk: Optional[_Trunk] = None
        self._plans: Dict[Tuple[int, int, int], dict] = {}
        self.use_graphs = True

    def _make_backbone(self, weights: Optional[str]) -> nn.Module:
        from torchvision.models import ResNet152_Weights, resnet152
        resnet = resnet152(weights=None if weights is None else getattr(ResNet152_Weights, weights))
        return nn.Sequential(*list(resnet.children())[:-2])                    # resnet.py:33

    def _generate_grid_spatial(self, num_regions: Optional[int] = None) -> torch.Tensor:
        """models/base.py:244-270."""
        n = self.num_regions if num_regions is None else num_regions
        g = int(n ** 0.5)
        return grid_spatial(n)[: g * g].to(self.device)

    def _trunk_parts(self) -> Tuple[nn.Sequential, nn.Sequential]:
        """(conv1 .. layer3 as modules 0..6, layer4) of the parameter container."""
        return self.backbone, self.backbone[7]

    def _get_trunk(self) -> _Trunk:
        base, top = self._trunk_parts()
        ver = sum(p._version for p in self.backbone.parameters()) + sum(b._version for b in self.backbone.buffers())
        if self._trunk is None or self._trunk.version != ver or self._trunk.device != base[0].weight.device:
            self._trunk = _Trunk(SimpleNamespace(base=base, top=top), ver)
        return self._trunk

    def _run(self, plan: dict) -> None:
        t: _Trunk = plan["trunk"]
        fmap = t.run_layer(t.base(plan["img"]), t.layer4, "l4")               # [B, h, w, 2048] bf16
        b, h, w, ch = fmap.shape
        g = int(self.num_regions ** 0.5)
        (kh, sh), (kw, sw) = adaptive_windows(h, g), adaptive_windows(w, g)
        if kh != kw or sh != sw:
            raise VbError("non-square feature maps are not supported by the grid extractor")
        if kh == 1:
            cells = fmap.view(b * g * g, 1, ch)
        else:
            col = t.buf("cells", (b * g * g, kh * kw * ch))
            ops.im2col_nhwc(fmap, col, kh, kw, sh, 0)
            cells = col.view(b * g * g, kh * kw, ch)
        ops.avgpool_nhwc(cells, plan["feats"])

    def _plan(self, b: int, h: int, w: int) -> dict:
        trunk = self._get_trunk()
        key = (b, h, w)
        plan = self._plans.get(key)
        if plan is None or plan["trunk"] is not trunk:
            g = int(self.num_regions ** 0.5)
            plan = self._plans[key] = {"trunk": trunk, "img": torch.zeros(b, 3, h, w, device=trunk.device),
                                       "feats": torch.zeros(b * g * g, 2048, device=trunk.device), "graph": None, "gen": -1}
        return plan

    @torch.no_grad()
    def extract_batch(self, imgs: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        """Preprocessed fp32 NCHW images on the GPU -> ([B, g*g, output_dim] fp32, [B, g*g, 5] fp32)."""
        if not imgs.is_cuda:
            raise VbError("extract_batch needs CUDA tensors; there is no CPU fallback")
        b, _, h, w = imgs.shape
        with torch.cuda.device(imgs.device):
            plan = self._plan(b, h, w)
            trunk: _Trunk = plan["trunk"]
            plan["img"].copy_(imgs)
            if self.use_graphs and plan["graph"] is not None and plan["gen"] == trunk.arena_gen:
                plan["graph"].replay()
            else:
                self._run(plan)                       # eager: sizes the scratch arena and produces this result
                if self.use_graphs:
                    torch.cuda.synchronize()
                    graph = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(graph):
                        self._run(plan)
                    plan["graph"], plan["gen"] = graph, trunk.arena_gen
            feats = plan["feats"].view(b, -1, 2048)
            if self.output_dim > 2048:                                          # resnet.py:65-73
                feats = torch.cat([feats, feats.new_zeros(b, feats.shape[1], self.output_dim - 2048)], dim=-1)
            else:
                feats = feats[..., : self.output_dim].clone()
            spatial = self._generate_grid_spatial()
            return feats, spatial.unsqueeze(0).expand(b, *spatial.shape).clone()

    @torch.no_grad()
    def extract_features(self, image) -> Tuple[torch.Tensor, torch.Tensor]:
        """resnet.py:51-76: PIL image -> ([g*g, output_dim], [num_regions grid boxes, 5])."""
        feats, spatial = self.extract_batch(self.transform(image).unsqueeze(0).to(self.device))
        return feats[0], spatial[0]

    @torch.no_grad()
    def forward(self, images: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        """resnet.py:78-85: the same per-image host preprocessing via PIL, then ONE batched pass through the trunk."""
        batch = torch.stack([self.transform(self._to_pil(img.cpu())) for img in images]).to(self.device)
        return self.extract_batch(batch)


# ---------------------------------------------------------------------------------------------------------------------
# the Visual Genome ResNet-101 variant (reference models/feature_extractors/resnet_vg.py)
# ---------------------------------------------------------------------------------------------------------------------
class VGResNet101Backbone(nn.Module):
    """Parameter container with the reference's layout (resnet_vg.py:29-54): ``RCNN_base`` = conv1 .. layer3, ``RCNN_top`` =
    layer4 of a torchvision ResNet-101 (the Visual Genome Faster R-CNN checkpoint's names)."""

    def __init__(self, weights: Optional[str] = "IMAGENET1K_V1"):
        super().__init__()
        from torchvision.models import ResNet101_Weights, resnet101
        resnet = resnet101(weights=None if weights is None else getattr(ResNet101_Weights, weights))
        self.RCNN_base = nn.Sequential(resnet.conv1, resnet.bn1, resnet.relu, resnet.maxpool, resnet.layer1, resnet.layer2,
                                       resnet.layer3)
        self.RCNN_top = resnet.layer4


def load_vg_backbone_weights(model: VGResNet101Backbone, checkpoint_path: str) -> dict:
    """Same contract as the reference loader (resnet_vg.py:70-120): take ``RCNN_base.*`` / ``RCNN_top.*`` tensors of the
    checkpoint (``RCNN_top.0.X`` is the model's ``RCNN_top.X``) whose shapes match, ignore everything else, report counts."""
    checkpoint = torch.load(checkpoint_path, map_location="cpu")
    state = checkpoint.get("model", checkpoint)
    own = model.state_dict()
    loaded, skipped = {}, {}
    for key, value in state.items():
        if not (key.startswith("RCNN_base") or key.startswith("RCNN_top")):
            skipped[key] = "not backbone"
            continue
        name = "RCNN_top." + key[len("RCNN_top.0."):] if key.startswith("RCNN_top.0.") else key
        if name not in own:
            skipped[key] = "key not in model"
        elif own[name].shape != value.shape:
            skipped[key] = f"shape mismatch: {own[name].shape} vs {value.shape}"
        else:
            loaded[name] = value
    model.load_state_dict(loaded, strict=False)
    return {"loaded": len(loaded), "total_model": len(own), "skipped": len(skipped), "skipped_keys": list(skipped.keys())[:10]}


class ResNetVGExtractor(ResNetFeatureExtractor):
    """Reference ``ResNetVGExtractor`` (resnet_vg.py:123-259): the same grid extraction over the VG-pretrained ResNet-101;
    ``weights_path`` names the Visual Genome checkpoint, the torchvision weights are kept when it does not exist."""

    def __init__(self, output_dim: int = 2048, num_regions: int = 36, weights_path: Optional[str] = None,
                 device: Optional[str] = None, *, weights: Optional[str] = "IMAGENET1K_V1", image_size: int = 224):
        import os
        super().__init__(output_dim, num_regions, device, weights=weights, image_size=image_size)
        weights_path = "weights/faster_rcnn_res101_vg.pth" if weights_path is None else weights_path
        self.has_vg_weights = False
        if os.path.exists(weights_path):
            self.has_vg_weights = load_vg_backbone_weights(self.backbone, weights_path)["loaded"] > 0
        self.grid_size = int(num_regions ** 0.5)

    def _make_backbone(self, weights: Optional[str]) -> nn.Module:
        return VGResNet101Backbone(weights)

    def _trunk_parts(self) -> Tuple[nn.Sequential, nn.Sequential]:
        return self.backbone.RCNN_base, self.backbone.RCNN_top
