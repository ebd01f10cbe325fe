"""B200-native ResNet-152 RoI feature extractor: drop-in for the reference's
``models/feature_extractors/resnet152_roi.py`` (``ResNet152Backbone`` :35-74, ``ResNet152ROIExtractor`` :77-324).

Same constructor, same ``extract_features(PIL.Image) -> ([num_regions, 2048], [num_regions, 5])`` and
``forward(images[B,3,H,W]) -> ([B,N,2048], [B,N,5])``, same ``backbone.base / backbone.top`` parameter layout (torchvision
resnet152 state_dict keys), same proposal boxes (bit-exact) and RoIPool index arithmetic (bit-exact).

Device side, every op is a kernel of ``libvilbert_b200.so``: activations are NHWC bf16, each of the 155 convolutions is the
tcgen05 GEMM over ``[pixels, kh*kw*Cin]`` (1x1: the activation itself; 3x3 / strided: an im2col gather) with the eval-mode
BatchNorm folded into the GEMM epilogue's per-channel scale / bias, the bottleneck identity as the epilogue's residual
operand and the ReLU fused; max-pool, RoIPool / RoIAlign, global average pool and the proposal NMS are bandwidth kernels.
The whole trunk for one batch shape is captured in a CUDA graph.  There is no CPU or PyTorch fallback.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Tuple

import numpy as np
import torch
import torch.nn as nn

from . import ops
from ._lib import VbError

ACT_NONE, ACT_RELU = ops.ACT_NONE, ops.ACT_RELU


# ---------------------------------------------------------------------------------------------------------------------
# proposals (reference :180-311) -- host arithmetic mirrors the reference's Python floats exactly; scoring + NMS on the GPU
# ---------------------------------------------------------------------------------------------------------------------
SCALES = (0.15, 0.25, 0.35, 0.5, 0.7)
ASPECT_RATIOS = (0.5, 0.75, 1.0, 1.33, 2.0)


def grid_boxes(num_regions: int, img_h: int, img_w: int) -> np.ndarray:
    """Reference ``_generate_grid_proposals`` (:191-206): floor(sqrt(N))^2 equal cells, row-major."""
    g = int(num_regions ** 0.5)
    ch, cw = img_h / g, img_w / g
    out = [[j * cw, i * ch, (j + 1) * cw, (i + 1) * ch] for i in range(g) for j in range(g)]
    return np.asarray(out, dtype=np.float32).reshape(-1, 4)


def sliding_window_boxes(img_h: int, img_w: int) -> np.ndarray:
    """Candidates of ``_generate_multi_scale_proposals`` (:208-240): 5 scales x 5 aspect ratios, stride max(0.4*side, 20),
    accumulated in Python doubles exactly as the reference does, then rounded to fp32 once."""
    out: List[List[float]] = []
    for scale in SCALES:
        for ar in ASPECT_RATIOS:
            bw = img_w * scale
            bh = bw / ar
            bh = min(bh, img_h * 0.95)
            bw = min(bw, img_w * 0.95)
            sx, sy = max(bw * 0.4, 20), max(bh * 0.4, 20)
            x = 0
            while x + bw <= img_w:
                y = 0
                while y + bh <= img_h:
                    out.append([x, y, x + bw, y + bh])
                    y += sy
                x += sx
    return np.asarray(out, dtype=np.float32).reshape(-1, 4)


def normalize_boxes(boxes: np.ndarray, img_w: int, img_h: int) -> np.ndarray:
    """Reference ``_normalize_boxes`` (:295-311) in fp32: x/W, y/H, clamp to [0,1], append w*h."""
    b = boxes.astype(np.float32).copy()
    b[:, [0, 2]] /= np.float32(img_w)
    b[:, [1, 3]] /= np.float32(img_h)
    b = np.clip(b, np.float32(0), np.float32(1))
    area = (b[:, 2] - b[:, 0]) * (b[:, 3] - b[:, 1])
    return np.concatenate([b, area[:, None]], axis=1).astype(np.float32)


def select_boxes(cands: torch.Tensor, num_regions: int, img_h: int, img_w: int) -> torch.Tensor:
    """Reference ``_select_diverse_boxes`` (:251-293): area score, NMS(0.5), first N kept (padded from the suppressed ones in
    index order when NMS leaves fewer).  `cands` is a CUDA fp32 [n,4] tensor; returns the chosen rows."""
    scores = torch.empty(cands.shape[0], dtype=torch.float32, device=cands.device)
    ops.box_area_score(cands, img_w, img_h, scores, 0.15)
    keep = ops.nms(cands, scores, 0.5)
    if keep.numel() < num_regions:
        kept = set(keep.tolist())
        rest = [i for i in range(cands.shape[0]) if i not in kept][: num_regions - keep.numel()]
        keep = torch.cat([keep, torch.tensor(rest, dtype=torch.long, device=cands.device)])
    return cands[keep[:num_regions]]


def generate_proposals(num_regions: int, img_h: int, img_w: int, use_multi_scale: bool, device) -> torch.Tensor:
    """Reference ``_generate_proposals`` (:180-249): fp32 [num_regions, 4] on `device`."""
    if not use_multi_scale:
        return torch.from_numpy(grid_boxes(num_regions, img_h, img_w)).to(device)
    cands = torch.from_numpy(sliding_window_boxes(img_h, img_w)).to(device)
    if cands.shape[0] > num_regions:
        cands = select_boxes(cands, num_regions, img_h, img_w)
    elif cands.shape[0] < num_regions:
        cands = torch.cat([cands, torch.from_numpy(grid_boxes(num_regions, img_h, img_w)).to(device)], dim=0)
    return cands[:num_regions].contiguous()


# ---------------------------------------------------------------------------------------------------------------------
# backbone: parameter container with the reference's layout + the kernel engine
# ---------------------------------------------------------------------------------------------------------------------
class _ConvSpec:
    """One convolution + its folded eval-mode BatchNorm, laid out for the GEMM: w bf16 [Cout, kh*kw*Cin] with column
    (ky*kw + kx)*Cin + ci, scale / bias fp32 [Cout]."""

    def __init__(self, conv: nn.Conv2d, bn: nn.BatchNorm2d, kpad: Optional[int] = None):
        w = conv.weight.detach()
        cout, cin, kh, kw = w.shape
        self.cin, self.cout, self.kh, self.kw = cin, cout, kh, kw
        self.stride, self.pad = conv.stride[0], conv.padding[0]
        flat = w.permute(0, 2, 3, 1).reshape(cout, kh * kw * cin)
        k = flat.shape[1] if kpad is None else kpad
        wb = torch.zeros(cout, k, dtype=torch.bfloat16, device=w.device)
        wb[:, : flat.shape[1]] = flat.to(torch.bfloat16)
        self.w = wb
        var = bn.running_var.detach().float()
        scale = bn.weight.detach().float() / torch.sqrt(var + bn.eps)
        self.scale = scale.contiguous()
        self.bias = (bn.bias.detach().float() - bn.running_mean.detach().float() * scale).contiguous()
        if conv.bias is not None:
            self.bias = (self.bias + conv.bias.detach().float() * scale).contiguous()


class ResNet152Backbone(nn.Module):
    """Reference class of the same name (:35-74): ``base`` = conv1..layer3 (stride 16, 1024 channels), ``top`` = layer4,
    global average pool.  The torchvision modules only HOLD the parameters (same state_dict keys as the reference); the
    arithmetic runs in ``_Trunk`` below."""

    def __init__(self, weights: Optional[str] = "IMAGENET1K_V2"):
        super().__init__()
        from torchvision.models import ResNet152_Weights, resnet152
        w = None if weights is None else getattr(ResNet152_Weights, weights)
        resnet = resnet152(weights=w)
        self.base = nn.Sequential(resnet.conv1, resnet.bn1, resnet.relu, resnet.maxpool, resnet.layer1, resnet.layer2,
                                  resnet.layer3)
        self.top = resnet.layer4
        self.avgpool = nn.AdaptiveAvgPool2d((1, 1))
        self._trunk: Optional["_Trunk"] = None

    def _apply(self, fn, *a, **k):
        out = super()._apply(fn, *a, **k)
        self._trunk = None
        return out

    def trunk(self) -> "_Trunk":
        dev = self.base[0].weight.device
        if dev.type != "cuda":
            raise VbError("ResNet152Backbone (B200) runs on CUDA only; there is no CPU fallback")
        ver = sum(p._version for p in self.parameters()) + sum(b._version for b in self.buffers())
        if self._trunk is None or self._trunk.version != ver:
            self._trunk = _Trunk(self, ver)
        return self._trunk

    # boundary adapters with the reference's NCHW fp32 signatures (the extractor itself stays in NHWC bf16 end to end)
    def forward_base(self, x: torch.Tensor) -> torch.Tensor:
        t = self.trunk()
        y = t.base(x.float().contiguous())
        return y.permute(0, 3, 1, 2).float()

    def forward_top(self, x: torch.Tensor) -> torch.Tensor:
        t = self.trunk()
        nhwc = x.permute(0, 2, 3, 1).contiguous().to(torch.bfloat16)
        return t.top(nhwc).clone()


class _Trunk:
    """Prepared weights + scratch arena + kernel schedule for one backbone."""

    def __init__(self, bb: ResNet152Backbone, version: int):
        self.version = version
        self.device = bb.base[0].weight.device
        with torch.no_grad():
            self.stem = _ConvSpec(bb.base[0], bb.base[1], kpad=152)
            self.stages = [self._prep_layer(bb.base[i]) for i in (4, 5, 6)]
            self.layer4 = self._prep_layer(bb.top)
        self.arena: Dict[str, torch.Tensor] = {}
        self.arena_gen = 0
        # 3x3 and strided convolutions as implicit GEMMs (TMA im2col loads).  VB_CONV_EXPLICIT=1: materialise the
        # [pixels, kh*kw*Cin] matrix first (the round-1 path, kept for A/B measurements and as the kernel's own cross-check)
        import os
        self.implicit = os.environ.get("VB_CONV_EXPLICIT", "0") != "1"

    @staticmethod
    def _prep_layer(layer: nn.Sequential):
        blocks = []
        for blk in layer:
            ds = None
            if blk.downsample is not None:
                ds = _ConvSpec(blk.downsample[0], blk.downsample[1])
            blocks.append((_ConvSpec(blk.conv1, blk.bn1), _ConvSpec(blk.conv2, blk.bn2), _ConvSpec(blk.conv3, blk.bn3), ds))
        return blocks

    # ---- scratch arena: named, grow-only, reused across blocks (all launches are ordered on one stream)
    def buf(self, name: str, shape, dtype=torch.bfloat16) -> torch.Tensor:
        n = int(np.prod(shape))
        t = self.arena.get(name)
        if t is None or t.numel() < n or t.dtype != dtype:
            if torch.cuda.is_current_stream_capturing():
                raise VbError("scratch arena grew during graph capture (warm-up pass missing)")
            t = torch.empty(n, dtype=dtype, device=self.device)
            self.arena[name] = t
            self.arena_gen += 1
        return t[:n].view(shape)

    # ---- one convolution (+BN +ReLU +residual) = [im2col +] GEMM
    def conv(self, x: torch.Tensor, c: _ConvSpec, out_name: str, relu: bool, residual: Optional[torch.Tensor] = None):
        b, h, w, cin = x.shape
        assert cin == c.cin, (cin, c.cin)
        ho = (h + 2 * c.pad - c.kh) // c.stride + 1
        wo = (w + 2 * c.pad - c.kw) // c.stride + 1
        conv = None
        if c.kh == 1 and c.kw == 1 and c.stride == 1:
            a = x.view(-1, cin)
        elif self.implicit and cin % 64 == 0:
            a, conv = x, (c.kh, c.kw, c.stride, c.pad)      # implicit GEMM: TMA gathers the filter taps (im2col mode)
        else:
            a = self.buf("col", (b * ho * wo, c.kh * c.kw * cin))
            ops.im2col_nhwc(x, a, c.kh, c.kw, c.stride, c.pad)
        out = self.buf(out_name, (b, ho, wo, c.cout))
        ops.gemm(a, c.w, out.view(-1, c.cout), scale=c.scale, bias=c.bias, act=ACT_RELU if relu else ACT_NONE,
                 aux=None if residual is None else residual.view(-1, c.cout),
                 aux_mode=ops.AUX_NONE if residual is None else ops.AUX_ADD, conv=conv)
        return out

    def bottleneck(self, x: torch.Tensor, blk, out_name: str):
        """torchvision Bottleneck.forward (v1.5: the stride sits on the 3x3): relu(bn3(conv3(..)) + identity)."""
        c1, c2, c3, ds = blk
        t1 = self.conv(x, c1, "t1", relu=True)
        t2 = self.conv(t1, c2, "t2", relu=True)
        identity = x if ds is None else self.conv(x, ds, "ds", relu=False)
        return self.conv(t2, c3, out_name, relu=True, residual=identity)

    def run_layer(self, x: torch.Tensor, blocks, tag: str):
        """Blocks ping-pong between two buffers of their own layer (the layer's input lives in the previous layer's)."""
        names = (tag + ".a", tag + ".b")
        for i, blk in enumerate(blocks):
            x = self.bottleneck(x, blk, names[i & 1])
        return x

    def base(self, img: torch.Tensor) -> torch.Tensor:
        """conv1 .. layer3 (reference ``forward_base`` :65-67): fp32 NCHW [B,3,H,W] -> bf16 NHWC [B,H/16,W/16,1024]."""
        b, _, h, w = img.shape
        c = self.stem
        ho, wo = (h + 2 * c.pad - c.kh) // c.stride + 1, (w + 2 * c.pad - c.kw) // c.stride + 1
        col = self.buf("col", (b * ho * wo, c.w.shape[1]))
        ops.stem_im2col(img, col, c.kh, c.kw, c.stride, c.pad)
        s = self.buf("stem", (b, ho, wo, c.cout))
        ops.gemm(col, c.w, s.view(-1, c.cout), scale=c.scale, bias=c.bias, act=ACT_RELU)
        hp, wp = (ho + 2 - 3) // 2 + 1, (wo + 2 - 3) // 2 + 1
        x = self.buf("pool", (b, hp, wp, c.cout))
        ops.maxpool_nhwc(s, x, 3, 2, 1)
        for i, blocks in enumerate(self.stages):
            x = self.run_layer(x, blocks, f"l{i + 1}")
        return x

    def top(self, pooled: torch.Tensor) -> torch.Tensor:
        """layer4 + global average pool + flatten (reference ``forward_top`` :69-74): bf16 NHWC [R,p,p,1024] -> fp32 [R,2048]."""
        x = self.run_layer(pooled, self.layer4, "l4")
        r, h, w, ch = x.shape
        out = self.buf("feat", (r, ch), torch.float32)
        ops.avgpool_nhwc(x.view(r, h * w, ch), out)
        return out


# ---------------------------------------------------------------------------------------------------------------------
# the extractor
# ---------------------------------------------------------------------------------------------------------------------
class ResNet152ROIExtractor(nn.Module):
    """Reference ``ResNet152ROIExtractor`` (:77-324).  Extra keyword-only arguments: ``weights`` (torchvision weight name or
    None for random init; the reference hard-codes IMAGENET1K_V2), ``image_size`` (the reference resizes to 600) and
    ``pool_mode`` ("roi_pool" as the reference, or "roi_align": torchvision.ops.roi_align, sampling_ratio 2)."""

    def __init__(self, output_dim: int = 2048, num_regions: int = 36, roi_size: int = 14, use_multi_scale: bool = True,
                 device: Optional[str] = None, *, weights: Optional[str] = "IMAGENET1K_V2", image_size: int = 600,
                 pool_mode: str = "roi_pool"):
        super().__init__()
        if device is None:
            device = "cuda"
        if not str(device).startswith("cuda"):
            raise VbError("ResNet152ROIExtractor (B200) runs on CUDA only; there is no CPU fallback")
        if pool_mode not in ("roi_pool", "roi_align"):
            raise VbError("pool_mode must be 'roi_pool' or 'roi_align'")
        self.output_dim, self.num_regions, self.device = output_dim, num_regions, device
        self.roi_size, self.use_multi_scale = roi_size, use_multi_scale
        self.image_size, self.pool_mode = image_size, pool_mode
        self.backbone = ResNet152Backbone(weights=weights)
        self.backbone.to(self.device).eval()
        for p in self.backbone.parameters():
            p.requires_grad = False
        from torchvision import transforms
        self.transform = transforms.Compose([
            transforms.Resize((image_size, image_size)), transforms.ToTensor(),
            transforms.Normalize(mean=[0.485, 0.456, 0.406], std=[0.229, 0.224, 0.225])])
        self._to_pil = transforms.ToPILImage()
        self._boxes: Dict[Tuple[int, int], Tuple[torch.Tensor, torch.Tensor]] = {}
        self._plans: Dict[Tuple[int, int, int], dict] = {}
        self.use_graphs = True

    # -- proposals are a pure function of the image size: computed once (GPU score + NMS), cached
    def _generate_proposals(self, img_h: int, img_w: int) -> torch.Tensor:
        return self._proposals(img_h, img_w)[0]

    def _proposals(self, img_h: int, img_w: int):
        key = (img_h, img_w)
        if key not in self._boxes:
            boxes = generate_proposals(self.num_regions, img_h, img_w, self.use_multi_scale, torch.device(self.device))
            spatial = torch.from_numpy(normalize_boxes(boxes.cpu().numpy(), img_w, img_h)).to(boxes.device)
            self._boxes[key] = (boxes, spatial)
        return self._boxes[key]

    def _normalize_boxes(self, boxes: torch.Tensor, img_w: int, img_h: int) -> torch.Tensor:
        return torch.from_numpy(normalize_boxes(boxes.detach().cpu().numpy(), img_w, img_h)).to(boxes.device)

    # -- the device pipeline for a batch of preprocessed images
    def _run(self, plan: dict):
        t: _Trunk = plan["trunk"]
        fmap = t.base(plan["img"])
        b, fh, fw, ch = fmap.shape
        r = plan["rois"].shape[0]
        pooled = t.buf("roi", (r, self.roi_size, self.roi_size, ch))
        if self.pool_mode == "roi_pool":
            ops.roi_pool_nhwc(fmap, plan["rois"], pooled, 1.0 / 16.0)
        else:
            ops.roi_align_nhwc(fmap, plan["rois"], pooled, 1.0 / 16.0, sampling_ratio=2, aligned=False)
        feats = t.top(pooled)
        plan["feats"].copy_(feats)

    def _plan(self, b: int, h: int, w: int) -> dict:
        trunk = self.backbone.trunk()
        key = (b, h, w)
        plan = self._plans.get(key)
        if plan is not None and plan["trunk"] is trunk:
            return plan
        dev = trunk.device
        boxes, spatial = self._proposals(h, w)
        n = boxes.shape[0]
        rois = torch.cat([torch.arange(b, device=dev, dtype=torch.float32).repeat_interleave(n)[:, None], boxes.repeat(b, 1)], dim=1)
        plan = {"trunk": trunk, "img": torch.zeros(b, 3, h, w, device=dev), "rois": rois.contiguous(),
                "feats": torch.zeros(b * n, 2048, device=dev), "spatial": spatial, "graph": None, "gen": -1, "runs": 0}
        self._plans[key] = plan
        return plan

    @torch.no_grad()
    def extract_batch(self, imgs: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        """Preprocessed (resized, normalised) fp32 NCHW images on the GPU -> ([B,N,2048] fp32, [B,N,5] fp32)."""
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
                self._run(plan)                       # eager: sizes the scratch arena, loads modules, produces this result
                if self.use_graphs:
                    torch.cuda.synchronize()
                    g = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(g):
                        self._run(plan)
                    plan["graph"], plan["gen"] = g, trunk.arena_gen
            n = self.num_regions
            return plan["feats"].view(b, n, -1).clone(), plan["spatial"].unsqueeze(0).expand(b, n, 5).clone()

    @torch.no_grad()
    def extract_features(self, image) -> Tuple[torch.Tensor, torch.Tensor]:
        """Reference ``extract_features`` (:144-178): PIL image -> ([num_regions, output_dim], [num_regions, 5])."""
        img = self.transform(image).unsqueeze(0).to(self.device)
        feats, spatial = self.extract_batch(img)
        return feats[0], spatial[0]

    @torch.no_grad()
    def forward(self, images: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        """Reference ``forward`` (:313-324): the same per-image host preprocessing (via PIL, as the reference does), then ONE
        batched pass through the trunk instead of a Python loop."""
        batch = torch.stack([self.transform(self._to_pil(img.cpu())) for img in images]).to(self.device)
        return self.extract_batch(batch)
