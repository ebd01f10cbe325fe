"""Visual Genome Faster R-CNN region extractor on the RoI stage's kernels: drop-in for the reference's
``models/feature_extractors/fasterrcnn_vg.py`` (``VGFasterRCNN`` :36-97, ``load_vg_weights`` :100-167,
``FasterRCNNVGExtractor`` :170-481).

Same constructor, same parameter tree (``model.RCNN_base / RCNN_top / RCNN_cls_score / RCNN_bbox_pred``: the Visual Genome
checkpoint's names), same ``extract_features(PIL) -> ([num_regions, 2048], [num_regions, 5])`` and
``forward(images) -> ([B, N, 2048], [B, N, 5])``.  The path is the RoI stage's (RoIPool-14 -> layer4 -> mean) behind a
classifier-scored proposal step:

    picture 600 x 1000 -> conv1 .. layer3 (ResNet-101, stride 16)                      resnet152_roi._Trunk.base
    <= 200 sliding-window candidates (host arithmetic, a function of the picture size)  vg_grid_candidates
    [VG checkpoint present] RoIPool-14 of every candidate -> layer4 -> mean -> 1601-way
        class scores (tcgen05 GEMM, fp32 out) -> max over the 1600 object classes        vb_roi_pool_nhwc, _Trunk.top, vb_gemm_bf16, vb_rowmax_f32
    NMS(0.3) -> first num_regions survivors, padded with the last one                    vb_nms, vb_select_regions
    features of the chosen regions + normalised boxes                                   vb_select_regions (rows already computed
                                                                                         for the scores are reused, not recomputed)

Every step runs on the device without a host read, so the whole extraction of one batch shape is ONE CUDA graph.  CUDA only;
there is no CPU or PyTorch fallback.
"""
from __future__ import annotations

import os
from types import SimpleNamespace
from typing import Dict, List, Optional, Tuple

import numpy as np
import torch
import torch.nn as nn

from . import ops
from ._lib import VbError
from .resnet152_roi import _Trunk

NUM_VG_CLASSES = 1601                       # 1600 object classes + background (fasterrcnn_vg.py:47)
_CLS_PAD = (NUM_VG_CLASSES + 7) // 8 * 8    # class-score rows padded to a 16-byte multiple for the TMA store

SCALES = (0.2, 0.3, 0.4, 0.5, 0.7)
ASPECT_RATIOS = (0.5, 1.0, 2.0)


def vg_grid_candidates(img_h: int, img_w: int, num_proposals: int = 100) -> np.ndarray:
    """Candidates of ``_generate_proposals`` (:283-338): 5 scales x 3 aspect ratios, stride half a box, cut at
    2 * num_proposals boxes, topped up with grid cells when fewer than num_proposals came out.  Accumulated in Python doubles
    like the reference, rounded to fp32 once."""
    out: List[List[float]] = []
    limit = num_proposals * 2

    def windows():
        for scale in SCALES:
            for ar in ASPECT_RATIOS:
                bw = img_w * scale
                bh = min(bw / ar, img_h * 0.9)
                bw = min(bw, img_w * 0.9)
                sx, sy = max(bw * 0.5, 1), max(bh * 0.5, 1)
                x = 0
                while x + bw <= img_w:
                    y = 0
                    while y + bh <= img_h:
                        out.append([x, y, x + bw, y + bh])
                        if len(out) >= limit:
                            return
                        y += sy
                    x += sx
    windows()
    if len(out) < num_proposals:
        g = int((num_proposals - len(out)) ** 0.5) + 1
        cw, ch = img_w / g, img_h / g
        out += [[j * cw, i * ch, min((j + 1) * cw, img_w), min((i + 1) * ch, img_h)] for i in range(g) for j in range(g)]
    return np.asarray(out[:limit], dtype=np.float32).reshape(-1, 4)


class VGFasterRCNN(nn.Module):
    """Parameter container with the reference's layout (:36-97).  The torchvision modules only HOLD the parameters; the
    arithmetic runs in ``_Trunk`` and the GEMM.  The three methods keep the reference's NCHW fp32 signatures as boundary
    adapters (the extractor itself stays in NHWC bf16 end to end)."""

    NUM_VG_CLASSES = NUM_VG_CLASSES

    def __init__(self, weights: Optional[str] = "IMAGENET1K_V1"):
        super().__init__()
        from torchvision.models import ResNet101_Weights, resnet101
        resnet = resnet101(weights=None if weights is None else getattr(ResNet101_Weights, weights))
        self.RCNN_base = nn.Sequential(resnet.conv1, resnet.bn1, resnet.relu, resnet.maxpool, resnet.layer1, resnet.layer2,
                                       resnet.layer3)
        self.RCNN_top = resnet.layer4
        self.avgpool = nn.AdaptiveAvgPool2d((1, 1))
        self.RCNN_cls_score = nn.Linear(2048, NUM_VG_CLASSES)
        self.RCNN_bbox_pred = nn.Linear(2048, NUM_VG_CLASSES * 4)
        self._engine: Optional["_Engine"] = None

    def _apply(self, fn, *a, **k):
        out = super()._apply(fn, *a, **k)
        self._engine = None
        return out

    def engine(self) -> "_Engine":
        dev = self.RCNN_cls_score.weight.device
        if dev.type != "cuda":
            raise VbError("VGFasterRCNN (B200) runs on CUDA only; there is no CPU fallback")
        ver = sum(p._version for p in self.parameters()) + sum(b._version for b in self.buffers())
        if self._engine is None or self._engine.version != ver:
            self._engine = _Engine(self, ver)
        return self._engine

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        """Base features (:78-80): fp32 NCHW [B,3,H,W] -> fp32 NCHW [B,1024,H/16,W/16]."""
        return self.engine().trunk.base(x.float().contiguous()).permute(0, 3, 1, 2).float()

    def extract_top_features(self, pooled_features: torch.Tensor) -> torch.Tensor:
        """:82-93: RoI-pooled [N,1024,p,p] -> layer4 -> mean -> [N,2048]."""
        nhwc = pooled_features.permute(0, 2, 3, 1).contiguous().to(torch.bfloat16)
        return self.engine().trunk.top(nhwc).clone()

    def get_class_scores(self, features: torch.Tensor) -> torch.Tensor:
        """:95-97: [N,2048] -> [N,1601]."""
        e = self.engine()
        return e.class_scores(features.float().contiguous())[:, :NUM_VG_CLASSES].clone()


class _Engine:
    """Prepared trunk + classifier operands of one ``VGFasterRCNN``."""

    def __init__(self, model: VGFasterRCNN, version: int):
        self.version = version
        self.trunk = _Trunk(SimpleNamespace(base=model.RCNN_base, top=model.RCNN_top), version)
        dev = self.trunk.device
        with torch.no_grad():
            w = torch.zeros(_CLS_PAD, 2048, dtype=torch.bfloat16, device=dev)
            w[:NUM_VG_CLASSES] = model.RCNN_cls_score.weight.detach().to(torch.bfloat16)
            b = torch.zeros(_CLS_PAD, dtype=torch.float32, device=dev)
            b[:NUM_VG_CLASSES] = model.RCNN_cls_score.bias.detach().float()
        self.cls_w, self.cls_b = w, b

    def class_scores(self, feats: torch.Tensor) -> torch.Tensor:
        """fp32 [n,2048] -> fp32 [n, 1608] (columns >= 1601 are padding)."""
        t = self.trunk
        n = feats.shape[0]
        fb = ops.cast_bf16(feats, t.buf("cls.in", (n, 2048)))
        return ops.gemm(fb, self.cls_w, t.buf("cls.out", (n, _CLS_PAD), torch.float32), bias=self.cls_b)


def load_vg_weights(model: VGFasterRCNN, checkpoint_path: str) -> int:
    """Same contract as the reference loader (:100-167): every checkpoint tensor whose (re-spelled) key exists in the model
    with the same shape is loaded (``RCNN_top.0.X`` is the model's ``RCNN_top.X``); returns how many were."""
    checkpoint = torch.load(checkpoint_path, map_location="cpu")
    state = checkpoint.get("model", checkpoint)
    own = model.state_dict()
    loaded = {}
    for key, value in state.items():
        name = "RCNN_top." + key[len("RCNN_top.0."):] if key.startswith("RCNN_top.0.") else key
        if name in own and own[name].shape == value.shape:
            loaded[name] = value
    model.load_state_dict(loaded, strict=False)
    return len(loaded)


class FasterRCNNVGExtractor(nn.Module):
    """Reference ``FasterRCNNVGExtractor`` (:170-481).  Extra keyword-only arguments: ``weights`` (torchvision weight name or
    None for random init; the reference hard-codes IMAGENET1K_V1) and ``image_size`` ((height, width) the picture is resized
    to; the reference hard-codes (600, 1000))."""

    NUM_VG_CLASSES = NUM_VG_CLASSES
    DEFAULT_WEIGHTS_URL = "https://drive.google.com/file/d/18n_3V1rywgeADZ3oONO0DsuuS9eMW6sN/view"

    def __init__(self, output_dim: int = 2048, num_regions: int = 36, weights_path: Optional[str] = None,
                 confidence_threshold: float = 0.2, nms_threshold: float = 0.3, device: Optional[str] = None, *,
                 weights: Optional[str] = "IMAGENET1K_V1", image_size: Tuple[int, int] = (600, 1000)):
        super().__init__()
        device = "cuda" if device is None else device
        if not str(device).startswith("cuda"):
            raise VbError("FasterRCNNVGExtractor (B200) runs on CUDA only; there is no CPU fallback")
        from torchvision import transforms
        self.output_dim, self.num_regions, self.device = output_dim, num_regions, device
        self.confidence_threshold, self.nms_threshold = confidence_threshold, nms_threshold
        self.weights_path = weights_path
        weights_path = "weights/faster_rcnn_res101_vg.pth" if weights_path is None else weights_path
        self.has_vg_weights = os.path.exists(weights_path)
        self.model = VGFasterRCNN(weights)
        if self.has_vg_weights:
            load_vg_weights(self.model, weights_path)
        self.model.to(device).eval()
        for p in self.model.parameters():
            p.requires_grad = False
        self.image_size = tuple(image_size)
        self.transform = transforms.Compose([
            transforms.Resize(self.image_size), transforms.ToTensor(),
            transforms.Normalize(mean=[0.485, 0.456, 0.406], std=[0.229, 0.224, 0.225])])
        self._to_pil = transforms.ToPILImage()
        self._plans: Dict[Tuple[int, int, int], dict] = {}
        self.use_graphs = True

    # -- the device pipeline for one batch of preprocessed pictures
    def _run(self, plan: dict) -> None:
        e: _Engine = plan["engine"]
        t, b, n, nc = e.trunk, plan["b"], self.num_regions, plan["cands"].shape[0]
        h, w = plan["hw"]
        fmap = t.base(plan["img"])
        ch = fmap.shape[-1]
        scored = self.has_vg_weights
        top = None
        if scored:
            # _score_proposals (:345-365) for every candidate of every picture in one pass
            pooled = t.buf("roi", (b * nc, 14, 14, ch))
            ops.roi_pool_nhwc(fmap, plan["rois"], pooled, 1.0 / 16.0)
            top = t.top(pooled)                                                     # fp32 [b*nc, 2048]
            ops.rowmax(e.class_scores(top), plan["scores"], 1, NUM_VG_CLASSES)
        for i in range(b):
            sl = slice(i * n, (i + 1) * n)
            if nc > n:                                                              # _select_top_regions (:367-392)
                ops.nms_device(plan["cands"], plan["scores"][i * nc:(i + 1) * nc], self.nms_threshold, plan["ws"], plan["keep"][i],
                               plan["nkeep"][i:i + 1])
            # chosen rows: the features computed for the scores are reused (same arithmetic per row as a second pass)
            ops.select_regions(plan["cands"], plan["keep"][i], plan["nkeep"][i:i + 1], n, w, h, boxes=plan["boxes"][sl],
                               spatial=plan["spatial"][sl], index=plan["index"][sl], rois=plan["sel_rois"][sl], batch_index=i,
                               feat_src=top[i * nc:(i + 1) * nc] if scored else None, feat_dst=plan["feats"][sl] if scored else None)
        if not scored:
            # without the checkpoint every candidate scores 1.0 (:340-343): only the chosen regions go through layer4
            pooled = t.buf("roi", (b * n, 14, 14, ch))
            ops.roi_pool_nhwc(fmap, plan["sel_rois"], pooled, 1.0 / 16.0)
            plan["feats"].copy_(t.top(pooled))

    def _plan(self, b: int, h: int, w: int) -> dict:
        engine = self.model.engine()
        key = (b, h, w)
        plan = self._plans.get(key)
        if plan is not None and plan["engine"] is engine and plan["n"] == self.num_regions:
            return plan
        dev, n = engine.trunk.device, self.num_regions
        cands = torch.from_numpy(vg_grid_candidates(h, w)).to(dev)
        nc = cands.shape[0]
        rois = torch.cat([torch.arange(b, device=dev, dtype=torch.float32).repeat_interleave(nc)[:, None], cands.repeat(b, 1)], dim=1)
        keep = torch.arange(max(nc, 1), device=dev, dtype=torch.int32).repeat(b, 1).contiguous()     # nc <= n: identity + padding
        plan = {"engine": engine, "b": b, "n": n, "hw": (h, w), "img": torch.zeros(b, 3, h, w, device=dev), "cands": cands,
                "rois": rois.contiguous(), "scores": torch.ones(b * max(nc, 1), device=dev),
                "ws": torch.zeros(2 * max(nc, 1), dtype=torch.int32, device=dev), "keep": keep,
                "nkeep": torch.full((b,), nc, dtype=torch.int32, device=dev),
                "boxes": torch.zeros(b * n, 4, device=dev), "spatial": torch.zeros(b * n, 5, device=dev),
                "index": torch.zeros(b * n, dtype=torch.int32, device=dev), "feats": torch.zeros(b * n, 2048, device=dev),
                "graph": None, "gen": -1}
        plan["sel_rois"] = torch.zeros(b * n, 5, device=dev)
        self._plans[key] = plan
        return plan

    @torch.no_grad()
    def extract_batch(self, imgs: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        """Preprocessed (resized, normalised) fp32 NCHW pictures on the GPU -> ([B,N,output 2048] fp32, [B,N,5] fp32)."""
        if not imgs.is_cuda:
            raise VbError("extract_batch needs CUDA tensors; there is no CPU fallback")
        b, _, h, w = imgs.shape
        with torch.cuda.device(imgs.device):
            plan = self._plan(b, h, w)
            trunk: _Trunk = plan["engine"].trunk
            plan["img"].copy_(imgs)
            if self.use_graphs and plan["graph"] is not None and plan["gen"] == trunk.arena_gen:
                plan["graph"].replay()
            else:
                self._run(plan)                       # eager: sizes the scratch arena and produces this result
                if self.use_graphs:
                    torch.cuda.synchronize()
                    g = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(g):
                        self._run(plan)
                    plan["graph"], plan["gen"] = g, trunk.arena_gen
            n = self.num_regions
            return plan["feats"].view(b, n, -1).clone(), plan["spatial"].view(b, n, 5).clone()

    # -- the reference's method surface
    @torch.no_grad()
    def extract_features(self, image) -> Tuple[torch.Tensor, torch.Tensor]:
        """:252-281: PIL picture -> ([num_regions, output_dim], [num_regions, 5])."""
        feats, spatial = self.extract_batch(self.transform(image).unsqueeze(0).to(self.device))
        return feats[0], spatial[0]

    @torch.no_grad()
    def forward(self, images: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        """:471-481: the same per-picture host preprocessing via PIL, then ONE batched pass."""
        batch = torch.stack([self.transform(self._to_pil(img.cpu())) for img in images]).to(self.device)
        return self.extract_batch(batch)

    def selected(self, b: int = 1) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
        """(boxes [B,N,4], candidate index [B,N], candidate scores [B,nc]) of the last ``extract_batch`` of that batch size at
        the configured picture size -- what ``_select_top_regions`` returned in the reference; for tests and diagnostics."""
        plan = self._plans[(b, *self.image_size)]
        n = self.num_regions
        return plan["boxes"].view(b, n, 4).clone(), plan["index"].view(b, n).long(), plan["scores"].view(b, -1).clone()
