"""Visual Genome Faster R-CNN region extractor with RPN proposals on the RoI stage's kernels: drop-in for
the reference's ``models/feature_extractors/fasterrcnn_vg_rpn.py`` (``RPN`` :34-174, ``VGFasterRCNNWithRPN`` :177-239,
``load_vg_checkpoint`` :242-287, ``FasterRCNNVGRPNExtractor`` :290-563).

Same constructor, same parameter tree (``model.RCNN_base / RCNN_top / RCNN_rpn.{RPN_Conv, RPN_cls_score, RPN_bbox_pred} /
RCNN_cls_score / RCNN_bbox_pred``), same ``extract_features(PIL) -> ([num_regions, 2048], [num_regions, 5])``.  One picture =
one CUDA graph, no host read inside:

    picture resized to 600 / max 1000 (aspect kept, host, PIL)               _resize_image
    conv1 .. layer3 (ResNet-101, stride 16)                                   resnet152_roi._Trunk.base
    RPN: 3x3 conv + ReLU (implicit GEMM), objectness | box-delta heads (one GEMM, fp32 out)        vb_gemm_bf16
    softmax, 12 anchors per cell, deltas, clipping, min-size filter          vb_rpn_decode
    top pre_nms_top_n by objectness, NMS(0.7), first post_nms_top_n          vb_rank_sort_desc, vb_gather_sorted, vb_nms_sorted
    RoIPool-14 -> layer4 -> mean -> 1601-way class scores -> max             vb_roi_pool_nhwc, _Trunk.top, vb_gemm_bf16, vb_rowmax_f32
    top num_regions by class score, boxes / scale normalised                 vb_rank_sort_desc, vb_select_regions

Only when fewer than ``num_regions`` proposals survive (the reference pads with grid cells, :490-535) does the host read a count and
run that branch step by step.  CUDA only; there is no CPU or PyTorch fallback.
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
from .fasterrcnn_vg import _CLS_PAD, NUM_VG_CLASSES
from .resnet152_roi import _Trunk

ANCHOR_SCALES, ANCHOR_RATIOS, FEAT_STRIDE = (4, 8, 16, 32), (0.5, 1.0, 2.0), 16


def base_anchors() -> np.ndarray:
    """``RPN._generate_anchors`` (:110-118): 4 scales x 3 ratios around the origin, Python doubles rounded to fp32 once."""
    rows = []
    for scale in ANCHOR_SCALES:
        for ratio in ANCHOR_RATIOS:
            h = scale * FEAT_STRIDE * (ratio ** 0.5)
            w = scale * FEAT_STRIDE / (ratio ** 0.5)
            rows.append([-w / 2, -h / 2, w / 2, h / 2])
    return np.asarray(rows, dtype=np.float32)


def pad_grid_boxes(num_needed: int, img_w: int, img_h: int) -> np.ndarray:
    """Grid cells appended by ``_pad_regions`` (:497-516)."""
    g = int(num_needed ** 0.5) + 1
    cw, ch = img_w / g, img_h / g
    rows: List[List[float]] = []
    for i in range(g):
        for j in range(g):
            if len(rows) >= num_needed:
                break
            rows.append([j * cw, i * ch, min((j + 1) * cw, img_w), min((i + 1) * ch, img_h)])
        if len(rows) >= num_needed:
            break
    return np.asarray(rows, dtype=np.float32).reshape(-1, 4)


class RPN(nn.Module):
    """Parameter container with the checkpoint's names (:34-58); the arithmetic runs in ``_Engine``."""

    def __init__(self, in_channels: int = 1024, num_anchors: int = 12):
        super().__init__()
        self.num_anchors = num_anchors
        self.RPN_Conv = nn.Conv2d(in_channels, 512, kernel_size=3, padding=1)
        self.RPN_cls_score = nn.Conv2d(512, num_anchors * 2, kernel_size=1)
        self.RPN_bbox_pred = nn.Conv2d(512, num_anchors * 4, kernel_size=1)
        self.anchor_scales, self.anchor_ratios, self.feat_stride = list(ANCHOR_SCALES), list(ANCHOR_RATIOS), FEAT_STRIDE


class VGFasterRCNNWithRPN(nn.Module):
    """Reference class of the same name (:177-239): parameter container + NCHW fp32 boundary adapters."""

    NUM_VG_CLASSES = NUM_VG_CLASSES

    def __init__(self, weights: Optional[str] = "IMAGENET1K_V1"):
        super().__init__()
        from torchvision.models import ResNet101_Weights, resnet101
        resnet = resnet101(weights=None if weights is None else getattr(ResNet101_Weights, weights))
        self.RCNN_base = nn.Sequential(resnet.conv1, resnet.bn1, resnet.relu, resnet.maxpool, resnet.layer1, resnet.layer2,
                                       resnet.layer3)
        self.RCNN_top = resnet.layer4
        self.RCNN_rpn = RPN(in_channels=1024, num_anchors=12)
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
            raise VbError("VGFasterRCNNWithRPN (B200) runs on CUDA only; there is no CPU fallback")
        ver = sum(p._version for p in self.parameters()) + sum(b._version for b in self.buffers())
        if self._engine is None or self._engine.version != ver:
            self._engine = _Engine(self, ver)
        return self._engine

    def get_base_features(self, x: torch.Tensor) -> torch.Tensor:
        """:220-222: fp32 NCHW [B,3,H,W] -> fp32 NCHW [B,1024,H/16,W/16]."""
        return self.engine().trunk.base(x.float().contiguous()).permute(0, 3, 1, 2).float()

    def get_proposals(self, features: torch.Tensor, img_size: Tuple[int, int]) -> Tuple[torch.Tensor, torch.Tensor]:
        """:224-228 = ``RPN.forward`` (:60-104): NCHW fp32 features [1,1024,fh,fw] -> (proposals [fh*fw*12, 4] clipped to the
        picture, foreground probabilities [fh*fw*12]); no size filter (that is ``_filter_proposals``)."""
        e = self.engine()
        fmap = features.permute(0, 2, 3, 1).contiguous().to(torch.bfloat16)
        _, fh, fw, _ = fmap.shape
        a = fh * fw * 12
        props, scores = torch.empty(a, 4, device=fmap.device), torch.empty(a, device=fmap.device)
        nv = torch.zeros(1, dtype=torch.int32, device=fmap.device)
        ops.rpn_decode(e.rpn_heads(fmap), fh, fw, e.anchors, FEAT_STRIDE, img_size[0], img_size[1], 0.0, props, scores, nv)
        return props, scores

    def extract_roi_features(self, pooled: torch.Tensor) -> torch.Tensor:
        """:230-235: RoI-pooled [N,1024,p,p] -> layer4 -> mean -> [N,2048]."""
        return self.engine().trunk.top(pooled.permute(0, 2, 3, 1).contiguous().to(torch.bfloat16)).clone()

    def get_class_scores(self, features: torch.Tensor) -> torch.Tensor:
        """:237-239: [N,2048] -> [N,1601]."""
        return self.engine().class_scores(features.float().contiguous())[:, :NUM_VG_CLASSES].clone()


class _Engine:
    """Prepared trunk + RPN + classifier operands of one ``VGFasterRCNNWithRPN``."""

    def __init__(self, model: VGFasterRCNNWithRPN, version: int):
        self.version = version
        self.trunk = _Trunk(SimpleNamespace(base=model.RCNN_base, top=model.RCNN_top), version)
        dev = self.trunk.device
        rpn = model.RCNN_rpn
        with torch.no_grad():
            w = torch.zeros(_CLS_PAD, 2048, dtype=torch.bfloat16, device=dev)
            w[:NUM_VG_CLASSES] = model.RCNN_cls_score.weight.detach().to(torch.bfloat16)
            b = torch.zeros(_CLS_PAD, dtype=torch.float32, device=dev)
            b[:NUM_VG_CLASSES] = model.RCNN_cls_score.bias.detach().float()
            self.cls_w, self.cls_b = w, b
            cw = rpn.RPN_Conv.weight.detach()                                      # [512, 1024, 3, 3] -> [512, (ky*3+kx)*1024 + ci]
            self.rpn_w = cw.permute(0, 2, 3, 1).reshape(cw.shape[0], -1).to(torch.bfloat16).contiguous()
            self.rpn_b = rpn.RPN_Conv.bias.detach().float().contiguous()
            na = rpn.num_anchors
            hw = torch.cat([rpn.RPN_cls_score.weight.detach().reshape(2 * na, -1), rpn.RPN_bbox_pred.weight.detach().reshape(4 * na, -1)])
            self.head_w = hw.to(torch.bfloat16).contiguous()                       # [6A, 512]: objectness rows, then delta rows
            self.head_b = torch.cat([rpn.RPN_cls_score.bias.detach(), rpn.RPN_bbox_pred.bias.detach()]).float().contiguous()
        self.num_anchors = na
        self.anchors = base_anchors()

    def rpn_heads(self, fmap: torch.Tensor) -> torch.Tensor:
        """bf16 NHWC [1,fh,fw,1024] -> fp32 [fh*fw, 6A]: relu(conv3x3) (implicit GEMM), then both 1x1 heads in one GEMM."""
        t = self.trunk
        _, fh, fw, _ = fmap.shape
        x = ops.gemm(fmap, self.rpn_w, t.buf("rpn.x", (fh * fw, self.rpn_w.shape[0])), bias=self.rpn_b, act=ops.ACT_RELU, conv=(3, 3, 1, 1))
        return ops.gemm(x, self.head_w, t.buf("rpn.heads", (fh * fw, self.head_w.shape[0]), torch.float32), bias=self.head_b)

    def class_scores(self, feats: torch.Tensor) -> torch.Tensor:
        t = self.trunk
        n = feats.shape[0]
        fb = ops.cast_bf16(feats, t.buf("cls.in", (n, 2048)))
        return ops.gemm(fb, self.cls_w, t.buf("cls.out", (n, _CLS_PAD), torch.float32), bias=self.cls_b)


def load_vg_checkpoint(model: VGFasterRCNNWithRPN, checkpoint_path: str) -> dict:
    """Same contract as the reference loader (:242-287): statistics {"loaded", "total", "skipped"}."""
    checkpoint = torch.load(checkpoint_path, map_location="cpu")
    state = checkpoint.get("model", checkpoint)
    own = model.state_dict()
    loaded, skipped = {}, {}
    for key, value in state.items():
        name = "RCNN_top." + key[len("RCNN_top.0."):] if key.startswith("RCNN_top.0.") else key
        if name not in own:
            skipped[key] = "key not in model"
        elif own[name].shape != value.shape:
            skipped[key] = f"shape mismatch: {own[name].shape} vs {value.shape}"
        else:
            loaded[name] = value
    model.load_state_dict(loaded, strict=False)
    return {"loaded": len(loaded), "total": len(own), "skipped": len(skipped)}


class FasterRCNNVGRPNExtractor(nn.Module):
    """Reference ``FasterRCNNVGRPNExtractor`` (:290-563).  Extra keyword-only argument: ``weights`` (torchvision weight name or
    None for random init; the reference hard-codes IMAGENET1K_V1)."""

    def __init__(self, output_dim: int = 2048, num_regions: int = 36, weights_path: Optional[str] = None, nms_threshold: float = 0.7,
                 pre_nms_top_n: int = 6000, post_nms_top_n: int = 300, min_box_size: float = 16, device: Optional[str] = None, *,
                 weights: Optional[str] = "IMAGENET1K_V1"):
        super().__init__()
        device = "cuda" if device is None else device
        if not str(device).startswith("cuda"):
            raise VbError("FasterRCNNVGRPNExtractor (B200) runs on CUDA only; there is no CPU fallback")
        if pre_nms_top_n > 8192:
            raise VbError("pre_nms_top_n above 8192 is not supported by vb_nms_sorted")
        from torchvision import transforms
        self.output_dim, self.num_regions, self.device = output_dim, num_regions, device
        self.nms_threshold, self.pre_nms_top_n, self.post_nms_top_n = nms_threshold, pre_nms_top_n, post_nms_top_n
        self.min_box_size = min_box_size
        weights_path = "weights/faster_rcnn_res101_vg.pth" if weights_path is None else weights_path
        self.model = VGFasterRCNNWithRPN(weights)
        self.has_vg_weights = os.path.exists(weights_path) and load_vg_checkpoint(self.model, weights_path)["loaded"] > 0
        self.model.to(device).eval()
        for p in self.model.parameters():
            p.requires_grad = False
        self.transform = transforms.Compose([transforms.ToTensor(),
                                             transforms.Normalize(mean=[0.485, 0.456, 0.406], std=[0.229, 0.224, 0.225])])
        self._to_pil = transforms.ToPILImage()
        self.target_size, self.max_size = 600, 1000
        self._plans: Dict[tuple, dict] = {}
        self.use_graphs = True

    def _resize_image(self, image):
        """:373-385: aspect-preserving resize to a 600-pixel short side, capped at 1000 on the long side."""
        from PIL import Image
        w, h = image.size
        scale = self.target_size / min(w, h)
        if max(w, h) * scale > self.max_size:
            scale = self.max_size / max(w, h)
        return image.resize((int(w * scale), int(h * scale)), Image.BILINEAR), scale

    # -- the device pipeline of one picture
    def _run(self, plan: dict) -> None:
        e: _Engine = plan["engine"]
        t, n = e.trunk, self.num_regions
        h, w = plan["hw"]
        post = self.post_nms_top_n
        fmap = t.base(plan["img"])
        _, fh, fw, ch = fmap.shape
        ops.rpn_decode(e.rpn_heads(fmap), fh, fw, e.anchors, FEAT_STRIDE, h, w, self.min_box_size, plan["props"], plan["scores"],
                       plan["n_valid"])
        ops.rank_sort_desc(plan["scores"], plan["order"])                                              # :455-458 (top-k = prefix)
        ops.gather_sorted(plan["props"], plan["scores"], plan["order"], plan["n_valid"], plan["top_boxes"], plan["top_scores"],
                          plan["n_top"])
        ops.nms_sorted(plan["top_boxes"], plan["n_top"], self.nms_threshold, plan["keep"], plan["n_keep"])   # :461-465
        # the survivors as RoIs (rows beyond the count repeat the last one; they are masked out of the ranking below)
        ops.select_regions(plan["top_boxes"], plan["keep"], plan["n_keep"], post, w, h, boxes=plan["kept_boxes"], rois=plan["kept_rois"])
        pooled = t.buf("roi", (post, 14, 14, ch))
        ops.roi_pool_nhwc(fmap, plan["kept_rois"], pooled, 1.0 / 16.0)
        top = t.top(pooled)                                                                            # fp32 [post, 2048]
        ops.rowmax(e.class_scores(top), plan["region_scores"], 1, NUM_VG_CLASSES)                      # :419-422
        ops.rank_sort_desc(plan["region_scores"], plan["region_order"], limit=plan["n_keep"])          # :427-430
        ops.select_regions(plan["kept_boxes"], plan["region_order"], plan["n_keep"], n, plan["orig_wh"][0], plan["orig_wh"][1],
                           boxes=plan["boxes"], spatial=plan["spatial"], index=plan["index"], feat_src=top, feat_dst=plan["feats"],
                           box_div=plan["scale"])
        plan["top_feats"] = top

    def _plan(self, h: int, w: int, orig_w: int, orig_h: int, scale: float) -> dict:
        engine = self.model.engine()
        key = (h, w, orig_w, orig_h, self.num_regions)
        plan = self._plans.get(key)
        if plan is not None and plan["engine"] is engine:
            return plan
        dev, n, post, pre = engine.trunk.device, self.num_regions, self.post_nms_top_n, self.pre_nms_top_n
        fh, fw = self._fmap_hw(h, w)
        a = fh * fw * engine.num_anchors
        i32 = dict(dtype=torch.int32, device=dev)
        plan = {"engine": engine, "hw": (h, w), "orig_wh": (orig_w, orig_h), "scale": scale, "img": torch.zeros(1, 3, h, w, device=dev),
                "props": torch.zeros(a, 4, device=dev), "scores": torch.zeros(a, device=dev), "order": torch.zeros(a, **i32),
                "n_valid": torch.zeros(1, **i32), "top_boxes": torch.zeros(pre, 4, device=dev), "top_scores": torch.zeros(pre, device=dev),
                "n_top": torch.zeros(1, **i32), "keep": torch.zeros(post, **i32), "n_keep": torch.zeros(1, **i32),
                "kept_boxes": torch.zeros(post, 4, device=dev), "kept_rois": torch.zeros(post, 5, device=dev),
                "region_scores": torch.zeros(post, device=dev), "region_order": torch.zeros(post, **i32),
                "boxes": torch.zeros(n, 4, device=dev), "spatial": torch.zeros(n, 5, device=dev), "index": torch.zeros(n, **i32),
                "feats": torch.zeros(n, 2048, device=dev), "graph": None, "gen": -1, "top_feats": None}
        self._plans[key] = plan
        return plan

    @staticmethod
    def _fmap_hw(h: int, w: int) -> Tuple[int, int]:
        """Spatial size of the stride-16 map: conv1 7x7/2 pad 3, max-pool 3x3/2 pad 1, two stride-2 3x3 pad 1 stages."""
        def down(v, k, s, p):
            return (v + 2 * p - k) // s + 1
        for k, s, p in ((7, 2, 3), (3, 2, 1), (3, 2, 1), (3, 2, 1)):
            h, w = down(h, k, s, p), down(w, k, s, p)
        return h, w

    @torch.no_grad()
    def extract_preprocessed(self, img: torch.Tensor, scale: float, orig_w: int, orig_h: int) -> Tuple[torch.Tensor, torch.Tensor]:
        """A resized, normalised fp32 NCHW picture [1,3,h,w] on the GPU (+ its resize factor and original size) ->
        ([N, 2048] fp32, [N, 5] fp32)."""
        if not img.is_cuda:
            raise VbError("extract_preprocessed needs CUDA tensors; there is no CPU fallback")
        _, _, h, w = img.shape
        with torch.cuda.device(img.device):
            plan = self._plan(h, w, orig_w, orig_h, scale)
            trunk: _Trunk = plan["engine"].trunk
            plan["img"].copy_(img)
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
            feats, spatial = plan["feats"].clone(), plan["spatial"].clone()
            m = int(plan["n_keep"].item())            # the one host read: did enough proposals survive?
            if m < self.num_regions:
                feats, spatial = self._pad_regions(plan, m)
            return feats, spatial

    def _pad_regions(self, plan: dict, m: int) -> Tuple[torch.Tensor, torch.Tensor]:
        """Fewer survivors than regions (:431-434, 490-535): the survivors in NMS order, then grid cells of the resized picture,
        each through RoIPool-14 -> layer4 -> mean; boxes / scale normalised on the host exactly like the device path."""
        from .resnet152_roi import normalize_boxes
        e: _Engine = plan["engine"]
        t, n = e.trunk, self.num_regions
        h, w = plan["hw"]
        dev = t.device
        kept = plan["kept_boxes"][:m]
        grid = torch.from_numpy(pad_grid_boxes(n - m, w, h)).to(dev)
        rois = torch.cat([torch.zeros(grid.shape[0], 1, device=dev), grid], dim=1).contiguous()
        fmap = t.base(plan["img"])
        pooled = t.buf("roi.pad", (grid.shape[0], 14, 14, fmap.shape[-1]))
        ops.roi_pool_nhwc(fmap, rois, pooled, 1.0 / 16.0)
        grid_feats = t.top(pooled).clone()
        # the survivors' rows: recompute them too (the arena buffer behind plan["top_feats"] was just reused)
        kp = t.buf("roi.pad2", (max(m, 1), 14, 14, fmap.shape[-1]))
        feats = grid_feats
        if m > 0:
            ops.roi_pool_nhwc(fmap, plan["kept_rois"][:m].contiguous(), kp[:m], 1.0 / 16.0)
            feats = torch.cat([t.top(kp[:m]).clone(), grid_feats], dim=0)
        boxes = torch.cat([kept, grid], dim=0)[:n]
        scaled = (boxes.cpu().numpy().astype(np.float32) / np.float32(plan["scale"])).astype(np.float32)
        spatial = torch.from_numpy(normalize_boxes(scaled, plan["orig_wh"][0], plan["orig_wh"][1])).to(dev)
        return feats[:n].contiguous(), spatial

    # -- the reference's method surface
    @torch.no_grad()
    def extract_features(self, image) -> Tuple[torch.Tensor, torch.Tensor]:
        """:387-440: PIL picture -> ([num_regions, 2048], [num_regions, 5])."""
        orig_w, orig_h = image.size
        resized, scale = self._resize_image(image)
        img = self.transform(resized).unsqueeze(0).to(self.device)
        return self.extract_preprocessed(img, scale, orig_w, orig_h)

    @torch.no_grad()
    def forward(self, images: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        """:552-563: per picture (their resized sizes differ), stacked."""
        outs = [self.extract_features(self._to_pil(img.cpu())) for img in images]
        return torch.stack([f for f, _ in outs]), torch.stack([s for _, s in outs])

    def selected(self, key=None) -> dict:
        """Intermediate results of the last ``extract_features`` (tests / diagnostics): proposals, objectness, survivors, class
        scores, the chosen boxes and their positions among the survivors."""
        plan = self._plans[key] if key is not None else list(self._plans.values())[-1]
        m = int(plan["n_keep"].item())
        return {"proposals": plan["props"].clone(), "scores": plan["scores"].clone(), "n_valid": int(plan["n_valid"].item()),
                "kept_boxes": plan["kept_boxes"][:m].clone(), "region_scores": plan["region_scores"][:m].clone(),
                "boxes": plan["boxes"].clone(), "index": plan["index"].long().clone(), "n_keep": m}
