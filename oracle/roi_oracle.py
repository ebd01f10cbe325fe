"""ORACLE (test infrastructure, never shipped, never on the product path).

CPU, fp32 restatement of the reference's ResNet-152 RoI feature stage
(/root/reference/src/multimodalclassification/models/feature_extractors/resnet152_roi.py), written as pure functions over a
``state_dict`` with the reference backbone's key names (``base.0.weight``, ``base.4.0.conv1.weight``, ``top.2.bn3.bias``
...) so that it travels to the GPU box (the reference itself does not).  Only ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s CPU legs may import it.

Third-party arithmetic on this path (not under /root/reference; versions unpinned by the reference):
torchvision 0.26 ``resnet152`` (Bottleneck v1.5: stride on the 3x3), ``ops.RoIPool`` and ``ops.nms``.  Their published
algorithms are restated here (``roi_pool``, ``nms``: plain numpy loops following torchvision/csrc/ops/cpu/*.cpp).

Pinning: the reference holds no golden vectors for this path, so ``oracle/make_golden_roi.py`` imports the
reference extractor in the authoring container (torchvision weights replaced by ``seeded_backbone_state`` below, as there is
no network for the ImageNet checkpoint), runs it on a seeded image and commits proposals / RoIPool / NMS / feature vectors
under ``tests/golden/roi_*.npz``; ``tests/test_roi_oracle_cpu.py`` checks this file against them: bit-equal for boxes, NMS
order, RoIPool and normalised boxes; <= 1e-4 relative for the fp32 features (different conv summation order).
"""
from __future__ import annotations

import math
from typing import Dict, List, Tuple

import numpy as np
import torch
import torch.nn.functional as F

LAYERS = (("base.4", 3, 64, 1), ("base.5", 8, 128, 2), ("base.6", 36, 256, 2), ("top", 3, 512, 2))   # torchvision resnet152


# ------------------------------------------------------------------------------------------------ seeded weights
def seeded_backbone_state(seed: int = 0, blocks: Tuple[int, int, int, int] = (3, 8, 36, 3)) -> Dict[str, torch.Tensor]:
    """Deterministic stand-in for the ImageNet checkpoint, keyed like the reference's ``ResNet152Backbone.state_dict()``
    (``blocks`` = bottlenecks per stage: torchvision resnet152 by default, (3, 4, 23, 3) = resnet101).
    He-normal convolutions; BatchNorm statistics and affine parameters drawn so that folding them is not a no-op; the last
    BatchNorm of every bottleneck is damped so that 50 un-normalised residual additions keep activations O(1)."""
    g = torch.Generator().manual_seed(seed)
    sd: Dict[str, torch.Tensor] = {}

    def conv(name, cout, cin, k):
        sd[name + ".weight"] = torch.randn(cout, cin, k, k, generator=g) * math.sqrt(2.0 / (cin * k * k))

    def bn(name, c, gain):
        sd[name + ".weight"] = (0.75 + 0.5 * torch.rand(c, generator=g)) * gain
        sd[name + ".bias"] = (torch.rand(c, generator=g) - 0.5) * 0.2
        sd[name + ".running_mean"] = (torch.rand(c, generator=g) - 0.5) * 0.2
        sd[name + ".running_var"] = 0.75 + 0.5 * torch.rand(c, generator=g)
        sd[name + ".num_batches_tracked"] = torch.tensor(0, dtype=torch.long)

    conv("base.0", 64, 3, 7)
    bn("base.1", 64, 1.0)
    cin = 64
    for (prefix, _, width, stride), nblocks in zip(LAYERS, blocks):
        for b in range(nblocks):
            p = f"{prefix}.{b}"
            conv(p + ".conv1", width, cin, 1); bn(p + ".bn1", width, 1.0)
            conv(p + ".conv2", width, width, 3); bn(p + ".bn2", width, 1.0)
            conv(p + ".conv3", width * 4, width, 1); bn(p + ".bn3", width * 4, 0.25)
            if b == 0:
                conv(p + ".downsample.0", width * 4, cin, 1); bn(p + ".downsample.1", width * 4, 0.7)
            cin = width * 4
    return sd


# ------------------------------------------------------------------------------------------------ ResNet-152 trunk
def _bn(sd, name, x):
    return F.batch_norm(x, sd[name + ".running_mean"], sd[name + ".running_var"], sd[name + ".weight"], sd[name + ".bias"],
                        training=False, eps=1e-5)


def _bottleneck(sd, p, x, stride):
    """torchvision.models.resnet.Bottleneck.forward (v1.5)."""
    out = F.relu(_bn(sd, p + ".bn1", F.conv2d(x, sd[p + ".conv1.weight"])))
    out = F.relu(_bn(sd, p + ".bn2", F.conv2d(out, sd[p + ".conv2.weight"], stride=stride, padding=1)))
    out = _bn(sd, p + ".bn3", F.conv2d(out, sd[p + ".conv3.weight"]))
    if (p + ".downsample.0.weight") in sd:
        x = _bn(sd, p + ".downsample.1", F.conv2d(x, sd[p + ".downsample.0.weight"], stride=stride))
    return F.relu(out + x)


def _layer(sd, prefix, blocks, stride, x):
    """``blocks`` is the torchvision resnet152 count; the count actually present in ``sd`` wins (resnet101 weights)."""
    b = 0
    while f"{prefix}.{b}.conv1.weight" in sd:
        x = _bottleneck(sd, f"{prefix}.{b}", x, stride if b == 0 else 1)
        b += 1
    return x


def forward_base(sd, img: torch.Tensor) -> torch.Tensor:
    """resnet152_roi.py:49-57, 65-67: conv1, bn1, relu, maxpool, layer1..3.  [B,3,H,W] -> [B,1024,H/16,W/16]."""
    x = F.relu(_bn(sd, "base.1", F.conv2d(img, sd["base.0.weight"], stride=2, padding=3)))
    x = F.max_pool2d(x, 3, 2, 1)
    for prefix, blocks, _, stride in LAYERS[:3]:
        x = _layer(sd, prefix, blocks, stride, x)
    return x


def forward_top(sd, pooled: torch.Tensor) -> torch.Tensor:
    """resnet152_roi.py:69-74: layer4, global average pool, flatten.  [R,1024,p,p] -> [R,2048]."""
    x = _layer(sd, "top", 3, 2, pooled)
    return x.mean(dim=(2, 3))


# ------------------------------------------------------------------------------------------------ RoIPool / NMS / boxes
def _round_half_away(x: np.float32) -> int:
    return int(math.floor(float(x) + 0.5)) if x >= 0 else -int(math.floor(-float(x) + 0.5))


def roi_pool(fmap: np.ndarray, rois: np.ndarray, pooled: int, spatial_scale: float) -> np.ndarray:
    """torchvision.ops.roi_pool (csrc/ops/cpu/roi_pool_kernel.cpp): fmap [N,C,H,W] fp32, rois [R,5] -> [R,C,p,p]."""
    _, c, h, w = fmap.shape
    out = np.zeros((rois.shape[0], c, pooled, pooled), dtype=np.float32)
    ss = np.float32(spatial_scale)
    for r, roi in enumerate(rois.astype(np.float32)):
        b = int(roi[0])
        x1, y1 = _round_half_away(roi[1] * ss), _round_half_away(roi[2] * ss)
        x2, y2 = _round_half_away(roi[3] * ss), _round_half_away(roi[4] * ss)
        rw, rh = max(x2 - x1 + 1, 1), max(y2 - y1 + 1, 1)
        bh, bw = np.float32(rh) / np.float32(pooled), np.float32(rw) / np.float32(pooled)
        for ph in range(pooled):
            hs = min(max(int(math.floor(np.float32(ph) * bh)) + y1, 0), h)
            he = min(max(int(math.ceil(np.float32(ph + 1) * bh)) + y1, 0), h)
            for pw in range(pooled):
                ws = min(max(int(math.floor(np.float32(pw) * bw)) + x1, 0), w)
                we = min(max(int(math.ceil(np.float32(pw + 1) * bw)) + x1, 0), w)
                if he > hs and we > ws:
                    out[r, :, ph, pw] = fmap[b, :, hs:he, ws:we].reshape(c, -1).max(axis=1)
    return out


def nms(boxes: np.ndarray, scores: np.ndarray, thr: float) -> np.ndarray:
    """torchvision.ops.nms, CPU kernel (csrc/ops/cpu/nms_kernel.cpp): stable descending sort, greedy, fp32 arithmetic."""
    boxes = boxes.astype(np.float32)
    x1, y1, x2, y2 = boxes[:, 0], boxes[:, 1], boxes[:, 2], boxes[:, 3]
    areas = (x2 - x1) * (y2 - y1)
    order = np.argsort(-scores.astype(np.float32), kind="stable")
    suppressed = np.zeros(len(boxes), dtype=bool)
    keep: List[int] = []
    for oi, i in enumerate(order):
        if suppressed[i]:
            continue
        keep.append(int(i))
        rest = order[oi + 1:]
        xx1, yy1 = np.maximum(x1[i], x1[rest]), np.maximum(y1[i], y1[rest])
        xx2, yy2 = np.minimum(x2[i], x2[rest]), np.minimum(y2[i], y2[rest])
        w = np.maximum(np.float32(0), xx2 - xx1)
        h = np.maximum(np.float32(0), yy2 - yy1)
        inter = w * h
        ovr = inter / (areas[i] + areas[rest] - inter)
        suppressed[rest[ovr.astype(np.float64) > float(thr)]] = True      # the C++ kernel's threshold is a double
    return np.asarray(keep, dtype=np.int64)


def grid_proposals(num_regions: int, img_h: int, img_w: int) -> np.ndarray:
    """resnet152_roi.py:191-206."""
    g = int(num_regions ** 0.5)
    cell_h, cell_w = img_h / g, img_w / g
    rows = []
    for i in range(g):
        for j in range(g):
            rows.append([j * cell_w, i * cell_h, (j + 1) * cell_w, (i + 1) * cell_h])
    return np.array(rows, dtype=np.float32).reshape(-1, 4)


def multi_scale_candidates(img_h: int, img_w: int) -> np.ndarray:
    """resnet152_roi.py:208-240 (Python-double accumulation, one rounding to fp32 at the end)."""
    rows = []
    for scale in (0.15, 0.25, 0.35, 0.5, 0.7):
        for ar in (0.5, 0.75, 1.0, 1.33, 2.0):
            box_w = img_w * scale
            box_h = box_w / ar
            box_h = min(box_h, img_h * 0.95)
            box_w = min(box_w, img_w * 0.95)
            stride_x, stride_y = max(box_w * 0.4, 20), max(box_h * 0.4, 20)
            x = 0
            while x + box_w <= img_w:
                y = 0
                while y + box_h <= img_h:
                    rows.append([x, y, x + box_w, y + box_h])
                    y += stride_y
                x += stride_x
    return np.array(rows, dtype=np.float32).reshape(-1, 4)


def area_scores(boxes: np.ndarray, img_h: int, img_w: int) -> np.ndarray:
    """resnet152_roi.py:262-270: 1 - |w/W * h/H - 0.15| in fp32, one rounding per operation."""
    b = boxes.astype(np.float32)
    widths = (b[:, 2] - b[:, 0]) / np.float32(img_w)
    heights = (b[:, 3] - b[:, 1]) / np.float32(img_h)
    return (np.float32(1.0) - np.abs(widths * heights - np.float32(0.15))).astype(np.float32)


def proposals(num_regions: int, img_h: int, img_w: int, multi_scale: bool = True) -> np.ndarray:
    """resnet152_roi.py:180-293."""
    if not multi_scale:
        return grid_proposals(num_regions, img_h, img_w)
    boxes = multi_scale_candidates(img_h, img_w)
    if len(boxes) > num_regions:
        keep = nms(boxes, area_scores(boxes, img_h, img_w), 0.5)
        if len(keep) < num_regions:
            kept = set(keep.tolist())
            rest = [i for i in range(len(boxes)) if i not in kept]
            keep = np.concatenate([keep, np.asarray(rest[: num_regions - len(keep)], dtype=np.int64)])
        boxes = boxes[keep[:num_regions]]
    elif len(boxes) < num_regions:
        boxes = np.concatenate([boxes, grid_proposals(num_regions, img_h, img_w)], axis=0)[:num_regions]
    return boxes[:num_regions]


def normalize_boxes(boxes: np.ndarray, img_w: int, img_h: int) -> np.ndarray:
    """resnet152_roi.py:295-311."""
    nb = boxes.astype(np.float32).copy()
    nb[:, 0] /= np.float32(img_w); nb[:, 2] /= np.float32(img_w)
    nb[:, 1] /= np.float32(img_h); nb[:, 3] /= np.float32(img_h)
    nb = np.minimum(np.maximum(nb, np.float32(0)), np.float32(1))
    areas = (nb[:, 2] - nb[:, 0]) * (nb[:, 3] - nb[:, 1])
    return np.concatenate([nb, areas[:, None]], axis=1)


# ------------------------------------------------------------------------------------------------ whole stage
def extract_features(sd, img: torch.Tensor, num_regions: int = 36, roi_size: int = 14, multi_scale: bool = True
                     ) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    """resnet152_roi.py:144-178 on an already preprocessed image [1,3,H,W] (fp32): (features [N,2048], spatial [N,5], boxes)."""
    h, w = img.shape[2], img.shape[3]
    with torch.no_grad():
        fmap = forward_base(sd, img)
        boxes = proposals(num_regions, h, w, multi_scale)
        rois = np.concatenate([np.zeros((len(boxes), 1), np.float32), boxes], axis=1)
        pooled = roi_pool(fmap.numpy(), rois, roi_size, 1.0 / 16.0)
        feats = forward_top(sd, torch.from_numpy(pooled))
    return feats.numpy(), normalize_boxes(boxes, w, h), boxes


def synthetic_image(seed: int = 7, h: int = 96, w: int = 128) -> np.ndarray:
    """uint8 HxWx3 test picture: smooth gradients + blocks + noise (so that resizing, pooling and max selection all matter)."""
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:h, 0:w].astype(np.float32)
    img = np.stack([xx / w * 255, yy / h * 255, (xx + yy) / (h + w) * 255], axis=2)
    for _ in range(12):
        y0, x0 = rng.integers(0, h - 8), rng.integers(0, w - 8)
        img[y0:y0 + rng.integers(4, 24), x0:x0 + rng.integers(4, 32)] = rng.integers(0, 255, size=3)
    img += rng.normal(0, 12, size=img.shape)
    return np.clip(img, 0, 255).astype(np.uint8)


# ------------------------------------------------------------------------------------------------ DINOv2 fusion tail (f-2)
def dinov2_fusion_tail(layer_features, proj_sd, num_regions: int = 36):
    """models/feature_extractors/dinov2_multilayer.py:342-381 (fusion_strategy="concat") for patch features WITHOUT the CLS
    token: L x [1, P, hidden] -> [num_regions, out].  proj_sd holds projection.{0,1,3}.{weight,bias}."""
    fused = torch.cat(list(layer_features), dim=-1)
    num_patches, dim = fused.shape[1], fused.shape[-1]
    g, t = int(num_patches ** 0.5), int(num_regions ** 0.5)
    grid = fused.permute(0, 2, 1).reshape(1, dim, g, g)
    resized = F.interpolate(grid, size=(t, t), mode="bilinear", align_corners=False)
    flat = resized.permute(0, 2, 3, 1).reshape(-1, dim)
    y = F.linear(flat, proj_sd["projection.0.weight"], proj_sd["projection.0.bias"])
    y = F.layer_norm(y, (y.shape[-1],), proj_sd["projection.1.weight"], proj_sd["projection.1.bias"], 1e-5)
    y = F.gelu(y)
    return F.linear(y, proj_sd["projection.3.weight"], proj_sd["projection.3.bias"])


def grid_spatial(num_regions: int) -> np.ndarray:
    """dinov2_multilayer.py:383-403."""
    g = int(num_regions ** 0.5)
    out = np.zeros((num_regions, 5), np.float32)
    for i in range(g):
        for j in range(g):
            x1, y1, x2, y2 = j / g, i / g, (j + 1) / g, (i + 1) / g
            out[i * g + j] = np.array([x1, y1, x2, y2, (x2 - x1) * (y2 - y1)], dtype=np.float32)
    return out


def seeded_fusion_inputs(seed: int = 21, layers: int = 4, grid: int = 37, hidden: int = 1024, out_dim: int = 2048):
    """Seeded stand-ins for the ViT layer outputs and the projection parameters (no hub checkpoint offline)."""
    g = torch.Generator().manual_seed(seed)
    feats = [torch.randn(1, grid * grid, hidden, generator=g) for _ in range(layers)]
    k = layers * hidden
    sd = {"projection.0.weight": (torch.rand(out_dim, k, generator=g) * 2 - 1) * math.sqrt(6.0 / (k + out_dim)),
          "projection.0.bias": (torch.rand(out_dim, generator=g) - 0.5) * 0.1,
          "projection.1.weight": 0.75 + 0.5 * torch.rand(out_dim, generator=g),
          "projection.1.bias": (torch.rand(out_dim, generator=g) - 0.5) * 0.2,
          "projection.3.weight": (torch.rand(out_dim, out_dim, generator=g) * 2 - 1) * math.sqrt(6.0 / (2 * out_dim)),
          "projection.3.bias": (torch.rand(out_dim, generator=g) - 0.5) * 0.1}
    return feats, sd


# ------------------------------------------------------------------------------------------------ grid extractor (f-4)
def grid_backbone_state(sd: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
    """The same weights under the key names of ``ResNetFeatureExtractor.backbone`` (models/feature_extractors/resnet.py:33:
    ``nn.Sequential(*list(resnet.children())[:-2])`` -> "0." conv1, "1." bn1, "4.".."7." layer1..layer4)."""
    return {("7." + k[4:] if k.startswith("top.") else k[5:]): v for k, v in sd.items()}


def vg_backbone_state(sd: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
    """The same weights under the key names of ``VGResNet101Backbone`` (models/feature_extractors/resnet_vg.py:29-54)."""
    return {("RCNN_top." + k[4:] if k.startswith("top.") else "RCNN_base." + k[5:]): v for k, v in sd.items()}


def grid_features(sd, img: torch.Tensor, num_regions: int = 36, output_dim: int = 2048) -> np.ndarray:
    """resnet.py:51-76 (= resnet_vg.py:205-240 with ResNet-101 weights) on a preprocessed image [1,3,H,W]: whole trunk, adaptive average pool to a
    sqrt(num_regions) grid, rows in raster order, zero-padded / truncated to ``output_dim``.  ``sd`` uses the RoI backbone's
    key names (``seeded_backbone_state``)."""
    g = int(num_regions ** 0.5)
    with torch.no_grad():
        x = _layer(sd, "top", 3, 2, forward_base(sd, img))
        x = F.adaptive_avg_pool2d(x, (g, g))
        f = x.view(1, x.shape[1], -1).permute(0, 2, 1).squeeze(0)
    if f.shape[-1] < output_dim:
        f = torch.cat([f, torch.zeros(f.shape[0], output_dim - f.shape[-1])], dim=-1)
    return f[:, :output_dim].numpy()


# ------------------------------------------------------------------------------------------------ VG Faster R-CNN extractor (f-4)
VG_CLASSES = 1601


def seeded_vg_heads(seed: int = 11) -> Dict[str, torch.Tensor]:
    """Seeded stand-ins for the Visual Genome checkpoint's ``RCNN_cls_score`` / ``RCNN_bbox_pred`` heads
    (fasterrcnn_vg.py:75-76); the class weights are wide enough that proposal scores are well separated."""
    g = torch.Generator().manual_seed(seed)
    return {"RCNN_cls_score.weight": torch.randn(VG_CLASSES, 2048, generator=g) * 0.05,
            "RCNN_cls_score.bias": (torch.rand(VG_CLASSES, generator=g) - 0.5) * 0.1,
            "RCNN_bbox_pred.weight": torch.randn(VG_CLASSES * 4, 2048, generator=g) * 0.001,
            "RCNN_bbox_pred.bias": torch.zeros(VG_CLASSES * 4)}


def vg_grid_candidates(img_h: int, img_w: int, num_proposals: int = 100) -> np.ndarray:
    """fasterrcnn_vg.py:283-343: 5 scales x 3 aspect ratios, stride half a box, at most 2 * num_proposals candidates (the
    inner ``break`` fires AFTER the append that reaches the limit and the y advance), grid cells appended when fewer than
    num_proposals came out; Python-double accumulation, one rounding to fp32."""
    rows: List[List[float]] = []
    limit = num_proposals * 2
    for scale in (0.2, 0.3, 0.4, 0.5, 0.7):
        for ar in (0.5, 1.0, 2.0):
            box_w = img_w * scale
            box_h = box_w / ar
            box_h = min(box_h, img_h * 0.9)
            box_w = min(box_w, img_w * 0.9)
            stride_x, stride_y = max(box_w * 0.5, 1), max(box_h * 0.5, 1)
            x = 0
            while x + box_w <= img_w:
                y = 0
                while y + box_h <= img_h:
                    rows.append([x, y, x + box_w, y + box_h])
                    y += stride_y
                    if len(rows) >= limit:
                        break
                x += stride_x
                if len(rows) >= limit:
                    break
            if len(rows) >= limit:
                break
        if len(rows) >= limit:
            break
    if len(rows) < num_proposals:
        grid = int((num_proposals - len(rows)) ** 0.5) + 1
        cell_w, cell_h = img_w / grid, img_h / grid
        for i in range(grid):
            for j in range(grid):
                rows.append([j * cell_w, i * cell_h, min((j + 1) * cell_w, img_w), min((i + 1) * cell_h, img_h)])
    return np.array(rows[:limit], dtype=np.float32).reshape(-1, 4)


def vg_select(boxes: np.ndarray, scores: np.ndarray, num_regions: int, nms_thr: float) -> np.ndarray:
    """fasterrcnn_vg.py:367-411 (``_select_top_regions`` + ``_pad_regions``): indices into ``boxes`` of the chosen regions.
    ``nms`` returns the survivors in descending score order, so the reference's ``topk`` over them is their first
    ``num_regions`` entries whenever the scores are distinct; among exactly tied scores torch.topk's order is unspecified and
    this restatement keeps the stable (index) order."""
    n = len(boxes)
    if n == 0:
        return np.zeros(0, dtype=np.int64)
    idx = np.arange(n, dtype=np.int64)
    if n > num_regions:
        idx = nms(boxes, scores, nms_thr)[:num_regions]
    if len(idx) < num_regions:
        idx = np.concatenate([idx, np.full(num_regions - len(idx), idx[-1], dtype=np.int64)])
    return idx[:num_regions]


def vg_scores(sd, heads, fmap: torch.Tensor, boxes: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
    """fasterrcnn_vg.py:345-365 (``_score_proposals``): RoIPool-14 -> layer4 -> mean -> class scores -> max over the 1600
    object classes (background column 0 excluded).  Returns (scores [n], top features [n, 2048])."""
    rois = np.concatenate([np.zeros((len(boxes), 1), np.float32), boxes], axis=1)
    with torch.no_grad():
        pooled = roi_pool(fmap.numpy(), rois, 14, 1.0 / 16.0)
        top = forward_top(sd, torch.from_numpy(pooled))
        cls = F.linear(top, heads["RCNN_cls_score.weight"], heads["RCNN_cls_score.bias"])
    return cls[:, 1:].max(dim=1)[0].numpy(), top.numpy()


def vg_extract_features(sd, heads, img: torch.Tensor, num_regions: int = 36, nms_thr: float = 0.3,
                        has_vg_weights: bool = True, scores: np.ndarray = None):
    """fasterrcnn_vg.py:252-281 on a preprocessed image [1,3,H,W]: (features [N,2048], spatial [N,5], boxes [N,4], candidate
    scores).  ``sd`` uses the RoI backbone's key names with ResNet-101 block counts; ``scores`` overrides the candidate scores
    (selection parity under given scores)."""
    h, w = img.shape[2], img.shape[3]
    with torch.no_grad():
        fmap = forward_base(sd, img)
    cands = vg_grid_candidates(h, w)
    if scores is None:
        scores = vg_scores(sd, heads, fmap, cands)[0] if has_vg_weights else np.ones(len(cands), np.float32)
    idx = vg_select(cands, scores, num_regions, nms_thr)
    boxes = cands[idx]
    rois = np.concatenate([np.zeros((len(boxes), 1), np.float32), boxes], axis=1)
    with torch.no_grad():
        feats = forward_top(sd, torch.from_numpy(roi_pool(fmap.numpy(), rois, 14, 1.0 / 16.0))).numpy()
    return feats, normalize_boxes(boxes, w, h), boxes, scores


# ------------------------------------------------------------------------------------------------ VG Faster R-CNN with RPN proposals (f-4)
RPN_SCALES, RPN_RATIOS, RPN_STRIDE = (4, 8, 16, 32), (0.5, 1.0, 2.0), 16


def seeded_rpn_state(seed: int = 13) -> Dict[str, torch.Tensor]:
    """Seeded stand-ins for the Visual Genome checkpoint's ``RCNN_rpn`` tensors (fasterrcnn_vg_rpn.py:34-56); the box deltas are
    kept small so that proposals stay anchor-like and many of them pass the min-size filter."""
    g = torch.Generator().manual_seed(seed)
    return {"RCNN_rpn.RPN_Conv.weight": torch.randn(512, 1024, 3, 3, generator=g) * math.sqrt(2.0 / (1024 * 9)),
            "RCNN_rpn.RPN_Conv.bias": (torch.rand(512, generator=g) - 0.5) * 0.1,
            "RCNN_rpn.RPN_cls_score.weight": torch.randn(24, 512, 1, 1, generator=g) * 0.004,
            "RCNN_rpn.RPN_cls_score.bias": (torch.rand(24, generator=g) - 0.5) * 0.1,
            "RCNN_rpn.RPN_bbox_pred.weight": torch.randn(48, 512, 1, 1, generator=g) * 0.01,
            "RCNN_rpn.RPN_bbox_pred.bias": (torch.rand(48, generator=g) - 0.5) * 0.05}


def rpn_base_anchors() -> np.ndarray:
    """fasterrcnn_vg_rpn.py:110-118: 4 scales x 3 ratios around (0, 0), Python-double arithmetic, one rounding to fp32."""
    rows = []
    for scale in RPN_SCALES:
        for ratio in RPN_RATIOS:
            h = scale * RPN_STRIDE * (ratio ** 0.5)
            w = scale * RPN_STRIDE / (ratio ** 0.5)
            rows.append([-w / 2, -h / 2, w / 2, h / 2])
    return np.array(rows, dtype=np.float32)


def rpn_heads(rpn_sd, fmap: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """:78-92: 3x3 conv + ReLU, then the two 1x1 heads, re-laid as [H*W*12, 2] objectness logits and [H*W*12, 4] deltas."""
    x = F.relu(F.conv2d(fmap, rpn_sd["RCNN_rpn.RPN_Conv.weight"], rpn_sd["RCNN_rpn.RPN_Conv.bias"], padding=1))
    cls = F.conv2d(x, rpn_sd["RCNN_rpn.RPN_cls_score.weight"], rpn_sd["RCNN_rpn.RPN_cls_score.bias"])
    box = F.conv2d(x, rpn_sd["RCNN_rpn.RPN_bbox_pred.weight"], rpn_sd["RCNN_rpn.RPN_bbox_pred.bias"])
    return cls.permute(0, 2, 3, 1).reshape(-1, 2), box.permute(0, 2, 3, 1).reshape(-1, 4)


def rpn_decode(cls: np.ndarray, box: np.ndarray, fh: int, fw: int, img_h: int, img_w: int) -> Tuple[np.ndarray, np.ndarray]:
    """:85-104, 106-174: foreground probability (softmax over the pair), anchors (base + cell centre), deltas applied in fp32 with
    one rounding per operation, clipped to the picture.  Returns (proposals [A,4], scores [A]), A = fh*fw*12, anchor fastest."""
    f32 = np.float32
    cls, box = cls.astype(f32), box.astype(f32)
    m = np.maximum(cls[:, 0], cls[:, 1])
    e0, e1 = np.exp(cls[:, 0] - m, dtype=f32), np.exp(cls[:, 1] - m, dtype=f32)
    scores = (e1 / (e0 + e1)).astype(f32)
    ys, xs = np.meshgrid(np.arange(fh) * RPN_STRIDE + RPN_STRIDE // 2, np.arange(fw) * RPN_STRIDE + RPN_STRIDE // 2, indexing="ij")
    shifts = np.stack([xs, ys, xs, ys], axis=-1).reshape(-1, 1, 4).astype(f32)
    anchors = (rpn_base_anchors()[None, :, :] + shifts).reshape(-1, 4).astype(f32)
    widths, heights = anchors[:, 2] - anchors[:, 0], anchors[:, 3] - anchors[:, 1]
    ctr_x, ctr_y = anchors[:, 0] + f32(0.5) * widths, anchors[:, 1] + f32(0.5) * heights
    dw, dh = np.minimum(box[:, 2], f32(4.0)), np.minimum(box[:, 3], f32(4.0))
    pcx, pcy = box[:, 0] * widths + ctr_x, box[:, 1] * heights + ctr_y
    pw, ph = np.exp(dw, dtype=f32) * widths, np.exp(dh, dtype=f32) * heights
    props = np.stack([pcx - f32(0.5) * pw, pcy - f32(0.5) * ph, pcx + f32(0.5) * pw, pcy + f32(0.5) * ph], axis=-1).astype(f32)
    props[:, 0::2] = np.clip(props[:, 0::2], f32(0), f32(img_w))
    props[:, 1::2] = np.clip(props[:, 1::2], f32(0), f32(img_h))
    return props, scores


def rpn_filter(props: np.ndarray, scores: np.ndarray, min_size: float = 16, pre_nms: int = 6000, post_nms: int = 300,
               nms_thr: float = 0.7) -> np.ndarray:
    """:442-469 (``_filter_proposals``): indices into ``props`` of the proposals that survive, in descending score order.  The
    reference's ``topk`` among exactly tied scores is unspecified; this restatement keeps the stable order."""
    w, h = props[:, 2] - props[:, 0], props[:, 3] - props[:, 1]
    idx = np.nonzero((w >= np.float32(min_size)) & (h >= np.float32(min_size)))[0]
    if len(idx) == 0:
        return idx
    order = idx[np.argsort(-scores[idx], kind="stable")][:pre_nms]
    keep = nms(props[order], scores[order], nms_thr)[:post_nms]
    return order[keep]


def rpn_resize(w: int, h: int, target: int = 600, max_size: int = 1000) -> Tuple[int, int, float]:
    """:373-385 (``_resize_image``): (new_w, new_h, scale)."""
    scale = target / min(w, h)
    if max(w, h) * scale > max_size:
        scale = max_size / max(w, h)
    return int(w * scale), int(h * scale), scale


def rpn_pad_grid(num_needed: int, img_w: int, img_h: int) -> np.ndarray:
    """:497-516: the grid cells appended when fewer than num_regions proposals survive."""
    g = int(num_needed ** 0.5) + 1
    cw, ch = img_w / g, img_h / g
    rows = []
    for i in range(g):
        for j in range(g):
            if len(rows) >= num_needed:
                break
            rows.append([j * cw, i * ch, min((j + 1) * cw, img_w), min((i + 1) * ch, img_h)])
        if len(rows) >= num_needed:
            break
    return np.array(rows, dtype=np.float32).reshape(-1, 4)


def vg_rpn_extract_features(sd, heads, rpn_sd, img: torch.Tensor, scale: float, orig_w: int, orig_h: int, num_regions: int = 36,
                            region_scores: np.ndarray = None, kept: np.ndarray = None):
    """fasterrcnn_vg_rpn.py:387-440 on a preprocessed picture [1,3,H,W]: (features [N,2048], spatial [N,5], boxes [N,4] in resized
    pixels, kept proposal boxes, their region scores).  ``kept`` / ``region_scores`` override the RPN survivors / their class scores
    (selection parity under given inputs)."""
    h, w = img.shape[2], img.shape[3]
    with torch.no_grad():
        fmap = forward_base(sd, img)
        if kept is None:
            cls, box = rpn_heads(rpn_sd, fmap)
            props, scores = rpn_decode(cls.numpy(), box.numpy(), fmap.shape[2], fmap.shape[3], h, w)
            kept = props[rpn_filter(props, scores)]
        if region_scores is None or len(kept) < num_regions:
            s, top = vg_scores(sd, heads, fmap, kept) if len(kept) else (np.zeros(0, np.float32), np.zeros((0, 2048), np.float32))
            region_scores = s if region_scores is None else region_scores
        else:
            top = None
        boxes = kept
        if len(kept) > num_regions:
            idx = np.argsort(-region_scores.astype(np.float32), kind="stable")[:num_regions]
            boxes = kept[idx]
        elif len(kept) < num_regions:
            boxes = np.concatenate([kept, rpn_pad_grid(num_regions - len(kept), w, h)], axis=0)[:num_regions]
        rois = np.concatenate([np.zeros((len(boxes), 1), np.float32), boxes.astype(np.float32)], axis=1)
        feats = forward_top(sd, torch.from_numpy(roi_pool(fmap.numpy(), rois, 14, 1.0 / 16.0))).numpy()
    spatial = normalize_boxes((boxes.astype(np.float32) / np.float32(scale)).astype(np.float32), orig_w, orig_h)
    return feats, spatial, boxes, kept, region_scores
