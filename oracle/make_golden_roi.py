"""Generate tests/golden/roi_stage.npz by running the UNMODIFIED reference RoI extractor (authoring container only:
/root/reference is not on the GPU box).  The only substitution is the ImageNet checkpoint, which cannot be downloaded here:
``resnet152`` is rebound to a weight-less constructor and ``oracle.roi_oracle.seeded_backbone_state`` is loaded instead.

    python oracle/make_golden_roi.py
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference/src")

from oracle import roi_oracle as ro  # noqa: E402


def main():
    import torchvision
    from PIL import Image
    from torchvision.ops import RoIPool, nms
    import multimodalclassification.models.feature_extractors.resnet152_roi as ref

    ref.resnet152 = lambda weights=None, **kw: torchvision.models.resnet152(weights=None, **kw)
    torch.manual_seed(0)
    torch.set_num_threads(os.cpu_count() or 1)
    ext = ref.ResNet152ROIExtractor(device="cpu")
    missing = ext.backbone.load_state_dict(ro.seeded_backbone_state(0), strict=True)
    print("loaded seeded backbone:", missing)
    out = {}

    # (1) whole stage on a seeded picture
    pic = ro.synthetic_image(7)
    feats, spatial = ext.extract_features(Image.fromarray(pic))
    out["image_u8"] = pic
    out["features"] = feats.numpy().astype(np.float32)
    out["spatial"] = spatial.numpy().astype(np.float32)
    img = ext.transform(Image.fromarray(pic)).unsqueeze(0)
    with torch.no_grad():
        fmap = ext.backbone.forward_base(img)
    out["fmap_digest"] = np.array([fmap.abs().mean().item(), fmap.abs().max().item(), fmap.std().item()], np.float64)
    out["fmap_probe"] = fmap[0, ::64, ::6, ::6].numpy().astype(np.float32)

    # (2) proposals at several image sizes, multi-scale and grid
    sizes = [(600, 600), (448, 448), (224, 224), (480, 640), (333, 500)]
    out["proposal_sizes"] = np.array(sizes, np.int64)
    for h, w in sizes:
        ext.use_multi_scale = True
        out[f"boxes_ms_{h}x{w}"] = ext._generate_proposals(h, w).numpy()
        out[f"spatial_ms_{h}x{w}"] = ext._normalize_boxes(ext._generate_proposals(h, w), w, h).numpy()
        ext.use_multi_scale = False
        out[f"boxes_grid_{h}x{w}"] = ext._generate_proposals(h, w).numpy()
    ext.use_multi_scale = True

    # (3) NMS: the reference's own candidate set + a random one with exact ties
    cands = torch.from_numpy(ro.multi_scale_candidates(600, 600))
    sc = torch.from_numpy(ro.area_scores(cands.numpy(), 600, 600))
    out["nms_cands"] = cands.numpy()
    out["nms_scores"] = sc.numpy()
    out["nms_keep"] = nms(cands, sc, 0.5).numpy()
    g = torch.Generator().manual_seed(3)
    xy = torch.rand(400, 2, generator=g) * 300
    wh = torch.rand(400, 2, generator=g) * 120 + 4
    rb = torch.cat([xy, xy + wh], dim=1)
    rs = (torch.randint(0, 12, (400,), generator=g).float() / 12.0)      # many exact ties
    out["nms_rand_boxes"], out["nms_rand_scores"] = rb.numpy(), rs.numpy()
    for thr in (0.3, 0.5, 0.7):
        out[f"nms_rand_keep_{int(thr * 10)}"] = nms(rb, rs, thr).numpy()

    # (4) RoIPool: bf16-representable map, boxes with .5 roundings, boxes past the border, tiny boxes
    fm = (torch.randn(2, 24, 38, 38, generator=g)).to(torch.bfloat16).float()
    rois = []
    for i in range(60):
        b = i % 2
        x1, y1 = torch.rand(2, generator=g).tolist()
        x1, y1 = x1 * 560 - 20, y1 * 560 - 20
        w_, h_ = (torch.rand(2, generator=g) * 300 + 1).tolist()
        rois.append([b, x1, y1, x1 + w_, y1 + h_])
    rois += [[0, 8.0, 8.0, 24.0, 24.0], [1, 0.0, 0.0, 600.0, 600.0], [0, 599.0, 599.0, 640.0, 640.0], [1, 100.0, 100.0, 100.0, 100.0],
             [0, 40.0, 56.0, 72.0, 88.0], [0, -50.0, -50.0, -20.0, -20.0]]
    rois = torch.tensor(rois, dtype=torch.float32)
    out["roi_fmap"], out["roi_rois"] = fm.numpy(), rois.numpy()
    for p in (14, 7):
        out[f"roi_pool_{p}"] = RoIPool((p, p), 1 / 16)(fm, rois).numpy()
    from torchvision.ops import roi_align
    out["roi_align_7"] = roi_align(fm, rois, (7, 7), spatial_scale=1 / 16, sampling_ratio=2, aligned=False).numpy()

    path = os.path.join(ROOT, "tests", "golden", "roi_stage.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path) // 1024, "KiB")
    print("features: mean |x| %.4f  max %.4f" % (np.abs(out["features"]).mean(), np.abs(out["features"]).max()))


if __name__ == "__main__":
    main()
