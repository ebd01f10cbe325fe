"""Generate tests/golden/fasterrcnn_vg_rpn.npz by running the UNMODIFIED reference extractor
(models/feature_extractors/fasterrcnn_vg_rpn.py, ``FasterRCNNVGRPNExtractor``) in the authoring container on a SEEDED checkpoint
with the Visual Genome file's key spelling (backbone, ``RCNN_rpn``, ``RCNN_cls_score``; see oracle/make_golden_vg.py for why)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference/src")
from oracle import roi_oracle as ro  # noqa: E402

CKPT = "/tmp/_vg_rpn_ckpt.pth"


def vg_rpn_checkpoint():
    sd = ro.vg_backbone_state(ro.seeded_backbone_state(1, (3, 4, 23, 3)))
    ck = {("RCNN_top.0." + k[9:] if k.startswith("RCNN_top.") else k): v for k, v in sd.items()}
    ck.update(ro.seeded_vg_heads(11))
    ck.update(ro.seeded_rpn_state(13))
    ck["RCNN_base.0.bias"] = torch.zeros(64)                            # no such key in the model
    return ck


def main():
    import torchvision
    from PIL import Image
    import multimodalclassification.models.feature_extractors.fasterrcnn_vg_rpn as ref
    ref.resnet101 = lambda weights=None, **kw: torchvision.models.resnet101(weights=None, **kw)
    torch.set_num_threads(os.cpu_count() or 1)
    torch.save({"model": vg_rpn_checkpoint()}, CKPT)
    out = {"image_u8": ro.synthetic_image(7)}
    pic = Image.fromarray(out["image_u8"])
    ext = ref.FasterRCNNVGRPNExtractor(weights_path=CKPT, device="cpu")
    assert ext.has_vg_weights
    stats = ref.load_vg_checkpoint(ref.VGFasterRCNNWithRPN(), CKPT)
    out["loader_stats"] = np.array([stats["loaded"], stats["total"], stats["skipped"]])
    out["model_keys"] = np.array(sorted(ext.model.state_dict().keys()))
    resized, scale = ext._resize_image(pic)
    out["resized_size"] = np.array(resized.size)
    out["scale"] = np.array(scale, dtype=np.float64)
    img = ext.transform(resized).unsqueeze(0)
    h, w = img.shape[2], img.shape[3]
    with torch.no_grad():
        base = ext.model.get_base_features(img)
        props, scores = ext.model.get_proposals(base, (h, w))
        boxes, kept_scores = ext._filter_proposals(props, scores, (h, w))
        roi = ext._extract_roi_features(base, boxes)
        region = ext.model.get_class_scores(roi)[:, 1:].max(dim=1)[0]
        feats, spatial = ext.extract_features(pic)
    out.update(fmap_hw=np.array(base.shape[2:]), proposals_full=props.numpy(), scores_full=scores.numpy(), kept_boxes=boxes.numpy(), kept_scores=kept_scores.numpy(), region_scores=region.numpy(),
               features=feats.numpy().astype(np.float32), spatial=spatial.numpy().astype(np.float32))
    # the RPN head outputs themselves (decode parity without the trunk's rounding)
    cls, box = ro.rpn_heads(ro.seeded_rpn_state(13), base)
    out.update(rpn_cls=cls.numpy(), rpn_box=box.numpy())
    # fewer proposals than regions: the grid padding branch (:490-535)
    ext.num_regions = 320
    f2, s2 = ext.extract_features(pic)
    out.update(padded_spatial=s2.numpy().astype(np.float32), padded_features=f2.numpy()[::8, ::16].astype(np.float32))
    for size in ((128, 96), (96, 128), (400, 100), (640, 480)):
        r, s = ext._resize_image(Image.new("RGB", size))
        out[f"resize_{size[0]}x{size[1]}"] = np.array([r.size[0], r.size[1], s], dtype=np.float64)
    path = os.path.join(ROOT, "tests", "golden", "fasterrcnn_vg_rpn.npz")
    np.savez_compressed(path, **out)
    srt = np.sort(region.numpy())[::-1]
    print("wrote", path, os.path.getsize(path) // 1024, "KiB; picture", (h, w), "anchors", props.shape[0], "kept", boxes.shape[0],
          "valid", int(((props[:, 2] - props[:, 0] >= 16) & (props[:, 3] - props[:, 1] >= 16)).sum()),
          "score range", float(scores.min()), float(scores.max()), "region score gap min", float(np.min(srt[:-1] - srt[1:])))


if __name__ == "__main__":
    main()
