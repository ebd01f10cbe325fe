"""Generate tests/golden/fasterrcnn_vg.npz by running the UNMODIFIED reference extractor
(models/feature_extractors/fasterrcnn_vg.py, ``FasterRCNNVGExtractor``) in the authoring container.

Neither the ImageNet ResNet-101 weights nor the Visual Genome checkpoint can be downloaded here: ``resnet101`` is rebound to
a weight-less constructor and a SEEDED checkpoint with the Visual Genome file's key spelling (``RCNN_base.*``,
``RCNN_top.0.*``, ``RCNN_cls_score``, ``RCNN_bbox_pred``, plus foreign ``RCNN_rpn`` keys) is written to /tmp and loaded by the
reference's own ``load_vg_weights``, which also switches on its classifier-scored proposal branch."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference/src")
from oracle import roi_oracle as ro  # noqa: E402

CKPT = "/tmp/_vg_frcnn_ckpt.pth"


def vg_checkpoint():
    sd = ro.vg_backbone_state(ro.seeded_backbone_state(1, (3, 4, 23, 3)))
    ck = {("RCNN_top.0." + k[9:] if k.startswith("RCNN_top.") else k): v for k, v in sd.items()}
    ck.update(ro.seeded_vg_heads(11))
    ck["RCNN_rpn.RPN_Conv.weight"] = torch.zeros(512, 1024, 3, 3)      # present in the real file, ignored by this extractor
    ck["RCNN_base.0.bias"] = torch.zeros(64)                            # no such key in the model
    return ck


def main():
    import torchvision
    from PIL import Image
    import multimodalclassification.models.feature_extractors.fasterrcnn_vg as ref
    ref.resnet101 = lambda weights=None, **kw: torchvision.models.resnet101(weights=None, **kw)
    torch.set_num_threads(os.cpu_count() or 1)
    ck = vg_checkpoint()
    torch.save({"model": ck}, CKPT)
    out = {"image_u8": ro.synthetic_image(7)}
    pic = Image.fromarray(out["image_u8"])

    ext = ref.FasterRCNNVGExtractor(weights_path=CKPT, device="cpu")
    assert ext.has_vg_weights
    out["loaded_count"] = np.array(ref.load_vg_weights(ref.VGFasterRCNN(), CKPT))
    out["model_keys"] = np.array(sorted(ext.model.state_dict().keys()))
    img = ext.transform(pic).unsqueeze(0)
    base = ext.model(img)
    cands, scores = ext._generate_proposals(base, img.shape[2], img.shape[3])
    boxes, sel_scores = ext._select_top_regions(cands, scores)
    feats, spatial = ext.extract_features(pic)
    out.update(candidates=cands.numpy(), scores=scores.numpy(), boxes=boxes.numpy(), sel_scores=sel_scores.numpy(),
               features=feats.numpy().astype(np.float32), spatial=spatial.numpy().astype(np.float32),
               base_probe=base[0, ::64, ::6, ::9].numpy().astype(np.float32))
    assert torch.equal(spatial[:, :4], ext._normalize_boxes(boxes, img.shape[3], img.shape[2])[:, :4])
    # fewer regions than NMS survivors / more regions than candidates survive NMS (padding branch)
    for n in (10, 100):
        ext.num_regions = n
        b, s = ext._select_top_regions(cands, scores)
        out[f"boxes_{n}"] = b.numpy()
    ext.num_regions = 36

    # no checkpoint on disk: every candidate scores 1.0 (torchvision-weights branch, fasterrcnn_vg.py:340-343)
    plain = ref.FasterRCNNVGExtractor(weights_path="/nonexistent.pth", device="cpu")
    assert not plain.has_vg_weights
    plain.model.load_state_dict({k: v for k, v in ext.model.state_dict().items()}, strict=True)
    pf, ps = plain.extract_features(pic)
    c2, s2 = plain._generate_proposals(base, img.shape[2], img.shape[3])
    b2, _ = plain._select_top_regions(c2, s2)
    out.update(plain_boxes=b2.numpy(), plain_features=pf.numpy()[:, ::8].astype(np.float32), plain_spatial=ps.numpy().astype(np.float32))
    # proposals at other picture sizes (pure host arithmetic)
    for h, w in ((600, 1000), (224, 224), (480, 640), (97, 1000)):
        c, _ = plain._generate_proposals(base, h, w)
        out[f"cands_{h}x{w}"] = c.numpy()
    path = os.path.join(ROOT, "tests", "golden", "fasterrcnn_vg.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path) // 1024, "KiB; candidates", cands.shape, "score range",
          float(scores.min()), float(scores.max()), "|f| max", float(feats.abs().max()))
    srt = np.sort(scores.numpy())[::-1]
    print("smallest gap among the candidate scores:", float(np.min(srt[:-1] - srt[1:])), "relative", float(np.min(srt[:-1] - srt[1:]) / srt[0]))


if __name__ == "__main__":
    main()
