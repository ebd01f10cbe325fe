"""Generate tests/golden/resnet_grid.npz by running the UNMODIFIED reference grid extractor
(models/feature_extractors/resnet.py, ``ResNetFeatureExtractor``) in the authoring container.  As for the RoI stage the
ImageNet checkpoint cannot be downloaded: ``resnet152`` is rebound to a weight-less constructor and the seeded weights of
``oracle.roi_oracle.seeded_backbone_state`` are loaded (strict) under the extractor's own key names."""
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
    import multimodalclassification.models.feature_extractors.resnet as ref
    ref.resnet152 = lambda weights=None, **kw: torchvision.models.resnet152(weights=None, **kw)
    torch.set_num_threads(os.cpu_count() or 1)
    out = {"image_u8": ro.synthetic_image(7)}
    for n in (36, 49, 9):
        ext = ref.ResNetFeatureExtractor(num_regions=n, device="cpu")
        print(n, ext.backbone.load_state_dict(ro.grid_backbone_state(ro.seeded_backbone_state(0)), strict=True))
        feats, spatial = ext.extract_features(Image.fromarray(out["image_u8"]))
        out[f"features_{n}"] = feats.numpy().astype(np.float32) if n == 36 else feats.numpy()[:, ::16].astype(np.float32)
        out[f"spatial_{n}"] = spatial.numpy().astype(np.float32)
    ext = ref.ResNetFeatureExtractor(output_dim=2304, num_regions=36, device="cpu")
    ext.backbone.load_state_dict(ro.grid_backbone_state(ro.seeded_backbone_state(0)), strict=True)
    padded, _ = ext.extract_features(Image.fromarray(out["image_u8"]))
    assert padded.shape == (36, 2304) and torch.equal(padded[:, :2048], torch.from_numpy(out["features_36"])) \
        and padded[:, 2048:].abs().max() == 0
    # the Visual Genome ResNet-101 variant (resnet_vg.py): no checkpoint on disk -> its torchvision-weights branch, then seeded
    import multimodalclassification.models.feature_extractors.resnet_vg as ref_vg
    ref_vg.resnet101 = lambda weights=None, **kw: torchvision.models.resnet101(weights=None, **kw)
    vg = ref_vg.ResNetVGExtractor(weights_path="/nonexistent.pth", device="cpu")
    print("vg", vg.backbone.load_state_dict(ro.vg_backbone_state(ro.seeded_backbone_state(1, (3, 4, 23, 3))), strict=True))
    feats, spatial = vg.extract_features(Image.fromarray(out["image_u8"]))
    out["vg_features_36"] = feats.numpy()[:, ::4].astype(np.float32)
    out["vg_spatial_36"] = spatial.numpy().astype(np.float32)
    # ... and its checkpoint loader on a checkpoint with the VG file's key spelling, foreign keys and a wrong shape
    ck = {("RCNN_top.0." + k[9:] if k.startswith("RCNN_top.") else k): v for k, v in vg.backbone.state_dict().items()}
    ck["RCNN_rpn.RPN_Conv.weight"] = torch.zeros(4)
    ck["RCNN_cls_score.weight"] = torch.zeros(4)
    ck["RCNN_base.0.weight"] = torch.zeros(64, 3, 3, 3)
    ck["RCNN_base.9.weight"] = torch.zeros(1)
    torch.save({"model": ck}, "/tmp/_vg_ckpt.pth")
    fresh = ref_vg.VGResNet101Backbone()
    stats = ref_vg.load_vg_backbone_weights(fresh, "/tmp/_vg_ckpt.pth")
    out["vg_loader_stats"] = np.array([stats["loaded"], stats["total_model"], stats["skipped"]])
    out["vg_loader_skipped_keys"] = np.array(stats["skipped_keys"])
    path = os.path.join(ROOT, "tests", "golden", "resnet_grid.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path) // 1024, "KiB; |f| mean %.4f max %.4f" % (np.abs(out["features_36"]).mean(),
                                                                                      np.abs(out["features_36"]).max()))


if __name__ == "__main__":
    main()
