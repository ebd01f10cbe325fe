"""Generate tests/golden/dinov2_tail.npz by running the reference's OWN fusion-tail code
(models/feature_extractors/dinov2_multilayer.py:342-403).  The ViT backbone comes from torch.hub (no network here), so the
extractor object is created without running its constructor and given seeded layer features and projection parameters;
everything from `extract_features` line 342 on is the unmodified reference code."""
import os
import sys

import numpy as np
import torch
import torch.nn as nn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference/src")
from oracle import roi_oracle as ro  # noqa: E402


def main():
    from PIL import Image
    from multimodalclassification.models.feature_extractors.dinov2_multilayer import DINOv2MultiLayerExtractor as Ref
    feats, sd = ro.seeded_fusion_inputs()
    ext = Ref.__new__(Ref)
    nn.Module.__init__(ext)
    ext.output_dim, ext.num_regions, ext.device, ext.fusion_strategy = 2048, 36, "cpu", "concat"
    ext.projection = nn.Sequential(nn.Linear(4096, 2048), nn.LayerNorm(2048), nn.GELU(), nn.Linear(2048, 2048))   # :250-255
    ext.load_state_dict(sd, strict=True)
    ext.transform = lambda img: torch.zeros(3, 8, 8)
    ext._extract_multilayer_features = lambda t: feats
    with torch.no_grad():
        out, spatial = ext.extract_features(Image.new("RGB", (8, 8)))
    path = os.path.join(ROOT, "tests", "golden", "dinov2_tail.npz")
    np.savez_compressed(path, projected=out.numpy().astype(np.float32), spatial=spatial.numpy().astype(np.float32))
    print("wrote", path, os.path.getsize(path) // 1024, "KiB; |out| mean %.4f max %.4f" % (out.abs().mean(), out.abs().max()))


if __name__ == "__main__":
    main()
