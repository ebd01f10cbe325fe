"""Generate tests/golden/roi_stage_448_align.npz: BASELINE.json configs[2] (3x448x448 image, RoIAlign 7x7 over 36 boxes, C5
features feeding ViLBERT) produced by the UNMODIFIED reference extractor class with two attributes rebound to the values that
configuration names (authoring container only: /root/reference is not on the GPU box):

* ``transform``  -> Resize((448, 448)) instead of the class's 600 x 600 (resnet152_roi.py:126-133)
* ``roi_pool``   -> ``torchvision.ops.RoIAlign((7, 7), 1/16, sampling_ratio=2)``, the pooling op of the reference's detection
                    extractor (feature_extractors/fasterrcnn_resnet152.py:130-134) in place of ``RoIPool((14, 14), 1/16)``

Everything else (backbone split, proposal generation + NMS, forward_top, box normalisation) is the reference's own code.  The
second half chains the stage into the encoder the way pipelines/model_training/nodes.py:129-148, 195-202 does: features and
boxes of two pictures -> ``ViLBERTForClassification(tiny config, v_feature_size 2048)`` -> logits / loss.

    python oracle/make_golden_roi_align.py
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference/src")

from oracle import roi_oracle as ro  # noqa: E402
from oracle import vilbert_oracle as vo  # noqa: E402


def chain_config():
    """Tiny encoder whose visual input width is the RoI stage's 2048."""
    c = vo.tiny_config()
    c["v_feature_size"] = 2048
    return c


def main():
    import torchvision
    from PIL import Image
    from torchvision import transforms
    from torchvision.ops import RoIAlign
    import multimodalclassification.models.feature_extractors.resnet152_roi as ref
    from multimodalclassification.models.vilbert_facebook_arch import ViLBERTForClassification

    ref.resnet152 = lambda weights=None, **kw: torchvision.models.resnet152(weights=None, **kw)
    torch.manual_seed(0)
    torch.set_num_threads(os.cpu_count() or 1)
    ext = ref.ResNet152ROIExtractor(roi_size=7, device="cpu")
    print("loaded seeded backbone:", ext.backbone.load_state_dict(ro.seeded_backbone_state(0), strict=True))
    ext.transform = transforms.Compose([transforms.Resize((448, 448)), transforms.ToTensor(),
                                        transforms.Normalize(mean=[0.485, 0.456, 0.406], std=[0.229, 0.224, 0.225])])
    ext.roi_pool = RoIAlign(output_size=(7, 7), spatial_scale=1 / 16, sampling_ratio=2)
    out = {}
    pics = [ro.synthetic_image(11), ro.synthetic_image(12)]
    feats, spatial = [], []
    for pic in pics:
        f, s = ext.extract_features(Image.fromarray(pic))
        feats.append(f.numpy().astype(np.float32))
        spatial.append(s.numpy().astype(np.float32))
    out["images_u8"] = np.stack(pics)
    out["features"] = np.stack(feats)
    out["spatial"] = np.stack(spatial)

    # chained: the extractor's output is the encoder's visual input (nodes.py:129-148 builds the batch, :195-202 calls the model)
    cfg = chain_config()
    sd = vo.seeded_state_dict(cfg)
    model = ViLBERTForClassification(cfg, num_labels=2).eval()
    print(model.load_state_dict(sd, strict=True))
    g = torch.Generator().manual_seed(5)
    ids = torch.randint(1, cfg["vocab_size"], (2, 24), generator=g)
    mask = torch.ones(2, 24, dtype=torch.int64)
    mask[1, 17:] = 0
    labels = torch.tensor([1, 0])
    with torch.no_grad():
        o = model(input_ids=ids, attention_mask=mask, token_type_ids=torch.zeros_like(ids),
                  visual_features=torch.from_numpy(out["features"]), visual_attention_mask=torch.ones(2, 36, dtype=torch.int64),
                  spatial_locations=torch.from_numpy(out["spatial"]), labels=labels)
    out["chain_input_ids"], out["chain_attention_mask"], out["chain_labels"] = ids.numpy(), mask.numpy(), labels.numpy()
    out["chain_logits"], out["chain_loss"] = o["logits"].numpy(), o["loss"].numpy()
    path = os.path.join(ROOT, "tests", "golden", "roi_stage_448_align.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path) // 1024, "KiB; features mean |x| %.4f max %.4f; chain logits %s loss %.5f" %
          (np.abs(out["features"]).mean(), np.abs(out["features"]).max(), out["chain_logits"].round(4).tolist(), float(out["chain_loss"])))


if __name__ == "__main__":
    main()
