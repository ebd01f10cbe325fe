"""The RoI-stage oracle (oracle/roi_oracle.py) against fixtures produced by the reference itself
(oracle/make_golden_roi.py -> tests/golden/roi_stage.npz).  CPU only."""
import os

import numpy as np
import pytest

from oracle import roi_oracle as ro

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "roi_stage.npz"))
SIZES = [tuple(int(v) for v in hw) for hw in G["proposal_sizes"]]


@pytest.mark.parametrize("h,w", SIZES)
def test_proposals_bit_exact(h, w):
    assert np.array_equal(ro.proposals(36, h, w, True), G[f"boxes_ms_{h}x{w}"])
    assert np.array_equal(ro.proposals(36, h, w, False), G[f"boxes_grid_{h}x{w}"])
    assert np.array_equal(ro.normalize_boxes(G[f"boxes_ms_{h}x{w}"], w, h), G[f"spatial_ms_{h}x{w}"])


def test_nms_order_bit_exact():
    assert np.array_equal(ro.area_scores(G["nms_cands"], 600, 600), G["nms_scores"])
    assert np.array_equal(ro.nms(G["nms_cands"], G["nms_scores"], 0.5), G["nms_keep"])
    for thr in (0.3, 0.5, 0.7):
        assert np.array_equal(ro.nms(G["nms_rand_boxes"], G["nms_rand_scores"], thr), G[f"nms_rand_keep_{int(thr * 10)}"])


@pytest.mark.parametrize("p", [14, 7])
def test_roi_pool_bit_exact(p):
    out = ro.roi_pool(G["roi_fmap"], G["roi_rois"], p, 1.0 / 16.0)
    assert np.array_equal(out, G[f"roi_pool_{p}"])


def test_whole_stage_matches_reference():
    """Same seeded weights, same picture, same host preprocessing as the fixture generator."""
    from PIL import Image
    from torchvision import transforms
    tf = transforms.Compose([transforms.Resize((600, 600)), transforms.ToTensor(),
                             transforms.Normalize(mean=[0.485, 0.456, 0.406], std=[0.229, 0.224, 0.225])])
    img = tf(Image.fromarray(G["image_u8"])).unsqueeze(0)
    assert np.array_equal(ro.synthetic_image(7), G["image_u8"])
    sd = ro.seeded_backbone_state(0)
    feats, spatial, boxes = ro.extract_features(sd, img)
    assert np.array_equal(boxes, G["boxes_ms_600x600"])
    assert np.array_equal(spatial, G["spatial"])
    ref = G["features"]
    err = np.abs(feats - ref).max() / np.abs(ref).max()
    assert err <= 1e-4, err
