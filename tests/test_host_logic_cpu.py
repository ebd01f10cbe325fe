"""Host-side logic of the product (no kernel launches) against the reference-produced fixtures, on the CPU: candidate box
generation, grid boxes, box normalisation (resnet152_roi.py), the DINOv2 grid boxes (dinov2_fusion.py), the gradient-range
merging of ddp.py and the flat parameter layout of vilbert.py."""
import os

import numpy as np
import torch

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "roi_stage.npz"))
D = np.load(os.path.join(os.path.dirname(__file__), "golden", "dinov2_tail.npz"))


def test_candidate_and_grid_boxes_bit_exact():
    from multimodal_classification_b200 import resnet152_roi as rr
    assert np.array_equal(rr.sliding_window_boxes(600, 600), G["nms_cands"])
    for h, w in [tuple(int(v) for v in hw) for hw in G["proposal_sizes"]]:
        assert np.array_equal(rr.grid_boxes(36, h, w), G[f"boxes_grid_{h}x{w}"])
        assert np.array_equal(rr.normalize_boxes(G[f"boxes_ms_{h}x{w}"], w, h), G[f"spatial_ms_{h}x{w}"])


def test_dinov2_grid_spatial_bit_exact():
    from multimodal_classification_b200.dinov2_fusion import grid_spatial
    assert np.array_equal(grid_spatial(36).numpy(), D["spatial"])


def test_merge_ranges():
    from multimodal_classification_b200 import ddp
    assert ddp.merge_ranges([(10, 20), (0, 10), (30, 40), (40, 50), (35, 38)]) == [(0, 20), (30, 50)]
    assert ddp.merge_ranges([]) == []


def test_flat_layout_keeps_fused_operands_adjacent_and_parameters_as_views():
    from multimodal_classification_b200.vilbert import ViLBERTForClassification, _FlatParams
    from oracle import vilbert_oracle as vo
    torch.manual_seed(0)
    m = ViLBERTForClassification(vo.tiny_config(), num_labels=2)
    before = {k: v.clone() for k, v in m.state_dict().items()}
    flat = _FlatParams(m, torch.device("cpu"))
    p = "bert.encoder.layer.0.attention.self"
    flat.check_contiguous([p + ".query.weight", p + ".key.weight", p + ".value.weight"])
    flat.check_contiguous([p + ".query.bias", p + ".key.bias", p + ".value.bias"])
    # flattening must not change a single value or key of the state_dict
    after = m.state_dict()
    assert list(after.keys()) == list(before.keys())
    assert all(torch.equal(after[k], before[k]) for k in before)
    assert flat.intact() and flat.w_end % 64 == 0 and flat.s_end <= flat.total
