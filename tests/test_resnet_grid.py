"""Grid feature extractor (reference models/feature_extractors/resnet.py) against tests/golden/
resnet_grid.npz, written by oracle/make_golden_grid.py from the reference's own ResNetFeatureExtractor on seeded weights."""
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import roi_oracle as ro

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "resnet_grid.npz"))


def preprocessed():
    from PIL import Image
    from torchvision import transforms
    tf = transforms.Compose([transforms.Resize((224, 224)), transforms.ToTensor(),
                             transforms.Normalize(mean=[0.485, 0.456, 0.406], std=[0.229, 0.224, 0.225])])
    return tf(Image.fromarray(G["image_u8"])).unsqueeze(0)


def test_oracle_matches_reference_grid_features():
    sd, img = ro.seeded_backbone_state(0), preprocessed()
    for n in (36, 49, 9):
        got = ro.grid_features(sd, img, n)
        ref = G[f"features_{n}"]
        got = got if n == 36 else got[:, ::16]
        assert got.shape == ref.shape
        assert np.abs(got - ref).max() <= 1e-4 * np.abs(ref).max(), n
        assert np.array_equal(ro.grid_spatial(n)[: int(n ** 0.5) ** 2], G[f"spatial_{n}"])
    assert ro.grid_features(sd, img, 36, 2304)[:, 2048:].max() == 0 and ro.grid_features(sd, img, 36, 100).shape == (36, 100)


def test_backbone_keys_are_the_reference_extractors():
    import torchvision
    ref_keys = list(torch.nn.Sequential(*list(torchvision.models.resnet152(weights=None).children())[:-2]).state_dict().keys())
    assert sorted(ro.grid_backbone_state(ro.seeded_backbone_state(0)).keys()) == sorted(ref_keys)


def test_adaptive_windows_agree_with_torch():
    from multimodal_classification_b200._lib import VbError
    from multimodal_classification_b200.resnet_grid import adaptive_windows
    x = torch.randn(1, 3, 7, 7)
    for g in (1, 2, 3, 6, 7):
        k, s = adaptive_windows(7, g)
        assert torch.allclose(F.avg_pool2d(x, k, s), F.adaptive_avg_pool2d(x, (g, g)), atol=1e-6), g
    for g in (4, 5, 10):
        with pytest.raises(VbError):
            adaptive_windows(7, g)


@pytest.fixture(scope="module")
def extractor():
    from multimodal_classification_b200.resnet_grid import ResNetFeatureExtractor
    ext = ResNetFeatureExtractor(device="cuda", weights=None)
    ext.backbone.load_state_dict(ro.grid_backbone_state(ro.seeded_backbone_state(0)), strict=True)
    return ext


@pytest.mark.gpu
def test_grid_features_vs_reference(extractor):
    from PIL import Image
    pic = Image.fromarray(G["image_u8"])
    feats, spatial = extractor.extract_features(pic)
    assert feats.shape == (36, 2048) and feats.dtype == torch.float32 and spatial.shape == (36, 5)
    assert np.array_equal(spatial.cpu().numpy(), G["spatial_36"])
    ref, got = G["features_36"], feats.cpu().numpy()
    rel, mx = np.linalg.norm(got - ref) / np.linalg.norm(ref), np.abs(got - ref).max() / np.abs(ref).max()
    print(f"grid features: rel-L2 {rel:.4f}  max-rel {mx:.4f}")
    assert rel <= 2e-2 and mx <= 2e-2, (rel, mx)
    again, _ = extractor.extract_features(pic)                    # graph replay
    assert torch.equal(feats, again)
    for n in (49, 9):
        extractor.num_regions = n
        extractor._plans.clear()
        f, s = extractor.extract_features(pic)
        r = G[f"features_{n}"]
        assert np.array_equal(s.cpu().numpy(), G[f"spatial_{n}"])
        assert np.abs(f.cpu().numpy()[:, ::16] - r).max() <= 2e-2 * np.abs(r).max(), n
    extractor.num_regions = 36
    extractor._plans.clear()


@pytest.mark.gpu
def test_grid_batch_padding_and_refusals(extractor):
    from multimodal_classification_b200._lib import VbError
    from multimodal_classification_b200.resnet_grid import ResNetFeatureExtractor
    imgs = torch.stack([torch.from_numpy(G["image_u8"]).permute(2, 0, 1), torch.from_numpy(G["image_u8"][::-1].copy()).permute(2, 0, 1)])
    f, s = extractor.forward(imgs)
    assert f.shape == (2, 36, 2048) and s.shape == (2, 36, 5)
    ref = G["features_36"]
    assert np.abs(f[0].cpu().numpy() - ref).max() <= 2e-2 * np.abs(ref).max()
    assert (f[0] - f[1]).abs().max().item() > 0.1
    extractor.output_dim = 2304
    wide, _ = extractor.forward(imgs)
    extractor.output_dim = 100
    narrow, _ = extractor.forward(imgs)
    extractor.output_dim = 2048
    assert wide.shape == (2, 36, 2304) and torch.equal(wide[..., :2048], f) and wide[..., 2048:].abs().max().item() == 0
    assert torch.equal(narrow, f[..., :100])
    with pytest.raises(VbError):
        ResNetFeatureExtractor(device="cpu", weights=None)
    extractor.num_regions = 25
    extractor._plans.clear()
    with pytest.raises(VbError, match="non-uniform"):
        extractor.forward(imgs)
    extractor.num_regions = 36
    extractor._plans.clear()


# ------------------------------------------------------------------------------------------------ VG ResNet-101 variant
VG_BLOCKS = (3, 4, 23, 3)


def test_oracle_matches_reference_vg_grid_features():
    got = ro.grid_features(ro.seeded_backbone_state(1, VG_BLOCKS), preprocessed(), 36)[:, ::4]
    ref = G["vg_features_36"]
    assert np.abs(got - ref).max() <= 1e-4 * np.abs(ref).max()
    assert np.array_equal(ro.grid_spatial(36), G["vg_spatial_36"])


def test_vg_checkpoint_loader_matches_reference_loader(tmp_path):
    """Same statistics and same resulting weights as the reference's ``load_vg_backbone_weights`` on a checkpoint with the VG
    file's key spelling (RCNN_top.0.*), foreign keys and one wrong shape (fixture: the reference loader's own report)."""
    from multimodal_classification_b200.resnet_grid import VGResNet101Backbone, load_vg_backbone_weights
    seeded = ro.vg_backbone_state(ro.seeded_backbone_state(1, VG_BLOCKS))
    src = VGResNet101Backbone(weights=None)
    src.load_state_dict(seeded, strict=True)
    ck = {("RCNN_top.0." + k[9:] if k.startswith("RCNN_top.") else k): v for k, v in src.state_dict().items()}
    ck["RCNN_rpn.RPN_Conv.weight"] = torch.zeros(4)
    ck["RCNN_cls_score.weight"] = torch.zeros(4)
    ck["RCNN_base.0.weight"] = torch.zeros(64, 3, 3, 3)
    ck["RCNN_base.9.weight"] = torch.zeros(1)
    path = str(tmp_path / "vg.pth")
    torch.save({"model": ck}, path)
    torch.manual_seed(3)
    fresh = VGResNet101Backbone(weights=None)
    stem = fresh.state_dict()["RCNN_base.0.weight"].clone()
    stats = load_vg_backbone_weights(fresh, path)
    assert [stats["loaded"], stats["total_model"], stats["skipped"]] == G["vg_loader_stats"].tolist()
    assert stats["skipped_keys"] == [str(k) for k in G["vg_loader_skipped_keys"]]
    got = fresh.state_dict()
    assert torch.equal(got["RCNN_base.0.weight"], stem)                         # wrong shape in the checkpoint: kept
    assert all(torch.equal(got[k], seeded[k]) for k in seeded if k != "RCNN_base.0.weight")


@pytest.mark.gpu
def test_vg_grid_features_vs_reference():
    from PIL import Image
    from multimodal_classification_b200.resnet_grid import ResNetVGExtractor
    ext = ResNetVGExtractor(device="cuda", weights=None, weights_path="/nonexistent.pth")
    assert not ext.has_vg_weights and ext.grid_size == 6
    ext.backbone.load_state_dict(ro.vg_backbone_state(ro.seeded_backbone_state(1, VG_BLOCKS)), strict=True)
    feats, spatial = ext.extract_features(Image.fromarray(G["image_u8"]))
    assert feats.shape == (36, 2048) and np.array_equal(spatial.cpu().numpy(), G["vg_spatial_36"])
    ref, got = G["vg_features_36"], feats.cpu().numpy()[:, ::4]
    rel, mx = np.linalg.norm(got - ref) / np.linalg.norm(ref), np.abs(got - ref).max() / np.abs(ref).max()
    print(f"vg grid features: rel-L2 {rel:.4f}  max-rel {mx:.4f}")
    assert rel <= 2e-2 and mx <= 2e-2, (rel, mx)
