"""Host schedule of the RoI / grid feature stage (multimodal_classification_b200/resnet152_roi.py, resnet_grid.py: BatchNorm
folding, NHWC weight layout, the 155-convolution trunk over a scratch arena, proposals, plan caching, padding / truncation)
run in the GPU-less container over the functional stand-ins of tests/ops_sim.py, against the fp32 oracle."""
import numpy as np
import pytest
import torch

from oracle import roi_oracle as ro


@pytest.fixture
def simulated(monkeypatch):
    if torch.cuda.is_available():
        pytest.skip("the stand-ins are for the GPU-less container")
    import ops_sim
    ops_sim.install(monkeypatch)
    ops_sim.install_device_shims(monkeypatch)
    # the backbone refuses CPU weights (product behaviour, tested in test_roi_gpu.py::test_cpu_device_is_refused); here the
    # same trunk cache is used without that check
    from multimodal_classification_b200 import resnet152_roi as rr

    def trunk(self):
        ver = sum(p._version for p in self.parameters()) + sum(b._version for b in self.buffers())
        if self._trunk is None or self._trunk.version != ver:
            self._trunk = rr._Trunk(self, ver)
        return self._trunk
    monkeypatch.setattr(rr.ResNet152Backbone, "trunk", trunk)


def test_roi_stage_schedule_matches_oracle(simulated):
    from multimodal_classification_b200.resnet152_roi import ResNet152ROIExtractor
    sd = ro.seeded_backbone_state(0)
    ext = ResNet152ROIExtractor(device="cuda", weights=None, image_size=96, roi_size=7)
    ext.use_graphs = False
    ext.backbone.load_state_dict(sd, strict=True)
    g = torch.Generator().manual_seed(5)
    img = torch.randn(2, 3, 96, 96, generator=g)
    feats, spatial = ext.extract_batch(img)
    assert feats.shape == (2, 36, 2048) and spatial.shape == (2, 36, 5)
    for b in range(2):
        ref_f, ref_s, ref_boxes = ro.extract_features(sd, img[b:b + 1], roi_size=7)
        assert np.array_equal(spatial[b].numpy(), ref_s)
        err = np.abs(feats[b].numpy() - ref_f).max() / np.abs(ref_f).max()
        assert err <= 2e-2, (b, err)
    assert np.array_equal(ext._generate_proposals(96, 96).numpy(), ro.proposals(36, 96, 96, True))
    again, _ = ext.extract_batch(img)                       # second call: same plan, same scratch arena
    assert torch.equal(again, feats)


def test_grid_extractor_schedule_matches_oracle(simulated):
    from multimodal_classification_b200.resnet_grid import ResNetFeatureExtractor
    sd = ro.seeded_backbone_state(0)
    ext = ResNetFeatureExtractor(device="cuda", weights=None)
    ext.use_graphs = False
    ext.backbone.load_state_dict(ro.grid_backbone_state(sd), strict=True)
    g = torch.Generator().manual_seed(6)
    img = torch.randn(1, 3, 224, 224, generator=g)
    feats, spatial = ext.extract_batch(img)
    ref = ro.grid_features(sd, img, 36)
    assert feats.shape == (1, 36, 2048) and np.array_equal(spatial[0].numpy(), ro.grid_spatial(36))
    assert np.abs(feats[0].numpy() - ref).max() <= 2e-2 * np.abs(ref).max()
    ext.output_dim = 100
    narrow, _ = ext.extract_batch(img)
    assert narrow.shape == (1, 36, 100) and torch.equal(narrow, feats[..., :100])
