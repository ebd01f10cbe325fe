"""Visual Genome Faster R-CNN extractor (reference models/feature_extractors/fasterrcnn_vg.py) against
tests/golden/fasterrcnn_vg.npz, written by oracle/make_golden_vg.py from the reference's own FasterRCNNVGExtractor on a seeded
checkpoint with the Visual Genome file's key spelling.

CPU: the oracle restatement against the fixture (candidates, selection: bit-equal; scores / features 1e-4), the host schedule
of the product module over the kernel stand-ins, the checkpoint loader.  GPU: the product path through the C ABI."""
import os

import numpy as np
import pytest
import torch

from oracle import roi_oracle as ro

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "fasterrcnn_vg.npz"))
SIZES = [(600, 1000), (224, 224), (480, 640), (97, 1000)]


def preprocessed(size=(600, 1000)):
    from PIL import Image
    from torchvision import transforms
    tf = transforms.Compose([transforms.Resize(size), transforms.ToTensor(),
                             transforms.Normalize(mean=[0.485, 0.456, 0.406], std=[0.229, 0.224, 0.225])])
    return tf(Image.fromarray(G["image_u8"])).unsqueeze(0)


def seeded():
    return ro.seeded_backbone_state(1, (3, 4, 23, 3)), ro.seeded_vg_heads(11)


def vg_checkpoint(path):
    """The fixture's checkpoint (oracle/make_golden_vg.py::vg_checkpoint), Visual Genome key spelling."""
    sd, heads = seeded()
    ck = {("RCNN_top.0." + k[9:] if k.startswith("RCNN_top.") else k): v for k, v in ro.vg_backbone_state(sd).items()}
    ck.update(heads)
    ck["RCNN_rpn.RPN_Conv.weight"] = torch.zeros(512, 1024, 3, 3)
    ck["RCNN_base.0.bias"] = torch.zeros(64)
    torch.save({"model": ck}, path)
    return path


# ------------------------------------------------------------------------------------------------ oracle vs the reference
def test_candidates_bit_equal_to_reference():
    from multimodal_classification_b200.fasterrcnn_vg import vg_grid_candidates
    for h, w in SIZES:
        ref = G[f"cands_{h}x{w}"]
        assert np.array_equal(ro.vg_grid_candidates(h, w), ref), (h, w)
        assert np.array_equal(vg_grid_candidates(h, w), ref), (h, w)          # the product's host arithmetic
    assert np.array_equal(ro.vg_grid_candidates(600, 1000), G["candidates"])


def test_selection_bit_equal_to_reference_given_its_scores():
    c, s = G["candidates"], G["scores"]
    assert np.array_equal(c[ro.vg_select(c, s, 36, 0.3)], G["boxes"])
    assert np.array_equal(s[ro.vg_select(c, s, 36, 0.3)], G["sel_scores"])
    for n in (10, 100):                                                       # 100 > NMS survivors: padding with the last one
        assert np.array_equal(c[ro.vg_select(c, s, n, 0.3)], G[f"boxes_{n}"]), n
    # without the checkpoint every score is 1.0 and the reference's torch.topk picks among exact ties in an unspecified
    # (implementation-dependent) order: both choices are NMS survivors, the stable one is the first 36 of them
    keep = ro.nms(c, np.ones(len(c), np.float32), 0.3)
    ours = ro.vg_select(c, np.ones(len(c), np.float32), 36, 0.3)
    assert np.array_equal(ours, keep[:36])
    surv = {tuple(b) for b in c[keep]}
    assert all(tuple(b) in surv for b in G["plain_boxes"]) and len({tuple(b) for b in G["plain_boxes"]}) == 36


def test_oracle_matches_reference_extractor():
    sd, heads = seeded()
    torch.set_num_threads(os.cpu_count() or 1)
    feats, spatial, boxes, scores = ro.vg_extract_features(sd, heads, preprocessed())
    assert np.abs(scores - G["scores"]).max() <= 1e-4 * np.abs(G["scores"]).max()
    assert np.array_equal(boxes, G["boxes"]) and np.array_equal(spatial, G["spatial"])
    assert np.abs(feats - G["features"]).max() <= 1e-4 * np.abs(G["features"]).max()


def test_checkpoint_loader_counts_like_the_reference(tmp_path):
    from multimodal_classification_b200.fasterrcnn_vg import VGFasterRCNN, load_vg_weights
    model = VGFasterRCNN(weights=None)
    assert sorted(model.state_dict().keys()) == sorted(G["model_keys"].tolist())
    assert load_vg_weights(model, vg_checkpoint(str(tmp_path / "vg.pth"))) == int(G["loaded_count"])
    sd, heads = seeded()
    assert torch.equal(model.RCNN_top[2].conv3.weight, sd["top.2.conv3.weight"])
    assert torch.equal(model.RCNN_cls_score.weight, heads["RCNN_cls_score.weight"])


# ------------------------------------------------------------------------------------------------ host schedule on the CPU
@pytest.fixture
def simulated(monkeypatch):
    if torch.cuda.is_available():
        pytest.skip("the stand-ins are for the GPU-less container")
    import ops_sim
    ops_sim.install(monkeypatch)
    ops_sim.install_device_shims(monkeypatch)
    from multimodal_classification_b200 import fasterrcnn_vg as fv

    def engine(self):                     # the product refuses CPU weights; the same cache without that check
        ver = sum(p._version for p in self.parameters()) + sum(b._version for b in self.buffers())
        if self._engine is None or self._engine.version != ver:
            self._engine = fv._Engine(self, ver)
        return self._engine
    monkeypatch.setattr(fv.VGFasterRCNN, "engine", engine)


@pytest.mark.parametrize("with_checkpoint", [True, False])
def test_schedule_matches_oracle(simulated, tmp_path, with_checkpoint):
    from multimodal_classification_b200.fasterrcnn_vg import FasterRCNNVGExtractor
    sd, heads = seeded()
    path = vg_checkpoint(str(tmp_path / "vg.pth")) if with_checkpoint else str(tmp_path / "absent.pth")
    ext = FasterRCNNVGExtractor(num_regions=12, weights_path=path, device="cuda", weights=None, image_size=(96, 160))
    ext.use_graphs = False
    assert ext.has_vg_weights == with_checkpoint
    if not with_checkpoint:
        ext.model.load_state_dict({**ro.vg_backbone_state(sd), **heads}, strict=True)
    g = torch.Generator().manual_seed(3)
    img = torch.randn(2, 3, 96, 160, generator=g)
    feats, spatial = ext.extract_batch(img)
    assert feats.shape == (2, 12, 2048) and spatial.shape == (2, 12, 5)
    boxes, index, scores = ext.selected(2)
    for b in range(2):
        # selection parity under the product's own (bf16-rounded) scores, then features of those boxes against the oracle
        rf, rs, rb, rscores = ro.vg_extract_features(sd, heads, img[b:b + 1], num_regions=12, has_vg_weights=with_checkpoint,
                                                     scores=scores[b].numpy())
        assert np.array_equal(boxes[b].numpy(), rb) and np.array_equal(spatial[b].numpy(), rs)
        assert np.abs(feats[b].numpy() - rf).max() <= 2e-2 * np.abs(rf).max()
        if with_checkpoint:
            true_scores = ro.vg_extract_features(sd, heads, img[b:b + 1], num_regions=12)[3]
            assert np.abs(scores[b].numpy() - true_scores).max() <= 2e-2 * np.abs(true_scores).max()
    again, _ = ext.extract_batch(img)
    assert torch.equal(again, feats)
    # more regions than candidates survive: padded with the last survivor, as _pad_regions does
    ext.num_regions = 250
    f2, s2 = ext.extract_batch(img[:1])
    assert f2.shape == (1, 250, 2048) and torch.equal(s2[0, -1], s2[0, 199]) and torch.equal(f2[0, -1], f2[0, 199])


# ------------------------------------------------------------------------------------------------ the product path on the B200
@pytest.fixture(scope="module")
def extractor(tmp_path_factory):
    from multimodal_classification_b200.fasterrcnn_vg import FasterRCNNVGExtractor
    path = vg_checkpoint(str(tmp_path_factory.mktemp("vg") / "vg.pth"))
    return FasterRCNNVGExtractor(weights_path=path, device="cuda", weights=None)


@pytest.mark.gpu
def test_select_kernels_bit_exact():
    """vb_rowmax_f32 / vb_nms / vb_select_regions against the reference's own selection on the reference's scores."""
    from multimodal_classification_b200 import ops
    dev = "cuda"
    c, s = torch.from_numpy(G["candidates"]).to(dev), torch.from_numpy(G["scores"]).to(dev)
    n = c.shape[0]
    g = torch.Generator().manual_seed(0)
    x = torch.randn(n, 1608, generator=g).to(dev)
    out = torch.empty(n, device=dev)
    ops.rowmax(x, out, 1, 1601)
    assert torch.equal(out, x[:, 1:1601].max(dim=1)[0])
    feat_src = torch.randn(n, 2048, generator=g).to(dev)
    for regions, key in ((36, "boxes"), (10, "boxes_10"), (100, "boxes_100")):
        ws, keep, nk = torch.zeros(2 * n, dtype=torch.int32, device=dev), torch.zeros(n, dtype=torch.int32, device=dev), torch.zeros(1, dtype=torch.int32, device=dev)
        ops.nms_device(c, s, 0.3, ws, keep, nk)
        boxes, spatial = torch.zeros(regions, 4, device=dev), torch.zeros(regions, 5, device=dev)
        index, rois = torch.zeros(regions, dtype=torch.int32, device=dev), torch.zeros(regions, 5, device=dev)
        feat = torch.zeros(regions, 2048, device=dev)
        ops.select_regions(c, keep, nk, regions, 1000, 600, boxes=boxes, spatial=spatial, index=index, rois=rois, batch_index=3,
                           feat_src=feat_src, feat_dst=feat)
        assert np.array_equal(boxes.cpu().numpy(), G[key]), key
        assert np.array_equal(spatial.cpu().numpy(), ro.normalize_boxes(G[key], 1000, 600))
        assert torch.equal(feat, feat_src[index.long()]) and torch.equal(rois[:, 1:], boxes) and bool((rois[:, 0] == 3).all())
        assert np.array_equal(c[index.long()].cpu().numpy(), G[key])
    assert np.array_equal(spatial.cpu().numpy()[:36], G["spatial"][:36]) or regions != 36


@pytest.mark.gpu
def test_vg_extractor_vs_reference(extractor):
    from PIL import Image
    pic = Image.fromarray(G["image_u8"])
    feats, spatial = extractor.extract_features(pic)
    assert feats.shape == (36, 2048) and feats.dtype == torch.float32 and spatial.shape == (36, 5)
    boxes, index, scores = extractor.selected(1)
    got_scores, ref_scores = scores[0].cpu().numpy(), G["scores"]
    err = np.abs(got_scores - ref_scores).max() / np.abs(ref_scores).max()
    print(f"candidate scores: max-rel {err:.4f}")
    assert err <= 2e-2, err
    # the selection is bit-exact with the reference's algorithm on the scores the bf16 trunk produced ...
    idx = ro.vg_select(G["candidates"], got_scores, 36, 0.3)
    assert np.array_equal(index[0].cpu().numpy(), idx)
    assert np.array_equal(spatial.cpu().numpy(), ro.normalize_boxes(G["candidates"][idx], 1000, 600))
    # ... and every region the reference also chose carries the reference's feature vector within the bf16 bar
    ref_rows = {tuple(b): i for i, b in enumerate(G["boxes"])}
    common = [(r, ref_rows[tuple(b)]) for r, b in enumerate(boxes[0].cpu().numpy()) if tuple(b) in ref_rows]
    print(f"regions shared with the reference's selection: {len(common)} of 36")
    assert len(common) >= 24
    got, ref = feats.cpu().numpy()[[r for r, _ in common]], G["features"][[i for _, i in common]]
    rel, mx = np.linalg.norm(got - ref) / np.linalg.norm(ref), np.abs(got - ref).max() / np.abs(ref).max()
    print(f"VG RoI features: rel-L2 {rel:.4f}  max-rel {mx:.4f}")
    assert rel <= 2e-2 and mx <= 2e-2, (rel, mx)
    again, _ = extractor.extract_features(pic)                    # graph replay
    assert torch.equal(feats, again)


@pytest.mark.gpu
def test_vg_extractor_features_of_all_candidates_vs_oracle(extractor):
    """Size-independent check of the scored branch: the feature row of EVERY chosen region equals the oracle's RoIPool-14 ->
    layer4 -> mean of that box on the same picture (whatever the selection), within the bf16 bar."""
    sd, heads = seeded()
    img = preprocessed()
    feats, spatial = extractor.extract_batch(img.cuda())
    boxes, index, scores = extractor.selected(1)
    torch.set_num_threads(os.cpu_count() or 1)
    rf, rs, rb, _ = ro.vg_extract_features(sd, heads, img, scores=scores[0].cpu().numpy())
    assert np.array_equal(boxes[0].cpu().numpy(), rb) and np.array_equal(spatial[0].cpu().numpy(), rs)
    err = np.abs(feats[0].cpu().numpy() - rf).max() / np.abs(rf).max()
    print(f"chosen-region features vs oracle: max-rel {err:.4f}")
    assert err <= 2e-2


@pytest.mark.gpu
def test_vg_extractor_without_checkpoint_and_batched(extractor, tmp_path):
    from multimodal_classification_b200.fasterrcnn_vg import FasterRCNNVGExtractor
    plain = FasterRCNNVGExtractor(weights_path=str(tmp_path / "absent.pth"), device="cuda", weights=None, num_regions=36)
    assert not plain.has_vg_weights
    plain.model.load_state_dict(extractor.model.state_dict(), strict=True)
    img = preprocessed().cuda()
    feats, spatial = plain.extract_batch(torch.cat([img, img.flip(3)]))
    boxes, index, _ = plain.selected(2)
    keep = ro.nms(G["candidates"], np.ones(200, np.float32), 0.3)[:36]
    assert np.array_equal(index[0].cpu().numpy(), keep) and np.array_equal(index[1].cpu().numpy(), keep)
    # rows the reference also chose (its top-k order among exact ties differs): same features
    ref_rows = {tuple(b): i for i, b in enumerate(G["plain_boxes"])}
    common = [(r, ref_rows[tuple(b)]) for r, b in enumerate(boxes[0].cpu().numpy()) if tuple(b) in ref_rows]
    assert len(common) >= 18
    got, ref = feats[0].cpu().numpy()[[r for r, _ in common]][:, ::8], G["plain_features"][[i for _, i in common]]
    assert np.abs(got - ref).max() <= 2e-2 * np.abs(ref).max()
    assert not torch.equal(feats[0], feats[1])
    # the scored extractor on a batch of two equals two single-picture runs
    f2, s2 = extractor.extract_batch(torch.cat([img, img.flip(3)]))
    f1, s1 = extractor.extract_batch(img)
    assert torch.equal(s2[0], s1[0]) and np.abs((f2[0] - f1[0]).cpu().numpy()).max() <= 1e-3 * float(f1.abs().max())


@pytest.mark.gpu
def test_vg_cpu_device_is_refused():
    from multimodal_classification_b200._lib import VbError
    from multimodal_classification_b200.fasterrcnn_vg import FasterRCNNVGExtractor
    with pytest.raises(VbError):
        FasterRCNNVGExtractor(device="cpu", weights=None)
