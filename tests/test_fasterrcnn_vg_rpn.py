"""Visual Genome Faster R-CNN extractor with RPN proposals (reference
models/feature_extractors/fasterrcnn_vg_rpn.py) against tests/golden/fasterrcnn_vg_rpn.npz, written by
oracle/make_golden_vg_rpn.py from the reference's own FasterRCNNVGRPNExtractor on a seeded Visual-Genome-spelled checkpoint.

CPU: the oracle restatement against the fixture (decode 1e-6 relative: exp differs in the last place; filter / NMS / top-k /
padding / box normalisation bit-equal given the reference's inputs), the product's host schedule over the kernel stand-ins, the
checkpoint loader.  GPU: the kernels and the product path through the C ABI."""
import os

import numpy as np
import pytest
import torch

from oracle import roi_oracle as ro

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "fasterrcnn_vg_rpn.npz"))
FH, FW = (int(v) for v in G["fmap_hw"])
W, H = (int(v) for v in G["resized_size"])


def seeded():
    return ro.seeded_backbone_state(1, (3, 4, 23, 3)), ro.seeded_vg_heads(11), ro.seeded_rpn_state(13)


def matched(ref_boxes, got_boxes, iou=0.7):
    """How many reference boxes have a counterpart (IoU >= iou) among `got_boxes`: the bf16 trunk moves a proposal by up to a few
    per cent of its anchor size, so survivors are compared by overlap, not by coordinates."""
    a, b = np.asarray(ref_boxes, np.float64), np.asarray(got_boxes, np.float64)
    x1, y1 = np.maximum(a[:, None, 0], b[None, :, 0]), np.maximum(a[:, None, 1], b[None, :, 1])
    x2, y2 = np.minimum(a[:, None, 2], b[None, :, 2]), np.minimum(a[:, None, 3], b[None, :, 3])
    inter = np.clip(x2 - x1, 0, None) * np.clip(y2 - y1, 0, None)
    area = lambda t: (t[:, 2] - t[:, 0]) * (t[:, 3] - t[:, 1])
    ious = inter / (area(a)[:, None] + area(b)[None, :] - inter)
    return int((ious.max(axis=1) >= iou).sum())


def vg_rpn_checkpoint(path):
    """The fixture's checkpoint (oracle/make_golden_vg_rpn.py::vg_rpn_checkpoint)."""
    sd, heads, rpn = seeded()
    ck = {("RCNN_top.0." + k[9:] if k.startswith("RCNN_top.") else k): v for k, v in ro.vg_backbone_state(sd).items()}
    ck.update(heads)
    ck.update(rpn)
    ck["RCNN_base.0.bias"] = torch.zeros(64)
    torch.save({"model": ck}, path)
    return path


def picture():
    from PIL import Image
    return Image.fromarray(G["image_u8"])


def preprocessed():
    from PIL import Image
    from torchvision import transforms
    pic = picture()
    nw, nh, scale = ro.rpn_resize(*pic.size)
    tf = transforms.Compose([transforms.ToTensor(), transforms.Normalize(mean=[0.485, 0.456, 0.406], std=[0.229, 0.224, 0.225])])
    return tf(pic.resize((nw, nh), Image.BILINEAR)).unsqueeze(0), scale, pic.size


# ------------------------------------------------------------------------------------------------ oracle vs the reference
def test_resize_anchors_and_decode_match_the_reference():
    from multimodal_classification_b200.fasterrcnn_vg_rpn import base_anchors
    for key in G.files:
        if key.startswith("resize_"):
            w, h = map(int, key[7:].split("x"))
            assert np.allclose(np.array(ro.rpn_resize(w, h), dtype=np.float64), G[key], rtol=0, atol=0), key
    assert np.array_equal(base_anchors(), ro.rpn_base_anchors())
    props, scores = ro.rpn_decode(G["rpn_cls"], G["rpn_box"], FH, FW, H, W)
    assert np.abs(props - G["proposals_full"]).max() <= 1e-6 * np.abs(G["proposals_full"]).max()
    assert np.abs(scores - G["scores_full"]).max() <= 1e-6


def test_filter_and_selection_bit_equal_given_the_references_inputs():
    keep = ro.rpn_filter(G["proposals_full"], G["scores_full"])
    assert np.array_equal(G["proposals_full"][keep], G["kept_boxes"]) and np.array_equal(G["scores_full"][keep], G["kept_scores"])
    idx = np.argsort(-G["region_scores"], kind="stable")[:36]
    scale = float(G["scale"])
    spatial = ro.normalize_boxes((G["kept_boxes"][idx] / np.float32(scale)).astype(np.float32), 128, 96)
    assert np.array_equal(spatial, G["spatial"])


def test_oracle_matches_reference_extractor():
    sd, heads, rpn = seeded()
    img, scale, (ow, oh) = preprocessed()
    torch.set_num_threads(os.cpu_count() or 1)
    feats, spatial, boxes, kept, region = ro.vg_rpn_extract_features(sd, heads, rpn, img, scale, ow, oh)
    assert kept.shape == G["kept_boxes"].shape and np.abs(kept - G["kept_boxes"]).max() <= 1e-3
    assert np.abs(region - G["region_scores"]).max() <= 1e-4 * np.abs(G["region_scores"]).max()
    assert np.abs(spatial - G["spatial"]).max() <= 1e-6
    assert np.abs(feats - G["features"]).max() <= 1e-4 * np.abs(G["features"]).max()
    # fewer survivors than regions: grid padding
    f2, s2, _, _, _ = ro.vg_rpn_extract_features(sd, heads, rpn, img, scale, ow, oh, num_regions=320)
    assert np.abs(s2 - G["padded_spatial"]).max() <= 1e-6
    assert np.abs(f2[::8, ::16] - G["padded_features"]).max() <= 1e-4 * np.abs(G["padded_features"]).max()


def test_checkpoint_loader_counts_like_the_reference(tmp_path):
    from multimodal_classification_b200.fasterrcnn_vg_rpn import VGFasterRCNNWithRPN, load_vg_checkpoint
    model = VGFasterRCNNWithRPN(weights=None)
    assert sorted(model.state_dict().keys()) == sorted(G["model_keys"].tolist())
    stats = load_vg_checkpoint(model, vg_rpn_checkpoint(str(tmp_path / "vg.pth")))
    assert [stats["loaded"], stats["total"], stats["skipped"]] == G["loader_stats"].tolist()
    assert torch.equal(model.RCNN_rpn.RPN_Conv.weight, seeded()[2]["RCNN_rpn.RPN_Conv.weight"])


def test_feature_map_size_formula_matches_the_trunk():
    """The plan sizes its anchor buffers from the picture size alone (FasterRCNNVGRPNExtractor._fmap_hw)."""
    from multimodal_classification_b200.fasterrcnn_vg_rpn import FasterRCNNVGRPNExtractor as E
    assert E._fmap_hw(H, W) == (FH, FW)
    sd = seeded()[0]
    for h, w in ((97, 131), (160, 224), (33, 250)):
        with torch.no_grad():
            fmap = ro.forward_base(sd, torch.zeros(1, 3, h, w))
        assert E._fmap_hw(h, w) == tuple(fmap.shape[2:]), (h, w)


# ------------------------------------------------------------------------------------------------ host schedule on the CPU
@pytest.fixture
def simulated(monkeypatch):
    if torch.cuda.is_available():
        pytest.skip("the stand-ins are for the GPU-less container")
    import ops_sim
    ops_sim.install(monkeypatch)
    ops_sim.install_device_shims(monkeypatch)
    from multimodal_classification_b200 import fasterrcnn_vg_rpn as fr

    def engine(self):                     # the product refuses CPU weights; the same cache without that check
        ver = sum(p._version for p in self.parameters()) + sum(b._version for b in self.buffers())
        if self._engine is None or self._engine.version != ver:
            self._engine = fr._Engine(self, ver)
        return self._engine
    monkeypatch.setattr(fr.VGFasterRCNNWithRPN, "engine", engine)


def test_schedule_matches_oracle(simulated, tmp_path):
    from multimodal_classification_b200.fasterrcnn_vg_rpn import FasterRCNNVGRPNExtractor
    sd, heads, rpn = seeded()
    ext = FasterRCNNVGRPNExtractor(num_regions=12, weights_path=vg_rpn_checkpoint(str(tmp_path / "vg.pth")), device="cuda",
                                   weights=None, post_nms_top_n=40)
    ext.use_graphs = False
    assert ext.has_vg_weights
    g = torch.Generator().manual_seed(3)
    img = torch.randn(1, 3, 160, 224, generator=g)
    feats, spatial = ext.extract_preprocessed(img, 1.6, 140, 100)
    assert feats.shape == (12, 2048) and spatial.shape == (12, 5)
    got = ext.selected()
    # the host schedule given the product's own intermediate values: survivors and class scores -> the oracle's choice
    rf, rs, rb, _, _ = ro.vg_rpn_extract_features(sd, heads, rpn, img, 1.6, 140, 100, num_regions=12, kept=got["kept_boxes"].numpy(),
                                                  region_scores=got["region_scores"].numpy())
    assert np.array_equal(got["boxes"].numpy(), rb) and np.abs(spatial.numpy() - rs).max() <= 1e-6
    assert np.abs(feats.numpy() - rf).max() <= 2e-2 * np.abs(rf).max()
    # ... and those intermediates against the oracle's own (bf16 stand-in arithmetic: survivors overlap, scores within the bar)
    _, _, _, okept, oregion = ro.vg_rpn_extract_features(sd, heads, rpn, img, 1.6, 140, 100, num_regions=12)
    okept = okept[:40]
    shared = matched(okept, got["kept_boxes"].numpy())
    assert shared >= 0.6 * len(okept), (shared, len(okept))
    # fewer survivors than regions: the padding branch
    ext.num_regions = 60
    f2, s2 = ext.extract_preprocessed(img, 1.6, 140, 100)
    m = ext.selected()["n_keep"]
    assert f2.shape == (60, 2048) and m <= 40
    grid = ro.rpn_pad_grid(60 - m, 224, 160)
    want = ro.normalize_boxes((grid / np.float32(1.6)).astype(np.float32), 140, 100)
    assert np.array_equal(s2.numpy()[m:], want[: 60 - m])


# ------------------------------------------------------------------------------------------------ kernels and the product path on the B200
@pytest.mark.gpu
def test_rpn_kernels_against_the_reference_values():
    from multimodal_classification_b200 import ops
    from multimodal_classification_b200.fasterrcnn_vg_rpn import base_anchors
    dev = "cuda"
    a = FH * FW * 12
    heads = torch.cat([torch.from_numpy(G["rpn_cls"]).view(FH * FW, 24), torch.from_numpy(G["rpn_box"]).view(FH * FW, 48)], dim=1).to(dev)
    props, scores = torch.empty(a, 4, device=dev), torch.empty(a, device=dev)
    nv = torch.zeros(1, dtype=torch.int32, device=dev)
    ops.rpn_decode(heads, FH, FW, base_anchors(), 16, H, W, 16, props, scores, nv)
    ref_p, ref_s = G["proposals_full"], G["scores_full"]
    assert np.abs(props.cpu().numpy() - ref_p).max() <= 1e-6 * np.abs(ref_p).max()
    valid = (ref_p[:, 2] - ref_p[:, 0] >= 16) & (ref_p[:, 3] - ref_p[:, 1] >= 16)
    got_s = scores.cpu().numpy()
    assert int(nv.item()) == int(valid.sum()) and np.all(np.isneginf(got_s[~valid]))
    assert np.abs(got_s[valid] - ref_s[valid]).max() <= 1e-6
    # sort / gather / NMS on the REFERENCE's proposals and scores: bit-equal survivors
    p_ref = torch.from_numpy(ref_p).to(dev)
    s_ref = torch.from_numpy(np.where(valid, ref_s, -np.inf).astype(np.float32)).to(dev)
    order = torch.zeros(a, dtype=torch.int32, device=dev)
    ops.rank_sort_desc(s_ref, order)
    assert torch.equal(order.long().cpu(), torch.sort(s_ref.cpu(), descending=True, stable=True)[1])
    top_b, top_s = torch.empty(6000, 4, device=dev), torch.empty(6000, device=dev)
    cnt, nk = torch.zeros(1, dtype=torch.int32, device=dev), torch.zeros(1, dtype=torch.int32, device=dev)
    ops.gather_sorted(p_ref, s_ref, order, nv, top_b, top_s, cnt)
    keep = torch.zeros(300, dtype=torch.int32, device=dev)
    ops.nms_sorted(top_b, cnt, 0.7, keep, nk)
    assert int(cnt.item()) == 6000 and int(nk.item()) == len(G["kept_boxes"])
    kept = top_b[keep[: int(nk.item())].long()].cpu().numpy()
    assert np.array_equal(kept, G["kept_boxes"]) and np.array_equal(top_s[keep.long()].cpu().numpy(), G["kept_scores"])
    # limit: elements beyond it sort last
    lim = torch.tensor([100], dtype=torch.int32, device=dev)
    rs = torch.from_numpy(G["region_scores"]).to(dev)
    ro_ = torch.zeros(rs.numel(), dtype=torch.int32, device=dev)
    ops.rank_sort_desc(rs, ro_, limit=lim)
    assert torch.equal(ro_[:100].long().cpu(), torch.sort(rs[:100].cpu(), descending=True, stable=True)[1])
    # box_div of the final selection: the reference's boxes / scale, then normalisation
    ops.rank_sort_desc(rs, ro_)
    kb = torch.from_numpy(G["kept_boxes"]).to(dev)
    n_all = torch.tensor([kb.shape[0]], dtype=torch.int32, device=dev)
    spatial = torch.zeros(36, 5, device=dev)
    ops.select_regions(kb, ro_, n_all, 36, 128, 96, spatial=spatial, box_div=float(G["scale"]))
    assert np.array_equal(spatial.cpu().numpy(), G["spatial"])


@pytest.fixture(scope="module")
def extractor(tmp_path_factory):
    from multimodal_classification_b200.fasterrcnn_vg_rpn import FasterRCNNVGRPNExtractor
    path = vg_rpn_checkpoint(str(tmp_path_factory.mktemp("vgrpn") / "vg.pth"))
    return FasterRCNNVGRPNExtractor(weights_path=path, device="cuda", weights=None)


@pytest.mark.gpu
def test_vg_rpn_extractor_vs_reference(extractor):
    pic = picture()
    feats, spatial = extractor.extract_features(pic)
    assert feats.shape == (36, 2048) and spatial.shape == (36, 5) and feats.dtype == torch.float32
    got = extractor.selected()
    # objectness of the proposals the reference also kept as valid: within the bf16 bar
    ref_s, ref_p = G["scores_full"], G["proposals_full"]
    valid = (ref_p[:, 2] - ref_p[:, 0] >= 16) & (ref_p[:, 3] - ref_p[:, 1] >= 16)
    gs = got["scores"].cpu().numpy()
    both = valid & np.isfinite(gs)
    err = np.abs(gs[both] - ref_s[both]).max()
    print(f"objectness: max |d| {err:.4f} over {both.sum()} proposals; survivors {got['n_keep']}")
    assert both.sum() >= 0.98 * valid.sum() and err <= 2e-2
    # the whole chain given the product's own survivors and class scores = the oracle's selection, boxes and features
    sd, heads, rpn = seeded()
    img, scale, (ow, oh) = preprocessed()
    torch.set_num_threads(os.cpu_count() or 1)
    rf, rs, rb, _, _ = ro.vg_rpn_extract_features(sd, heads, rpn, img, scale, ow, oh, kept=got["kept_boxes"].cpu().numpy(),
                                                  region_scores=got["region_scores"].cpu().numpy())
    assert np.array_equal(got["boxes"].cpu().numpy(), rb) and np.array_equal(spatial.cpu().numpy(), rs)
    rel = np.abs(feats.cpu().numpy() - rf).max() / np.abs(rf).max()
    print(f"features of the chosen regions vs oracle: max-rel {rel:.4f}")
    assert rel <= 2e-2
    # survivors shared with the reference's run (bf16 trunk: near-ties in objectness reorder the NMS input)
    shared = matched(G["kept_boxes"], got["kept_boxes"].cpu().numpy())
    print(f"reference survivors with a counterpart (IoU >= 0.7) among ours: {shared} of {len(G['kept_boxes'])}")
    assert shared >= 0.6 * len(G["kept_boxes"])
    again, _ = extractor.extract_features(pic)                    # graph replay
    assert torch.equal(feats, again)


@pytest.mark.gpu
def test_vg_rpn_padding_branch_and_batch(extractor):
    pic = picture()
    extractor.num_regions = 320
    try:
        feats, spatial = extractor.extract_features(pic)
        m = extractor.selected()["n_keep"]
        assert feats.shape == (320, 2048) and m <= 300
        grid = ro.rpn_pad_grid(320 - m, W, H)
        want = ro.normalize_boxes((grid / np.float32(float(G["scale"]))).astype(np.float32), 128, 96)
        assert np.array_equal(spatial.cpu().numpy()[m:], want[: 320 - m])
        ref = G["padded_features"]
        rows = [r for r in range(0, 320, 8) if r >= max(m, 300)]                                        # grid cells: same boxes in both runs
        got = feats.cpu().numpy()[rows][:, ::16]
        assert rows and np.abs(got - ref[[r // 8 for r in rows]]).max() <= 2e-2 * np.abs(ref).max()
    finally:
        extractor.num_regions = 36
    x = torch.from_numpy(G["image_u8"]).permute(2, 0, 1).float().div(255)
    f, s = extractor.forward(torch.stack([x, x.flip(2)]))
    assert f.shape == (2, 36, 2048) and s.shape == (2, 36, 5) and not torch.equal(f[0], f[1])


@pytest.mark.gpu
def test_vg_rpn_cpu_device_is_refused():
    from multimodal_classification_b200._lib import VbError
    from multimodal_classification_b200.fasterrcnn_vg_rpn import FasterRCNNVGRPNExtractor
    with pytest.raises(VbError):
        FasterRCNNVGRPNExtractor(device="cpu", weights=None)
