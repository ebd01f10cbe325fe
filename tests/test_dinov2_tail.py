"""DINOv2 multi-layer fusion tail: oracle vs the fixture produced by the reference's own code
(oracle/make_golden_dinov2.py), and the CUDA path vs the same fixture."""
import os

import numpy as np
import pytest
import torch

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "dinov2_tail.npz"))


def test_oracle_matches_reference_fixture():
    from oracle import roi_oracle as ro
    feats, sd = ro.seeded_fusion_inputs()
    with torch.no_grad():
        out = ro.dinov2_fusion_tail(feats, sd, 36).numpy()
    assert np.abs(out - G["projected"]).max() <= 1e-5 * np.abs(G["projected"]).max()
    assert np.array_equal(ro.grid_spatial(36), G["spatial"])


@pytest.mark.gpu
def test_cuda_tail_matches_reference_fixture():
    from multimodal_classification_b200.dinov2_fusion import DINOv2FusionTail
    from oracle import roi_oracle as ro
    feats, sd = ro.seeded_fusion_inputs()
    tail = DINOv2FusionTail(num_layers=4, hidden_size=1024, output_dim=2048, num_regions=36, device="cuda")
    assert list(tail.state_dict().keys()) == list(sd.keys())
    tail.load_state_dict(sd, strict=True)
    out, spatial = tail.fuse([f.cuda() for f in feats])
    assert out.shape == (1, 36, 2048) and out.dtype == torch.float32
    assert np.array_equal(spatial[0].cpu().numpy(), G["spatial"])
    ref = G["projected"]
    err = np.abs(out[0].cpu().numpy() - ref).max() / np.abs(ref).max()
    assert err <= 2e-2, err                      # bf16 operands, fp32 accumulation (north-star tolerance)
    # CLS-first layout (what the ViT hooks capture, dinov2_multilayer.py:318-321) and a batch of two
    with_cls = [torch.cat([torch.zeros(1, 1, 1024), f], dim=1).repeat(2, 1, 1).cuda() for f in feats]
    out2, _ = tail.fuse(with_cls, has_cls=True)
    assert out2.shape == (2, 36, 2048)
    assert torch.allclose(out2[0], out[0], atol=1e-6) and torch.allclose(out2[1], out[0], atol=1e-6)


@pytest.mark.gpu
def test_bilinear_concat_kernel_vs_interpolate():
    """The gather alone against F.interpolate(bilinear, align_corners=False) on the CPU, exact up to the bf16 rounding of
    the output."""
    import ctypes as C
    from multimodal_classification_b200 import _lib
    g = torch.Generator().manual_seed(1)
    for grid, target, h, layers, b in ((37, 6, 64, 2, 2), (16, 7, 32, 3, 1), (5, 5, 8, 1, 1)):
        feats = [torch.randn(b, grid * grid, h, generator=g) for _ in range(layers)]
        fused = torch.cat(feats, -1)
        ref = torch.nn.functional.interpolate(fused.permute(0, 2, 1).reshape(b, layers * h, grid, grid), size=(target, target),
                                              mode="bilinear", align_corners=False).permute(0, 2, 3, 1).reshape(-1, layers * h)
        dev = [f.cuda() for f in feats]
        out = torch.empty(b * target * target, layers * h, dtype=torch.bfloat16, device="cuda")
        ptrs = (C.c_void_p * layers)(*[f.data_ptr() for f in dev])
        _lib.check(_lib.lib().vb_bilinear_concat(ptrs, layers, out.data_ptr(), b, grid, target, h, grid * grid * h, h, 0,
                                                 torch.cuda.current_stream().cuda_stream), "vb_bilinear_concat")
        assert torch.equal(out.float().cpu(), ref.to(torch.bfloat16).float())


def test_oracle_identity_grid_keeps_every_patch_token():
    """num_regions = the patch count (16 x 16 at 224 px): the reference's resize is the identity, the tail is a per-token MLP."""
    from oracle import roi_oracle as ro
    feats, sd = ro.seeded_fusion_inputs(seed=5, grid=16)
    with torch.no_grad():
        out = ro.dinov2_fusion_tail(feats, sd, 256)
        fused = torch.cat(feats, -1)[0]
        y = torch.nn.functional.linear(fused, sd["projection.0.weight"], sd["projection.0.bias"])
        y = torch.nn.functional.gelu(torch.nn.functional.layer_norm(y, (2048,), sd["projection.1.weight"], sd["projection.1.bias"]))
        want = torch.nn.functional.linear(y, sd["projection.3.weight"], sd["projection.3.bias"])
    assert out.shape == (256, 2048) and torch.allclose(out, want, atol=1e-5)


@pytest.mark.gpu
def test_config4_all_patch_tokens_feed_the_encoder():
    """BASELINE config 4 end to end: 4 layers x (CLS + 16 x 16 patch tokens) x 1024 -> concat -> projection -> 256 regions x
    2048 + grid boxes -> two-stream encoder (blocked attention over 256 regions).  Tail against the oracle tail; encoder
    against the oracle encoder on the tail's own output."""
    from multimodal_classification_b200.dinov2_fusion import DINOv2FusionTail
    from multimodal_classification_b200.vilbert import ViLBERTForClassification
    from oracle import roi_oracle as ro
    from oracle import vilbert_oracle as vo
    feats, sd = ro.seeded_fusion_inputs(seed=5, grid=16)
    tail = DINOv2FusionTail(num_layers=4, hidden_size=1024, output_dim=2048, num_regions=256, device="cuda")
    tail.load_state_dict(sd, strict=True)
    with_cls = [torch.cat([torch.zeros(1, 1, 1024), f], dim=1).repeat(2, 1, 1).cuda() for f in feats]
    out, spatial = tail.fuse(with_cls, has_cls=True)
    assert out.shape == (2, 256, 2048) and spatial.shape == (2, 256, 5)
    with torch.no_grad():
        ref = ro.dinov2_fusion_tail(feats, sd, 256).numpy()
    err = np.abs(out[0].cpu().numpy() - ref).max() / np.abs(ref).max()
    assert err <= 2e-2, err
    assert np.array_equal(spatial[0].cpu().numpy(), ro.grid_spatial(256))
    cfg = vo.tiny_config()
    msd = vo.seeded_state_dict(cfg)
    model = ViLBERTForClassification(cfg, num_labels=2)
    model.load_state_dict(msd, strict=True)
    model = model.cuda().eval()
    batch = vo.synthetic_batch(cfg, batch=2, seq=32, regions=256, seed=3)
    batch["visual_features"], batch["spatial_locations"] = out.cpu(), spatial.cpu()
    with torch.no_grad():
        got = model(**{k: v.cuda() for k, v in batch.items()})
    want, _ = vo.loss_and_grads(msd, cfg, batch)
    scale = want["logits"].abs().max().item()
    assert (got["logits"].float().cpu() - want["logits"]).abs().max().item() <= 2e-2 * scale
    assert abs(got["loss"].item() - want["loss"].item()) <= 1e-3


def test_fusion_tail_host_logic_on_the_cpu(monkeypatch):
    """DINOv2FusionTail.fuse in the GPU-less container: its two direct C-ABI calls (vb_bilinear_concat, vb_gelu_bf16) are
    replaced by numpy restatements working through the pointers / strides fuse() passes, the GEMM and LayerNorm wrappers by
    tests/ops_sim.py.  Against the reference-produced fixture, with and without the CLS token, batch of two."""
    import ctypes as C
    import torch.nn.functional as F
    if torch.cuda.is_available():
        pytest.skip("the stand-ins are for the GPU-less container")
    import ops_sim
    from multimodal_classification_b200 import _lib
    from oracle import roi_oracle as ro

    class FakeLib:
        def vb_bilinear_concat(self, layers, n_layers, out, batch, grid, target, h, batch_stride, token_stride, first_token, stream):
            cols = []
            for l in range(n_layers):
                n = (batch - 1) * batch_stride + (first_token + grid * grid) * token_stride
                raw = np.ctypeslib.as_array((C.c_float * n).from_address(layers[l]))
                x = torch.from_numpy(np.stack([raw[b * batch_stride + first_token * token_stride:][: grid * grid * token_stride]
                                               .reshape(grid * grid, token_stride)[:, :h] for b in range(batch)]))
                r = F.interpolate(x.permute(0, 2, 1).reshape(batch, h, grid, grid), size=(target, target), mode="bilinear",
                                  align_corners=False)
                cols.append(r.permute(0, 2, 3, 1).reshape(batch * target * target, h))
            bits = torch.cat(cols, -1).to(torch.bfloat16).contiguous().view(torch.int16).numpy().view(np.uint16)
            np.ctypeslib.as_array((C.c_uint16 * bits.size).from_address(out))[:] = bits.reshape(-1)
            return 0

        def vb_gelu_bf16(self, x, y, n, stream):
            raw = np.ctypeslib.as_array((C.c_uint16 * n).from_address(x))
            v = torch.from_numpy((raw.astype(np.uint32) << 16).view(np.float32).copy())
            np.ctypeslib.as_array((C.c_uint16 * n).from_address(y))[:] = \
                F.gelu(v).to(torch.bfloat16).view(torch.int16).numpy().view(np.uint16)
            return 0

        def vb_last_error(self):
            return b""

    ops_sim.install(monkeypatch)
    ops_sim.install_device_shims(monkeypatch)
    monkeypatch.setattr(_lib, "lib", lambda: FakeLib())
    from multimodal_classification_b200.dinov2_fusion import DINOv2FusionTail
    feats, sd = ro.seeded_fusion_inputs()
    tail = DINOv2FusionTail(num_layers=4, hidden_size=1024, output_dim=2048, num_regions=36, device="cuda")
    tail.load_state_dict(sd, strict=True)
    out, spatial = tail.fuse(feats)
    ref = G["projected"]
    assert out.shape == (1, 36, 2048) and np.array_equal(spatial[0].numpy(), G["spatial"])
    assert np.abs(out[0].numpy() - ref).max() <= 2e-2 * np.abs(ref).max()
    with_cls = [torch.cat([torch.full((1, 1, 1024), 7.0), f], dim=1).repeat(2, 1, 1).contiguous() for f in feats]
    out2, _ = tail.fuse(with_cls, has_cls=True)
    # (CPU BLAS accumulates differently for 72 and 36 rows; the bf16 intermediates turn that into 1-ulp flips)
    assert out2.shape == (2, 36, 2048) and torch.allclose(out2[0], out[0], atol=5e-3) and torch.allclose(out2[1], out2[0], atol=5e-3)
