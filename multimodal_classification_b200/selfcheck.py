"""smoke(): one small forward+backward of the ViLBERT hot path on cuda:0, checked against the oracle
(oracle/ is test infrastructure: it is imported here only as the checker)."""
from __future__ import annotations

import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def compare_with_oracle(cfg, batch_kw, tol_logits=2e-2, check_grads=True, verbose=True):
    if ROOT not in sys.path:
        sys.path.insert(0, ROOT)
    from oracle import vilbert_oracle as vo
    from .vilbert import ViLBERTForClassification
    sd = vo.seeded_state_dict(cfg)
    batch = vo.synthetic_batch(cfg, **batch_kw)
    model = ViLBERTForClassification(cfg, num_labels=2)
    model.load_state_dict(sd, strict=True)
    model = model.cuda().eval()
    dev_batch = {k: v.cuda() for k, v in batch.items()}
    out = model(**dev_batch)
    out["loss"].backward()
    torch.cuda.synchronize()
    ref_out, ref_grads = vo.loss_and_grads(sd, cfg, batch)
    logits = out["logits"].detach().float().cpu()
    scale = ref_out["logits"].abs().max().item()
    err = (logits - ref_out["logits"]).abs().max().item()
    loss_err = abs(out["loss"].item() - ref_out["loss"].item())
    report = {"logit_err_over_max": err / scale, "loss_err": loss_err}
    if verbose:
        print(f"[selfcheck] max|dlogit|/max|logit| = {err / scale:.3e}  |dloss| = {loss_err:.3e}")
    assert err <= tol_logits * scale, (err, scale)
    assert loss_err <= 1e-3, loss_err
    if check_grads:
        worst = (1.0, "")
        for k, p in model.named_parameters():
            g_ref = ref_grads[k]
            if g_ref is None:
                assert p.grad is None, k
                continue
            assert p.grad is not None, k
            if ".key" in k and k.endswith(".bias"):
                assert p.grad.abs().max().item() < 1e-3, k   # mathematically zero
                continue
            g = p.grad.float().cpu().flatten().double()
            r = g_ref.flatten().double()
            cos = float((g @ r) / (g.norm() * r.norm() + 1e-30))
            if cos < worst[0]:
                worst = (cos, k)
        report["worst_grad_cosine"] = worst
        if verbose:
            print(f"[selfcheck] worst gradient cosine vs fp32 oracle: {worst[0]:.5f} ({worst[1]})")
        assert worst[0] >= 0.97, worst
    return report


def compare_roi_stage(verbose=True):
    """ResNet-152 conv1..layer3 + RoIPool + layer4 on one small seeded image against the fp32 oracle (bit-exact boxes,
    bf16 tolerance on the features)."""
    from oracle import roi_oracle as ro
    from .resnet152_roi import ResNet152ROIExtractor
    import numpy as np
    sd = ro.seeded_backbone_state(0)
    ext = ResNet152ROIExtractor(device="cuda", weights=None, image_size=224)
    ext.backbone.load_state_dict(sd, strict=True)
    g = torch.Generator().manual_seed(5)
    img = torch.randn(1, 3, 224, 224, generator=g)
    feats, spatial = ext.extract_batch(img.cuda())
    ref_f, ref_s, ref_boxes = ro.extract_features(sd, img)
    assert np.array_equal(ext._generate_proposals(224, 224).cpu().numpy(), ref_boxes)
    assert np.array_equal(spatial[0].cpu().numpy(), ref_s)
    err = np.abs(feats[0].cpu().numpy() - ref_f).max() / np.abs(ref_f).max()
    if verbose:
        print(f"[selfcheck] RoI stage: boxes bit-exact, max|dfeat|/max|feat| = {err:.3e}")
    assert err <= 2e-2, err
    return err


def compare_ingest(verbose=True):
    """One LMDB-shaped batch through vb_lmdb_regions against the oracle: spatial rows bit-exact, features = RNE bf16."""
    from oracle import ingest_oracle as io
    from . import ops
    import numpy as np
    rng = np.random.default_rng(3)
    feat = np.abs(rng.standard_normal((4 * 100, 2048))).astype(np.float32)
    boxes = rng.uniform(0, 1000, (4 * 100, 4)).astype(np.float32)
    f16 = torch.empty(400, 2048, dtype=torch.bfloat16, device="cuda")
    spatial = torch.empty(400, 5, device="cuda")
    ops.lmdb_regions(torch.from_numpy(feat).cuda(), f16, torch.from_numpy(boxes).cuda(), spatial)
    assert np.array_equal(spatial.cpu().numpy(), io.process_boxes(boxes, 400))
    assert torch.equal(f16.cpu().view(torch.int16), torch.from_numpy(feat).to(torch.bfloat16).view(torch.int16))
    if verbose:
        print("[selfcheck] ingest: spatial rows and bf16 features bit-exact")


def smoke():
    if not torch.cuda.is_available():
        raise RuntimeError("smoke() needs a CUDA device")
    if ROOT not in sys.path:
        sys.path.insert(0, ROOT)
    from oracle import vilbert_oracle as vo
    torch.cuda.set_device(0)
    compare_with_oracle(vo.tiny_config(), dict(batch=4, seq=128, regions=100, seed=1234))
    compare_roi_stage()
    compare_ingest()
    print("[selfcheck] smoke OK")
