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


def smoke():
    if not torch.cuda.is_available():
        raise RuntimeError("smoke() needs a CUDA device")
    if ROOT not in sys.path:
        sys.path.insert(0, ROOT)
    from oracle import vilbert_oracle as vo
    torch.cuda.set_device(0)
    compare_with_oracle(vo.tiny_config(), dict(batch=4, seq=128, regions=100, seed=1234))
    print("[selfcheck] smoke OK")
