"""B200 parity of the vilbert_core surface (multimodal_classification_b200/vilbert_core.py; reference
models/vilbert_core.py:593-657) against the reference-made fixture and the pinned oracle.  Its host schedule is also verified in
the CPU suite (tests/test_vilbert_core_cpu.py)."""
import os

import numpy as np
import pytest
import torch

from oracle import vilbert_core_oracle as co

pytestmark = pytest.mark.gpu
G = np.load(os.path.join(os.path.dirname(__file__), "golden", "vilbert_core_tiny.npz"))


def test_core_surface_against_reference_fixture_and_oracle():
    from test_vilbert_core_cpu import _model, _seeded_state
    cfg = co.tiny_core_config()
    model = _model(cfg)
    sd = _seeded_state(model)
    model.load_state_dict(sd, strict=False)
    model = model.cuda().eval()
    batch = co.synthetic_batch(cfg, batch=4, seq=32, regions=20, seed=1234)
    results = []
    for _ in range(3):                                   # eager, capture + replay, replay
        model.zero_grad(set_to_none=True)
        out = model(**{k: v.cuda() for k, v in batch.items()})
        out["loss"].backward()
        results.append((out["logits"].detach().clone(), model.vilbert.visual_embeddings.position_embeddings.weight.grad.clone()))
    assert all(torch.equal(r[0], results[0][0]) and torch.equal(r[1], results[0][1]) for r in results[1:])
    assert np.abs(results[0][0].float().cpu().numpy() - G["logits"]).max() <= 2e-2 * np.abs(G["logits"]).max()
    assert abs(out["loss"].item() - float(G["loss"])) <= 1e-3
    _, grads = co.loss_and_grads(sd, cfg, batch)
    worst = (1.0, "")
    for k, p in model.named_parameters():
        if k not in grads or (".key." in k and k.endswith(".bias")):
            continue
        g, r = p.grad.flatten().double().cpu(), grads[k].flatten().double()
        worst = min(worst, (float((g @ r) / (g.norm() * r.norm() + 1e-30)), k))
    assert worst[0] >= 0.97, worst
