"""CPU twin of tests/test_training_loop_gpu.py: the reference's training loop (nodes.py:757-760, 784-799) drives the drop-in
module over the functional kernel stand-ins of tests/ops_sim.py, so the HOST side of row a24 — stock AdamW updating views of the
flat master buffer in place, the version check that re-casts the bf16 shadows, clip_grad_norm_ over the gradient views, unused
parameters skipped — is covered in the GPU-less container."""
import numpy as np
import pytest
import torch

from oracle import vilbert_oracle as vo


@pytest.fixture
def simulated(monkeypatch):
    if torch.cuda.is_available():
        pytest.skip("the stand-ins are for the GPU-less container")
    import ops_sim
    ops_sim.install(monkeypatch)


def test_reference_loop_over_stand_ins(simulated):
    import test_training_loop_gpu as T
    from multimodal_classification_b200.vilbert import ViLBERTForClassification
    cfg = vo.tiny_config()
    sd = vo.seeded_state_dict(cfg)
    batches = T._batches(cfg)[:3]
    oracle = T._OracleModel(sd, cfg).eval()
    want = T._reference_loop(oracle, batches, torch.device("cpu"))
    model = ViLBERTForClassification(cfg, num_labels=2)
    model.load_state_dict(sd, strict=True)
    model.eval()
    got = T._reference_loop(model, batches, torch.device("cpu"))
    assert np.abs(np.array(got) - np.array(want)).max() <= 4e-3, (got, want)
    key = "bert.encoder.layer.0.intermediate.dense.weight"
    master = dict(model.named_parameters())[key].detach()
    assert not torch.equal(master, sd[key])
    with torch.no_grad():
        model(**batches[0])
    assert torch.equal(model._engine.flat.w(key), master.to(torch.bfloat16))      # shadow followed the in-place update
    for k, p in model.named_parameters():
        if "q_dense" in k:
            assert p.grad is None and torch.equal(p.detach(), sd[k]), k


def test_index_checks_and_zero_grad_fast_path_over_stand_ins(simulated):
    from multimodal_classification_b200._lib import VbError
    from multimodal_classification_b200.vilbert import ViLBERTForClassification
    cfg = vo.tiny_config()
    sd = vo.seeded_state_dict(cfg)
    model = ViLBERTForClassification(cfg, num_labels=2)
    model.load_state_dict(sd, strict=True)
    model.eval()
    b = vo.synthetic_batch(cfg, batch=2, seq=16, regions=8, seed=5)
    key = "bert.encoder.layer.1.output.dense.weight"
    p = dict(model.named_parameters())[key]
    model(**b)["loss"].backward()
    g1 = p.grad.clone()
    model.zero_grad(set_to_none=False)
    assert model._engine.grads_clean and float(p.grad.abs().max()) == 0.0
    model(**b)["loss"].backward()
    assert torch.equal(p.grad, g1)
    model(**b)["loss"].backward()
    assert torch.allclose(p.grad, 2 * g1, rtol=1e-6, atol=0)
    bad = dict(b)
    bad["labels"] = b["labels"].clone()
    bad["labels"][0] = 7
    with torch.no_grad():
        model(**bad)                                  # flagged by the staging pass ...
        with pytest.raises(VbError, match="labels outside"):
            model(**b)                                # ... raised when the next batch arrives
    ign = dict(b)
    ign["labels"] = b["labels"].clone()
    ign["labels"][0] = -100
    out = model(**ign)
    ref, _ = vo.loss_and_grads(sd, cfg, ign)
    assert abs(out["loss"].item() - ref["loss"].item()) <= 1e-3
    t = cfg["max_position_embeddings"] + 8
    with pytest.raises(VbError, match="max_position_embeddings"):
        model(**vo.synthetic_batch(cfg, batch=2, seq=t, regions=8, seed=3))
