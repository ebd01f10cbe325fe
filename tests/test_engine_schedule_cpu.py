"""The ViLBERT engine's HOST schedule (multimodal_classification_b200/vilbert.py: forward and backward kernel sequence, flat
parameter / gradient layout, fp32 residual ring, bucket order) run in the GPU-less container over the functional stand-ins of
tests/ops_sim.py and compared with the fp32 oracle: logits, loss and every parameter gradient.  The kernels themselves are
tested on the B200 (tests/test_*_gpu.py); this covers what sits above them."""
import pytest
import torch

from oracle import vilbert_oracle as vo


@pytest.fixture
def simulated(monkeypatch):
    if torch.cuda.is_available():
        pytest.skip("the stand-ins are for the GPU-less container")
    import ops_sim
    ops_sim.install(monkeypatch)


def _run(cfg, batch_kw, train=False):
    from multimodal_classification_b200.vilbert import ViLBERTForClassification
    sd = vo.seeded_state_dict(cfg)
    batch = vo.synthetic_batch(cfg, **batch_kw)
    model = ViLBERTForClassification(cfg, num_labels=2)
    model.load_state_dict(sd, strict=True)
    model.eval()
    out = model(**batch)
    out["loss"].backward()
    ref_out, ref_grads = vo.loss_and_grads(sd, cfg, batch)
    return model, out, ref_out, ref_grads


@pytest.mark.parametrize("batch_kw", [dict(batch=2, seq=24, regions=12, seed=3),
                                      dict(batch=3, seq=16, regions=9, seed=5, with_visual_mask=True, with_token_types=False)])
def test_engine_schedule_matches_oracle(simulated, batch_kw):
    cfg = vo.tiny_config()
    model, out, ref_out, ref_grads = _run(cfg, batch_kw)
    scale = ref_out["logits"].abs().max().item()
    assert (out["logits"].float() - ref_out["logits"]).abs().max().item() <= 2e-2 * scale
    assert abs(out["loss"].item() - ref_out["loss"].item()) <= 1e-3
    worst = (1.0, "")
    for k, p in model.named_parameters():
        g_ref = ref_grads[k]
        if g_ref is None:
            assert p.grad is None, k
            continue
        assert p.grad is not None, k
        if ".key" in k and k.endswith(".bias"):
            assert p.grad.abs().max().item() < 1e-3, k
            continue
        g, r = p.grad.flatten().double(), g_ref.flatten().double()
        cos = float((g @ r) / (g.norm() * r.norm() + 1e-30))
        worst = min(worst, (cos, k))
    assert worst[0] >= 0.97, worst
