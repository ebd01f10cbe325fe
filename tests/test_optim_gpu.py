"""Fused clip + AdamW + shadow refresh against torch.nn.utils.clip_grad_norm_ + torch.optim.AdamW (the reference's step,
pipelines/model_training/nodes.py:757-760, 795-799) on the same gradients, fp32, three steps with a changing LR."""

import pytest
import torch

pytestmark = pytest.mark.gpu


def _setup(freeze=0):
    from multimodal_classification_b200.vilbert import ViLBERTForClassification
    from oracle import vilbert_oracle as vo
    cfg = vo.tiny_config()
    torch.manual_seed(0)
    model = ViLBERTForClassification(cfg, num_labels=2).cuda().eval()
    if freeze:
        model.freeze_bert_layers(freeze)
    batch = {k: v.cuda() for k, v in vo.synthetic_batch(cfg, batch=4, seq=32, regions=20, seed=3).items()}
    return model, batch


@pytest.mark.parametrize("freeze,max_norm", [(0, 1.0), (2, 0.05), (0, None)])
def test_fused_adamw_matches_torch(freeze, max_norm):
    from multimodal_classification_b200.optim import FusedAdamW
    model, batch = _setup(freeze)
    opt = FusedAdamW(model, lr=1e-3, weight_decay=0.01, max_grad_norm=max_norm)
    sched = torch.optim.lr_scheduler.LambdaLR(opt, lambda s: 1.0 / (1 + s))
    # the yardstick: stock torch on detached copies of the same parameters, fed the SAME gradients every step
    ref_params = {k: p.detach().clone().requires_grad_(p.requires_grad) for k, p in model.named_parameters()}
    ref_opt = torch.optim.AdamW([p for p in ref_params.values()], lr=1e-3, weight_decay=0.01)
    ref_sched = torch.optim.lr_scheduler.LambdaLR(ref_opt, lambda s: 1.0 / (1 + s))
    for step in range(3):
        model.zero_grad(set_to_none=True)
        model(**batch)["loss"].backward()
        for k, p in model.named_parameters():
            ref_params[k].grad = None if p.grad is None else p.grad.detach().clone()
        with_grad = [p for p in ref_params.values() if p.grad is not None]
        if max_norm is not None:
            want_norm = torch.nn.utils.clip_grad_norm_(with_grad, max_norm).item()
        ref_opt.step(); ref_sched.step()
        opt.step(); sched.step()
        if max_norm is not None:
            assert abs(opt.grad_norm() - want_norm) <= 1e-5 * want_norm
        worst = 0.0
        for k, p in model.named_parameters():
            r = ref_params[k]
            worst = max(worst, ((p.detach() - r).abs().max() / (r.abs().max() + 1e-12)).item())
        assert worst <= 2e-6, (step, worst)
    flat = model._engine.flat
    # the bf16 shadow of the GEMM weights was rewritten by the same kernel
    assert torch.equal(flat.shadow, flat.master[:flat.w_end].to(torch.bfloat16))
    # parameters without a gradient did not move (frozen layers, unused q_dense*): identical to their initial values
    torch.manual_seed(0)
    from multimodal_classification_b200.vilbert import ViLBERTForClassification
    fresh = ViLBERTForClassification(model.config, num_labels=2)
    for (k, p), (_, q) in zip(model.named_parameters(), fresh.named_parameters()):
        if "q_dense" in k or not p.requires_grad:
            assert torch.equal(p.detach().cpu(), q.detach()), k
    # and the next forward uses the updated weights without a separate shadow refresh
    out = model(**batch)
    assert torch.isfinite(out["loss"]).item()
