"""Host logic of the fused optimizer (multimodal_classification_b200/optim.py: parameter ranges with gradients, pointer and
length plumbing over the flat buffers, optimizer-state layout, the "shadows are current" hand-over to the engine) in the
GPU-less container: the two C-ABI entry points are replaced by numpy restatements working on host memory through the very
pointers optim.py passes (formulas of csrc/optim.cu = torch's single-tensor AdamW), the engine runs over tests/ops_sim.py.
Three training steps must leave the parameters where clip_grad_norm_ + torch.optim.AdamW leave an identical model."""
import ctypes as C

import numpy as np
import pytest
import torch

from oracle import vilbert_oracle as vo


def _f32(ptr, n):
    return np.ctypeslib.as_array((C.c_float * n).from_address(ptr))


class FakeOptimLib:
    def __init__(self):
        self.ranges = []

    def vb_grad_sumsq(self, g, n, acc, stream):
        np.ctypeslib.as_array((C.c_double * 1).from_address(acc))[0] += float((_f32(g, n).astype(np.float64) ** 2).sum())
        return 0

    def vb_adamw_step(self, ref, stream):
        a = ref._obj
        self.ranges.append((a.n, a.shadow_n))
        p, g, m, v = (_f32(x, a.n) for x in (a.param, a.grad, a.exp_avg, a.exp_avg_sq))
        coef = np.float32(1.0)
        if a.max_norm > 0:
            total = np.float32(np.sqrt(np.ctypeslib.as_array((C.c_double * 1).from_address(a.grad_sumsq))[0]))
            coef = np.float32(min(a.max_norm / (total + np.float32(1e-6)), 1.0))
        gg = g * coef
        p *= np.float32(1.0 - a.lr * a.weight_decay)
        m += (gg - m) * np.float32(1.0 - a.beta1)
        v[:] = v * np.float32(a.beta2) + np.float32(1.0 - a.beta2) * gg * gg
        bc1 = np.float32(1.0 - a.beta1 ** a.step)
        bc2_sqrt = np.float32(np.sqrt(1.0 - a.beta2 ** a.step))
        p -= np.float32(a.lr) / bc1 * (m / (np.sqrt(v) / bc2_sqrt + np.float32(a.eps)))
        if a.shadow_n:
            bits = torch.from_numpy(p[: a.shadow_n].copy()).to(torch.bfloat16).view(torch.int16).numpy().view(np.uint16)
            np.ctypeslib.as_array((C.c_uint16 * a.shadow_n).from_address(a.shadow))[:] = bits
        return 0

    def vb_last_error(self):
        return b""


@pytest.fixture
def simulated(monkeypatch):
    if torch.cuda.is_available():
        pytest.skip("the stand-ins are for the GPU-less container")
    import ops_sim
    from multimodal_classification_b200 import _lib
    ops_sim.install(monkeypatch)
    fake = FakeOptimLib()
    monkeypatch.setattr(_lib, "lib", lambda: fake)
    return fake


def _model(cfg, sd):
    from multimodal_classification_b200.vilbert import ViLBERTForClassification
    m = ViLBERTForClassification(cfg, num_labels=2)
    m.load_state_dict(sd, strict=True)
    return m.eval()


def test_fused_adamw_tracks_torch_adamw_with_clipping(simulated):
    """Three training steps of the engine + FusedAdamW; after every backward the same gradients are also given to
    clip_grad_norm_ + torch.optim.AdamW acting on a detached copy of the parameters (comparing two separately simulated
    models instead would be ill-conditioned: Adam turns 1-ulp bf16 differences into O(lr) parameter differences)."""
    from multimodal_classification_b200.optim import FusedAdamW
    cfg = vo.tiny_config()
    sd = vo.seeded_state_dict(cfg)
    ours = _model(cfg, sd)
    ours.freeze_bert_layers(1)
    names = [k for k, p in ours.named_parameters() if p.requires_grad and "q_dense" not in k]
    hyper = dict(lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.01)
    fused = FusedAdamW(ours, max_grad_norm=0.05, **hyper)
    twins = {k: torch.nn.Parameter(sd[k].clone()) for k in names}
    stock = torch.optim.AdamW(list(twins.values()), **hyper)
    for i in range(3):
        ours.zero_grad(set_to_none=True)
        ours(**vo.synthetic_batch(cfg, batch=2, seq=16, regions=8, seed=30 + i))["loss"].backward()
        live = dict(ours.named_parameters())
        assert all(live[k].grad is not None for k in names)
        for k in names:
            twins[k].grad = live[k].grad.clone()
        norm = torch.nn.utils.clip_grad_norm_(list(twins.values()), 0.05)
        stock.step()
        fused.step()
        assert abs(fused.grad_norm() - float(norm)) <= 1e-4 * float(norm) and float(norm) > 0.05     # the clip is active
        for k in names:
            assert torch.allclose(live[k], twins[k], rtol=0, atol=2e-6), (i, k, (live[k] - twins[k]).abs().max().item())
    moved = max((dict(ours.named_parameters())[k] - sd[k]).abs().max().item() for k in names)
    assert moved > 1e-3
    for k, p in ours.named_parameters():
        if k not in names:
            assert torch.equal(p, sd[k]), k                        # frozen / unused parameters are not touched (no decay either)
    # the step wrote current bf16 shadows itself and told the engine so: the next forward must not need a recast
    flat = ours._engine.flat
    assert flat._version == flat.versions()
    assert torch.equal(flat.shadow, flat.master[: flat.w_end].to(torch.bfloat16))
    assert len(simulated.ranges) >= 3 and all(sn <= n for n, sn in simulated.ranges)


def test_fused_adamw_before_first_step_is_an_error(simulated):
    from multimodal_classification_b200._lib import VbError
    from multimodal_classification_b200.optim import FusedAdamW
    cfg = vo.tiny_config()
    with pytest.raises(VbError, match="before the first forward"):
        FusedAdamW(_model(cfg, vo.seeded_state_dict(cfg))).step()
