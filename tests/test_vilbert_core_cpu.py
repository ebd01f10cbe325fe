"""The vilbert_core surface (multimodal_classification_b200/vilbert_core.py) on the engine: parameter
tree = the reference's state_dict (authoring container), and the schedule — position table through a one-hot GEMM operand,
mean-pooled visual stream, cross-attention blocks under the engine's co-attention names — against the pinned oracle over the
functional kernel stand-ins of tests/ops_sim.py.  First GPU parity run pending (see the module docstring)."""
import os

import numpy as np
import pytest
import torch

from oracle import vilbert_core_oracle as co

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "vilbert_core_tiny.npz"))


def _bert_config(cfg):
    import transformers
    return transformers.BertConfig(vocab_size=cfg["vocab_size"], hidden_size=cfg["hidden_size"], num_hidden_layers=1,
                                   num_attention_heads=cfg["num_attention_heads"], intermediate_size=cfg["intermediate_size"],
                                   max_position_embeddings=cfg["max_position_embeddings"], type_vocab_size=cfg["type_vocab_size"])


def _model(cfg):
    from multimodal_classification_b200.vilbert_core import ViLBERTForClassification
    ref_cfg = {k: cfg[k] for k in ("v_feature_size", "v_num_hidden_layers", "max_regions", "t_num_hidden_layers", "num_co_layers",
                                   "classifier_dropout", "num_labels")}
    return ViLBERTForClassification(ref_cfg, num_labels=2, bert_config=_bert_config(cfg))


def test_state_dict_is_the_reference_state_dict():
    cfg = co.tiny_core_config()
    model = _model(cfg)
    own = {k: tuple(v.shape) for k, v in model.state_dict().items() if v.is_floating_point()}
    assert sorted(own) == sorted(str(k) for k in G["state_keys"])
    used = co.param_shapes(cfg)
    assert all(own[k] == used[k] for k in used)
    fb = dict(model._engine_named_parameters())
    assert len(fb) == len(dict(model.named_parameters()))                       # a renaming, nothing lost or duplicated
    assert sum(k.startswith("unused.") for k in fb) == 18                       # the text BertModel's encoder layer + pooler


@pytest.fixture
def simulated(monkeypatch):
    if torch.cuda.is_available():
        pytest.skip("the stand-ins are for the GPU-less container")
    import ops_sim
    ops_sim.install(monkeypatch)


def _seeded_state(model):
    shapes = {k: tuple(v.shape) for k, v in model.state_dict().items() if v.is_floating_point()}
    ordered = {str(k): shapes[str(k)] for k in G["state_keys"]}                 # the fixture's draw order
    return co.seeded_core_state(ordered, seed=0)


def test_core_schedule_matches_the_pinned_oracle(simulated):
    cfg = co.tiny_core_config()
    model = _model(cfg)
    sd = _seeded_state(model)
    model.load_state_dict(sd, strict=False)
    model.eval()
    batch = co.synthetic_batch(cfg, batch=4, seq=32, regions=20, seed=1234)
    out = model(**batch)
    out["loss"].backward()
    # against the fixture the reference class itself produced ...
    scale = np.abs(G["logits"]).max()
    assert np.abs(out["logits"].detach().numpy() - G["logits"]).max() <= 2e-2 * scale
    assert abs(out["loss"].item() - float(G["loss"])) <= 1e-3
    # ... and every gradient against the oracle
    _, grads = co.loss_and_grads(sd, cfg, batch)
    worst = (1.0, "")
    for k, p in model.named_parameters():
        if k not in grads:
            assert p.grad is None, k                                            # dead weights of the text BertModel
            continue
        assert p.grad is not None, k
        g, r = p.grad.flatten().double(), grads[k].flatten().double()
        if ".key." in k and k.endswith(".bias"):
            assert g.abs().max().item() < 1e-3, k                               # mathematically zero
            continue
        worst = min(worst, (float((g @ r) / (g.norm() * r.norm() + 1e-30)), k))
        assert abs(float(g.norm()) - float(r.norm())) <= 0.15 * float(r.norm()) + 1e-7, k
    assert worst[0] >= 0.97, worst
    # region-position rows beyond the batch's region count receive nothing
    pos = model.vilbert.visual_embeddings.position_embeddings.weight.grad
    assert pos[20:].abs().max().item() == 0 and pos[:20].abs().max().item() > 0


def test_core_surface_helpers_and_limits(simulated):
    from multimodal_classification_b200._lib import VbError
    cfg = co.tiny_core_config()
    model = _model(cfg)
    model.load_state_dict(_seeded_state(model), strict=False)
    model.eval()
    total, trainable = model.get_num_parameters()
    model.freeze_bert_layers(2)
    assert model.get_num_parameters()[1] < trainable == total
    batch = co.synthetic_batch(cfg, batch=2, seq=16, regions=10, seed=3)
    with torch.no_grad():
        logits = model(**batch)["logits"]
    assert torch.equal(model.predict(logits), logits.argmax(-1)) and torch.allclose(model.predict_proba(logits).sum(-1), torch.ones(2))
    too_many = co.synthetic_batch(cfg, batch=1, seq=8, regions=cfg["max_regions"] + 1, seed=3)
    with pytest.raises(VbError, match="region-position table"):
        model(**too_many)


def test_new_gemm_shapes_pass_the_library_argument_checks():
    """The three GEMM calls this surface adds (one-hot operands with padded row strides, K = region count / batch size) are
    legal for the real C-ABI library: its validation and tile selection accept them and it fails only where the CUDA driver is
    first needed (tensor-map encoding), exactly like an ordinary Linear; a misaligned row stride is rejected before that."""
    import ctypes as C
    from multimodal_classification_b200 import _lib
    if torch.cuda.is_available():
        pytest.skip("argument-only probe with fake pointers: GPU-less container only")
    lib = _lib.lib()

    def probe(**kw):
        a = _lib.GemmArgs()
        for k, v in kw.items():
            setattr(a, k, v)
        return lib.vb_gemm_bf16(C.byref(a), None), lib.vb_last_error()
    P, B, R, H = 1 << 20, 16, 100, 768
    Mv, Rp, Bp = B * R, 104, 16
    calls = {
        "position add": dict(a=P, b=P, d=P, aux=P, lda=Rp, ldb=H, ldd=H, ld_aux=H, m=Mv, n=H, k=R, b_mn_major=1, aux_mode=1),
        "position wgrad": dict(a=P, b=P, d=P, lda=Rp, ldb=H, ldd=H, m=R, n=H, k=Mv, a_mn_major=1, b_mn_major=1, d_is_f32=1),
        "mean-pool broadcast": dict(a=P, b=P, d=P, scale=P, lda=Bp, ldb=H, ldd=H, m=Mv, n=H, k=B, b_mn_major=1),
        "ordinary linear": dict(a=P, b=P, d=P, bias=P, lda=768, ldb=768, ldd=3072, m=2048, n=3072, k=768),
    }
    for name, kw in calls.items():
        rc, msg = probe(**kw)
        assert rc != 0 and b"cuTensorMapEncodeTiled" in msg, (name, rc, msg)
    rc, msg = probe(a=P, b=P, d=P, lda=100, ldb=H, ldd=H, m=Mv, n=H, k=R, b_mn_major=1)
    assert rc == -1 and b"lda/ldb" in msg
