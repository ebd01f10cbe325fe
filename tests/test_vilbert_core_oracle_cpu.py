"""Oracle of the reference's second two-stream surface (models/vilbert_core.py) against the fixture the
reference class itself produced (oracle/make_golden_core.py): logits, loss, pooled and sequence outputs, every parameter's
gradient norm, one full gradient tensor, and the optional-argument path.  CPU only: the CUDA engine for this surface is the
next step; the oracle comes first."""
import os

import numpy as np
import torch

from oracle import vilbert_core_oracle as co

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "vilbert_core_tiny.npz"))


def _seeded(cfg):
    """Re-draw seeded_core_state over the reference's full key list (shapes of the dead BertModel layers included)."""
    shapes = co.param_shapes(cfg)
    h, inter = cfg["hidden_size"], cfg["intermediate_size"]
    full = {}
    for k in (str(x) for x in G["state_keys"]):
        if k in shapes:
            full[k] = shapes[k]
        elif k.startswith("vilbert.bert.encoder.layer."):
            tail = k.split(".", 5)[5]
            dims = {"attention.self.query": (h, h), "attention.self.key": (h, h), "attention.self.value": (h, h),
                    "attention.output.dense": (h, h), "attention.output.LayerNorm": (h,), "intermediate.dense": (inter, h),
                    "output.dense": (h, inter), "output.LayerNorm": (h,)}[tail.rsplit(".", 1)[0]]
            full[k] = dims if k.endswith(".weight") else (dims[0],)
        elif k.startswith("vilbert.bert.pooler.dense"):
            full[k] = (h, h) if k.endswith(".weight") else (h,)
        else:
            raise AssertionError(f"unexpected reference key {k}")
    return co.seeded_core_state(full, seed=0)


def test_oracle_matches_reference_outputs_and_gradients():
    cfg = co.tiny_core_config()
    sd = _seeded(cfg)
    batch = co.synthetic_batch(cfg, batch=4, seq=32, regions=20, seed=1234)
    out, grads = co.loss_and_grads(sd, cfg, batch)
    for k in ("logits", "pooled_output", "text_pooled", "visual_pooled"):
        assert np.abs(out[k].numpy() - G[k]).max() <= 2e-5, k
    assert abs(out["loss"].item() - float(G["loss"])) <= 1e-6
    assert np.abs(out["text_output"][:, ::8, ::16].numpy() - G["text_output_probe"]).max() <= 5e-5
    assert np.abs(out["visual_output"][:, ::4, ::16].numpy() - G["visual_output_probe"]).max() <= 5e-5
    names, norms = [str(n) for n in G["names"]], G["grad_norms"]
    dead = [n for n, v in zip(names, norms) if v < 0]
    assert len(dead) == 18 and all(n.startswith(("vilbert.bert.encoder.", "vilbert.bert.pooler.")) for n in dead)
    for n, v in zip(names, norms):
        if v < 0:
            assert n not in grads
            continue
        g = grads[n]
        assert g is not None, n
        assert abs(float(g.norm()) - v) <= 1e-4 * max(v, 1e-6) + 1e-7, (n, float(g.norm()), v)
    probe = str(G["grad_probe_name"])
    assert np.abs(grads[probe].numpy() - G["grad_probe"]).max() <= 1e-6 + 1e-4 * np.abs(G["grad_probe"]).max()


def test_oracle_optional_arguments_path():
    cfg = co.tiny_core_config()
    sd = _seeded(cfg)
    b2 = co.synthetic_batch(cfg, batch=3, seq=20, regions=12, seed=7)
    for k in ("token_type_ids", "visual_attention_mask", "spatial_locations"):
        b2.pop(k)
    with torch.no_grad():
        out = co.forward(sd, cfg, **b2)
    assert np.abs(out["logits"].numpy() - G["logits_minimal"]).max() <= 2e-5
    assert abs(out["loss"].item() - float(G["loss_minimal"])) <= 1e-6


def test_used_parameter_shapes_are_a_subset_of_the_reference_state_dict():
    cfg = co.tiny_core_config()
    shapes = co.param_shapes(cfg)
    keys = {str(k) for k in G["state_keys"]}
    assert set(shapes) <= keys
    assert all(k.startswith(("vilbert.bert.encoder.", "vilbert.bert.pooler.")) or k in shapes for k in keys)
    # at the reference's full size (bert-base text stream): parameters the forward pass reads
    assert sum(int(np.prod(s)) for s in co.param_shapes(co.core_config()).values()) == 240_493_058
