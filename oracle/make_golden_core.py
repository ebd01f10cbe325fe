"""Generate tests/golden/vilbert_core_tiny.npz by running the UNMODIFIED reference class
``models/vilbert_core.py::ViLBERTForClassification`` in the authoring container.  The only substitution:
``BertModel.from_pretrained`` (no checkpoint offline) is replaced by a ``BertModel`` built from a ``BertConfig`` of the tiny
sizes; the module then copies those sizes into its own config exactly as it does for bert-base (vilbert_core.py:500-507).
Weights come from ``oracle.vilbert_core_oracle.seeded_core_state`` over the reference module's own state_dict shapes."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference/src")
from oracle import vilbert_core_oracle as co  # noqa: E402


def build_reference(cfg):
    import transformers
    import multimodalclassification.models.vilbert_core as ref
    bcfg = transformers.BertConfig(vocab_size=cfg["vocab_size"], hidden_size=cfg["hidden_size"], num_hidden_layers=1,
                                   num_attention_heads=cfg["num_attention_heads"], intermediate_size=cfg["intermediate_size"],
                                   max_position_embeddings=cfg["max_position_embeddings"], type_vocab_size=cfg["type_vocab_size"])

    class Stub:
        @staticmethod
        def from_pretrained(name):
            return transformers.BertModel(bcfg)
    ref.BertModel = Stub
    ref_cfg = {k: cfg[k] for k in ("v_feature_size", "v_num_hidden_layers", "max_regions", "t_num_hidden_layers", "num_co_layers",
                                   "classifier_dropout", "num_labels")}
    torch.manual_seed(0)
    return ref.ViLBERTForClassification(ref_cfg, num_labels=2).eval()


def main():
    cfg = co.tiny_core_config()
    model = build_reference(cfg)
    shapes = {k: tuple(v.shape) for k, v in model.state_dict().items() if v.is_floating_point()}
    sd = co.seeded_core_state(shapes, seed=0)
    print(model.load_state_dict(sd, strict=False))
    batch = co.synthetic_batch(cfg, batch=4, seq=32, regions=20, seed=1234)
    out = model(**batch)
    out["loss"].backward()
    named = dict(model.named_parameters())
    names = sorted(named)
    norms = np.array([-1.0 if named[k].grad is None else float(named[k].grad.norm()) for k in names])
    probe = "vilbert.encoder.c_layer.1.biattention_v.self.key.weight"
    res = {"names": np.array(names), "grad_norms": norms, "grad_probe": named[probe].grad.numpy(), "grad_probe_name": np.array(probe),
           "state_keys": np.array(list(model.state_dict().keys()))}
    for k in ("logits", "loss", "pooled_output", "text_pooled", "visual_pooled"):
        res[k] = out[k].detach().numpy()
    res["text_output_probe"] = out["text_output"].detach()[:, ::8, ::16].numpy()
    res["visual_output_probe"] = out["visual_output"].detach()[:, ::4, ::16].numpy()
    # a second batch without token types / visual mask / locations (all optional in the reference signature)
    b2 = co.synthetic_batch(cfg, batch=3, seq=20, regions=12, seed=7)
    for k in ("token_type_ids", "visual_attention_mask", "spatial_locations"):
        b2.pop(k)
    with torch.no_grad():
        o2 = model(**b2)
    res["logits_minimal"], res["loss_minimal"] = o2["logits"].numpy(), o2["loss"].numpy()
    path = os.path.join(ROOT, "tests", "golden", "vilbert_core_tiny.npz")
    np.savez_compressed(path, **res)
    unused = [k for k, n in zip(names, norms) if n < 0]
    print("wrote", path, os.path.getsize(path) // 1024, "KiB; logits", res["logits"].round(4).tolist(), "loss", float(res["loss"]),
          "; parameters without gradient:", len(unused), "of", len(names))


if __name__ == "__main__":
    main()
