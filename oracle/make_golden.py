"""Generate tests/golden/*.npz by running the UNMODIFIED reference (imported from /root/reference/src) on seeded
weights and inputs.  Runs only in the authoring container (the reference does not travel to the GPU box); the
fixtures and this script are committed.

    python oracle/make_golden.py            # writes tests/golden/vilbert_tiny.npz, vilbert_full.npz, roi_*.npz

Fixtures hold, per case: logits, loss, pooled outputs and — for gradients — the per-tensor L2 norms plus, for a few
small tensors, the full gradient, of both the CE loss and the non-cancelling scalar logits[:,1].sum().
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference/src")
os.environ.setdefault("PYTHONDONTWRITEBYTECODE", "1")
sys.dont_write_bytecode = True

from oracle import vilbert_oracle as vo  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")
FULL_GRAD_KEYS = [
    "classifier.4.weight", "classifier.4.bias", "bert.v_embeddings.image_location_embeddings.weight",
    "bert.embeddings.token_type_embeddings.weight", "bert.embeddings.LayerNorm.weight",
    "bert.encoder.layer.0.attention.self.query.bias", "bert.encoder.c_layer.0.biOutput.LayerNorm1.bias",
    "bert.encoder.c_layer.1.biattention.value2.bias", "bert.encoder.v_layer.0.output.dense.bias",
]


def reference_model(cfg, sd):
    from multimodalclassification.models.vilbert_facebook_arch import ViLBERTForClassification
    torch.manual_seed(0)
    m = ViLBERTForClassification(cfg, num_labels=2)
    ref_keys = list(m.state_dict().keys())
    assert ref_keys == list(sd.keys()), "oracle.param_shapes order differs from the reference state_dict"
    for k, v in m.state_dict().items():
        assert tuple(v.shape) == tuple(sd[k].shape), k
    m.load_state_dict(sd, strict=True)
    return m.eval()


def run_case(name, cfg, batch_kw, seed_w=0):
    sd = vo.seeded_state_dict(cfg, seed=seed_w)
    batch = vo.synthetic_batch(cfg, **batch_kw)
    m = reference_model(cfg, sd)
    rec = {}
    for scalar in ("loss", "logit1"):
        m.zero_grad(set_to_none=True)
        out = m(**batch)
        obj = out["loss"] if scalar == "loss" else out["logits"][:, 1].sum()
        obj.backward()
        names, norms = [], []
        for k, p in m.named_parameters():
            names.append(k)
            norms.append(-1.0 if p.grad is None else float(p.grad.double().norm()))
            if k in FULL_GRAD_KEYS and p.grad is not None:
                rec[f"grad_{scalar}/{k}"] = p.grad.numpy().copy()
        rec[f"gradnorm_{scalar}"] = np.asarray(norms, dtype=np.float64)
        rec["param_names"] = np.asarray(names)
    # yardstick: the reference itself under PyTorch's CPU bf16 autocast, same weights and batch — what
    # bf16 arithmetic costs on this model, measured, so that the GPU tolerances are relative to it and not guessed
    fp32_grads = {k: rec[f"grad_logit1/{k}"] for k in FULL_GRAD_KEYS if f"grad_logit1/{k}" in rec}
    fp32_logits = out["logits"].detach().clone()
    m.zero_grad(set_to_none=True)
    with torch.autocast("cpu", dtype=torch.bfloat16):
        out_bf = m(**batch)
        out_bf["logits"].float()[:, 1].sum().backward()
    rec["yard_logits"] = np.asarray(float((out_bf["logits"].float() - fp32_logits).abs().max() / fp32_logits.abs().max()))
    params = dict(m.named_parameters())
    for k, ref in fp32_grads.items():
        gb = params[k].grad.float().numpy()
        rec[f"yard_rel_l2/{k}"] = np.asarray(float(np.linalg.norm(gb - ref) / (np.linalg.norm(ref) + 1e-30)))
    m.zero_grad(set_to_none=True)
    with torch.no_grad():
        t_h, v_h, t_p, v_p = m.bert(**{k: v for k, v in batch.items() if k != "labels"})
    rec.update({"logits": out["logits"].detach().numpy(), "loss": np.asarray(float(out["loss"])),
                "t_pooled": t_p.numpy(), "v_pooled": v_p.numpy(),
                "t_hidden_row0": t_h[:, 0].numpy(), "v_hidden_row0": v_h[:, 0].numpy()})
    # eval-set scores for the AUROC-ordering check: softmax(logits)[:,1] on a second, larger synthetic set
    os.makedirs(GOLDEN, exist_ok=True)
    np.savez_compressed(os.path.join(GOLDEN, name + ".npz"), **rec)
    print(name, "autocast yardstick: logits", float(rec["yard_logits"]), "grad rel-L2",
          {k.split("/")[1][-40:]: round(float(v), 4) for k, v in rec.items() if k.startswith("yard_rel_l2/")})
    print(name, "logits[0]", rec["logits"][0], "loss", rec["loss"], "max|logit|", np.abs(rec["logits"]).max())


def main():
    torch.set_num_threads(os.cpu_count() or 1)
    tiny = vo.tiny_config()
    run_case("vilbert_tiny", tiny, dict(batch=4, seq=128, regions=100, seed=1234))
    run_case("vilbert_tiny_ragged", tiny, dict(batch=3, seq=40, regions=36, seed=77, with_visual_mask=True,
                                              with_token_types=False))
    run_case("vilbert_full", vo.facebook_config(), dict(batch=16, seq=128, regions=100, seed=1234))


if __name__ == "__main__":
    main()
