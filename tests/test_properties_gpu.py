"""Size-independent properties at BASELINE.json's full configuration (ViLBERT-base, 128 tokens x 100 regions), and the
AUROC-ordering check of the north-star against logits the reference itself produced on a fixed 512-sample synthetic eval set
(oracle/make_golden_eval.py -> tests/golden/vilbert_eval512.npz)."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
G = np.load(os.path.join(os.path.dirname(__file__), "golden", "vilbert_eval512.npz"))


@pytest.fixture(scope="module")
def model():
    from multimodal_classification_b200.vilbert import ViLBERTForClassification
    from oracle import vilbert_oracle as vo
    cfg = vo.facebook_config()
    m = ViLBERTForClassification(cfg, num_labels=2)
    m.load_state_dict(vo.seeded_state_dict(cfg), strict=True)
    return m.cuda().eval()


def _auroc(scores, labels):
    order = np.argsort(scores, kind="stable")
    ranks = np.empty(len(scores)); ranks[order] = np.arange(1, len(scores) + 1)
    pos = labels == 1
    n1, n0 = pos.sum(), (~pos).sum()
    return (ranks[pos].sum() - n1 * (n1 + 1) / 2) / (n1 * n0)


def test_eval_set_auroc_ordering_matches_reference(model):
    from oracle import vilbert_oracle as vo
    cfg = model.config
    ours = []
    with torch.no_grad():
        for s in G["seeds"]:
            b = {k: v.cuda() for k, v in vo.synthetic_batch(cfg, batch=64, seq=128, regions=100, seed=int(s)).items()}
            ours.append(model(**b)["logits"].float().cpu().numpy())
    ours = np.concatenate(ours)
    ref = G["logits"]
    scale = np.abs(ref).max()
    err = np.abs(ours - ref).max() / scale
    # bar: the reference's OWN bf16-autocast error on these 512 samples (1.5e-2, stored in the fixture) x 1.5 -- the rule SURVEY.md
    # §8c uses for gradients.  (On the 16-sample fixture the same quantity is 6e-3 against the north-star's 2e-2; the maximum over
    # 1 024 tiny random-init logits is a harsher statistic, and the bf16 residual stream costs ~1.4x autocast's fp32 one.)
    assert err <= 1.5 * float(G["yard_logits"]), (err, float(G["yard_logits"]))
    p_ref = torch.softmax(torch.from_numpy(ref), -1)[:, 1].numpy()
    p_our = torch.softmax(torch.from_numpy(ours), -1)[:, 1].numpy()
    # ordering: every pair of samples the reference separates by more than the tolerated score error keeps its order
    tol = 2.0 * np.abs(p_our - p_ref).max()
    d_ref = p_ref[:, None] - p_ref[None, :]
    d_our = p_our[:, None] - p_our[None, :]
    decided = np.abs(d_ref) > tol
    assert decided.mean() > 0.5, decided.mean()                # the check is not vacuous
    assert np.all(np.sign(d_ref[decided]) == np.sign(d_our[decided]))
    labels = G["labels"]
    a_ref, a_our = _auroc(p_ref, labels), _auroc(p_our, labels)
    print(f"max|dlogit|/max|logit| {err:.2e}  AUROC ref {a_ref:.4f} ours {a_our:.4f}  decided pairs {decided.mean():.3f}")
    assert abs(a_ref - a_our) <= 5e-3


def test_batch_permutation_equivariance(model):
    """Samples are independent through the whole model: permuting the batch permutes the logits bit for bit."""
    from oracle import vilbert_oracle as vo
    b = {k: v.cuda() for k, v in vo.synthetic_batch(model.config, batch=16, seq=128, regions=100, seed=1234).items()}
    perm = torch.randperm(16, generator=torch.Generator().manual_seed(1)).cuda()
    with torch.no_grad():
        a = model(**b)["logits"]
        p = model(**{k: v[perm] for k, v in b.items()})["logits"]
    assert torch.equal(a[perm], p)


def test_masked_tokens_do_not_matter(model):
    """Whatever sits at padded positions (attention_mask = 0) cannot change the logits: exp(-10000 + x - max) is exactly 0 in
    fp32, and the poolers read position 0 only (reference :404-408, :530-540)."""
    from oracle import vilbert_oracle as vo
    b = {k: v.cuda() for k, v in vo.synthetic_batch(model.config, batch=16, seq=128, regions=100, seed=4321).items()}
    with torch.no_grad():
        a = model(**b)["logits"].clone()
        ids = b["input_ids"].clone()
        noise = torch.randint(1, model.config["vocab_size"], ids.shape, generator=torch.Generator().manual_seed(2)).cuda()
        ids = torch.where(b["attention_mask"] == 0, noise, ids)
        c = model(**{**b, "input_ids": ids})["logits"]
    assert torch.equal(a, c)


def test_loss_is_mean_of_per_sample_losses(model):
    from oracle import vilbert_oracle as vo
    b = {k: v.cuda() for k, v in vo.synthetic_batch(model.config, batch=16, seq=128, regions=100, seed=99).items()}
    with torch.no_grad():
        out = model(**b)
    ce = torch.nn.functional.cross_entropy(out["logits"].float(), b["labels"])
    assert abs(out["loss"].item() - ce.item()) <= 1e-5
