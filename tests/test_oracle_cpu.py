"""Pin the oracle (oracle/vilbert_oracle.py) against fixtures produced by the unmodified reference
(oracle/make_golden.py, run in the authoring container).  CPU only."""
import os

import numpy as np
import pytest
import torch

from oracle import vilbert_oracle as vo

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")

CASES = {
    "vilbert_tiny": (vo.tiny_config, dict(batch=4, seq=128, regions=100, seed=1234)),
    "vilbert_tiny_ragged": (vo.tiny_config, dict(batch=3, seq=40, regions=36, seed=77, with_visual_mask=True,
                                                with_token_types=False)),
    "vilbert_full": (vo.facebook_config, dict(batch=16, seq=128, regions=100, seed=1234)),
}


def _load(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False)


@pytest.mark.parametrize("name", ["vilbert_tiny", "vilbert_tiny_ragged"])
def test_oracle_matches_reference_forward_and_grads(name):
    cfg_fn, kw = CASES[name]
    cfg = cfg_fn()
    g = _load(name)
    sd = vo.seeded_state_dict(cfg)
    batch = vo.synthetic_batch(cfg, **kw)
    for scalar in ("loss", "logit1"):
        out, grads = vo.loss_and_grads(sd, cfg, batch, scalar=scalar)
        np.testing.assert_allclose(out["logits"].numpy(), g["logits"], atol=1e-5, rtol=0)
        assert abs(float(out["loss"]) - float(g["loss"])) < 1e-6
        names = [str(n) for n in g["param_names"]]
        assert names == list(sd.keys())
        for n, ref_norm in zip(names, g[f"gradnorm_{scalar}"]):
            if ref_norm < 0:  # the reference never produced a gradient (q_dense1 / q_dense2)
                assert grads[n] is None, n
                assert "q_dense" in n
                continue
            got = float(grads[n].double().norm())
            assert abs(got - ref_norm) <= 1e-4 * ref_norm + 1e-8, (n, got, ref_norm)
        for key in g.files:
            if key.startswith(f"grad_{scalar}/"):
                n = key.split("/", 1)[1]
                ref = g[key]
                np.testing.assert_allclose(grads[n].numpy(), ref, atol=1e-5 * max(1e-6, np.abs(ref).max()) + 1e-9, rtol=1e-3)


def test_oracle_matches_reference_full_config_forward():
    cfg_fn, kw = CASES["vilbert_full"]
    cfg = cfg_fn()
    g = _load("vilbert_full")
    sd = vo.seeded_state_dict(cfg)
    batch = vo.synthetic_batch(cfg, **kw)
    with torch.no_grad():
        out = vo.forward(sd, cfg, **batch, return_hidden=True)
    np.testing.assert_allclose(out["logits"].numpy(), g["logits"], atol=1e-5, rtol=0)
    np.testing.assert_allclose(out["t_pooled"].numpy(), g["t_pooled"], atol=1e-5, rtol=0)
    np.testing.assert_allclose(out["v_pooled"].numpy(), g["v_pooled"], atol=1e-5, rtol=0)
    assert abs(float(out["loss"]) - float(g["loss"])) < 1e-6
    assert len(sd) == 523
    assert sum(v.numel() for v in sd.values()) == 248_826_882


def test_mask_construction_is_bit_exact():
    """(1.0 - m) * -10000.0 for int64 and float masks (vilbert_facebook_arch.py:530-540)."""
    m = torch.tensor([[1, 1, 0, 0]])
    e = vo.extended_mask(m)
    assert e.dtype == torch.float32 and e.shape == (1, 1, 1, 4)
    assert e.flatten().tolist() == [0.0, 0.0, -10000.0, -10000.0]
    assert torch.equal(vo.extended_mask(m.float()), e)
    assert vo.extended_mask(None) is None
