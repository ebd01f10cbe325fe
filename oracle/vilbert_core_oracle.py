"""ORACLE (test infrastructure, never shipped, never on the product path).

CPU, fp32 restatement of the reference's SECOND two-stream surface:
``models/vilbert_core.py::ViLBERTForClassification`` (/root/reference/src/multimodalclassification/models/vilbert_core.py:
271-657).  Same kernels' worth of arithmetic as the Facebook-architecture model, different wiring: both streams are 768 wide
(the text ``BertModel``'s width), the visual embedding adds a learned region-position table, every co-attention block is two
``BertCrossAttention`` modules (query from the own stream, key / value from the other, ``BertSelfOutput`` on the query's
residual) followed by one FFN per stream, the visual stream is mean-pooled, the classifier sees ``[text_pooled,
visual_pooled]``.  Pure functions over a ``state_dict`` with the reference module's key names, built from the blocks of
``oracle/vilbert_oracle.py``.  Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU legs may import it.

State-dict note: the reference keeps a whole ``transformers.BertModel`` under ``vilbert.bert`` but only calls its
``embeddings`` (vilbert_core.py:548-551); its 12 encoder layers and pooler are dead weights that never receive a gradient.

Pinning: no reference test covers this path; ``oracle/make_golden_core.py`` builds the reference class in the authoring
container (``BertModel.from_pretrained`` replaced by a config-built ``BertModel``: no checkpoint offline), loads
``seeded_core_state`` and commits logits / loss / pooled outputs / gradient norms as ``tests/golden/vilbert_core_tiny.npz``;
``tests/test_vilbert_core_oracle_cpu.py`` checks this file against them.
"""
from __future__ import annotations

from typing import Dict, Optional

import torch
import torch.nn.functional as F
from torch import Tensor

from .vilbert_oracle import extended_mask, layer_norm, linear, softmax_attention


def core_config() -> Dict:
    """``get_vilbert_config()`` (vilbert_core.py:668-688) after ``ViLBERTModel.__init__`` copied BERT-base's sizes into it."""
    return {"hidden_size": 768, "num_attention_heads": 12, "intermediate_size": 3072, "hidden_dropout_prob": 0.1,
            "attention_probs_dropout_prob": 0.1, "v_feature_size": 2048, "v_num_hidden_layers": 6, "max_regions": 100,
            "t_num_hidden_layers": 12, "num_co_layers": 6, "classifier_dropout": 0.5, "num_labels": 2,
            "vocab_size": 30522, "max_position_embeddings": 512, "type_vocab_size": 2}


def tiny_core_config() -> Dict:
    """Same structure, sizes small enough for a committed fixture (heads stay 64 wide; hidden 256 = the narrowest row the
    row-wise kernels take)."""
    c = core_config()
    c.update({"hidden_size": 256, "num_attention_heads": 4, "intermediate_size": 512, "v_feature_size": 64,
              "v_num_hidden_layers": 2, "t_num_hidden_layers": 4, "num_co_layers": 2, "max_regions": 40, "vocab_size": 500,
              "max_position_embeddings": 64})
    return c


# ------------------------------------------------------------------------------------------------ blocks
def _self_output(sd, p: str, ctx: Tensor, residual: Tensor) -> Tensor:
    """BertSelfOutput.forward (vilbert_core.py:158-164): LayerNorm(dense(ctx) + residual); dropout = identity in eval."""
    return layer_norm(linear(sd, p + ".dense", ctx) + residual, sd[p + ".LayerNorm.weight"], sd[p + ".LayerNorm.bias"])


def _ffn(sd, p_int: str, p_out: str, x: Tensor) -> Tensor:
    """BertIntermediate (:177-180, erf GELU) -> BertOutput (:194-200)."""
    i = F.gelu(linear(sd, p_int + ".dense", x))
    return layer_norm(linear(sd, p_out + ".dense", i) + x, sd[p_out + ".LayerNorm.weight"], sd[p_out + ".LayerNorm.bias"])


def bert_layer(sd, p: str, x: Tensor, mask: Optional[Tensor], heads: int) -> Tensor:
    """BertLayer.forward (:255-268)."""
    s = p + ".attention.self"
    ctx = softmax_attention(linear(sd, s + ".query", x), linear(sd, s + ".key", x), linear(sd, s + ".value", x), heads, mask)
    return _ffn(sd, p + ".intermediate", p + ".output", _self_output(sd, p + ".attention.output", ctx, x))


def cross_attention(sd, p: str, query_side: Tensor, key_side: Tensor, key_mask: Optional[Tensor], heads: int) -> Tensor:
    """BertCrossAttention.forward (:231-243) over BertCoAttention (:114-145): the mask belongs to the key side."""
    s = p + ".self"
    ctx = softmax_attention(linear(sd, s + ".query", query_side), linear(sd, s + ".key", key_side),
                            linear(sd, s + ".value", key_side), heads, key_mask)
    return _self_output(sd, p + ".output", ctx, query_side)


def connection_layer(sd, p: str, v: Tensor, t: Tensor, v_mask, t_mask, heads: int):
    """BertConnectionLayer.forward (:292-330): both directions read the INPUT hidden states, then one FFN per stream."""
    v_att = cross_attention(sd, p + ".biattention_v", v, t, t_mask, heads)
    t_att = cross_attention(sd, p + ".biattention_t", t, v, v_mask, heads)
    return _ffn(sd, p + ".intermediate_v", p + ".output_v", v_att), _ffn(sd, p + ".intermediate_t", p + ".output_t", t_att)


def text_embeddings(sd, input_ids: Tensor, token_type_ids: Optional[Tensor]) -> Tensor:
    """transformers BertEmbeddings.forward as called at vilbert_core.py:548-551."""
    p = "vilbert.bert.embeddings"
    if token_type_ids is None:
        token_type_ids = torch.zeros_like(input_ids)
    e = sd[p + ".word_embeddings.weight"][input_ids] + sd[p + ".token_type_embeddings.weight"][token_type_ids]
    e = e + sd[p + ".position_embeddings.weight"][: input_ids.shape[1]].unsqueeze(0)
    return F.layer_norm(e, (e.shape[-1],), sd[p + ".LayerNorm.weight"], sd[p + ".LayerNorm.bias"], 1e-12)


def visual_embeddings(sd, feats: Tensor, locs: Optional[Tensor]) -> Tensor:
    """ViLBERTEmbeddings.forward (:447-480): image projection [+ location projection] + region-position table -> LayerNorm."""
    p = "vilbert.visual_embeddings"
    e = linear(sd, p + ".image_embeddings", feats)
    if locs is not None:
        e = e + linear(sd, p + ".location_embeddings", locs)
    e = e + sd[p + ".position_embeddings.weight"][: feats.shape[1]].unsqueeze(0)
    return layer_norm(e, sd[p + ".LayerNorm.weight"], sd[p + ".LayerNorm.bias"])


def forward(sd: Dict[str, Tensor], cfg: Dict, input_ids: Tensor, attention_mask: Optional[Tensor] = None,
            token_type_ids: Optional[Tensor] = None, visual_features: Optional[Tensor] = None,
            visual_attention_mask: Optional[Tensor] = None, spatial_locations: Optional[Tensor] = None,
            labels: Optional[Tensor] = None) -> Dict[str, Tensor]:
    """ViLBERTForClassification.forward (:620-657) over ViLBERTModel.forward (:524-590) and ViLBERTEncoder.forward (:372-416):
    a text layer every step; after every second one a visual layer (while any are left) and a connection layer."""
    heads = cfg["num_attention_heads"]
    t = text_embeddings(sd, input_ids, token_type_ids)
    v = visual_embeddings(sd, visual_features, spatial_locations)
    t_mask, v_mask = extended_mask(attention_mask), extended_mask(visual_attention_mask)
    v_idx = co_idx = 0
    for t_idx in range(cfg["t_num_hidden_layers"]):
        t = bert_layer(sd, f"vilbert.encoder.t_layer.{t_idx}", t, t_mask, heads)
        if (t_idx + 1) % 2 == 0 and co_idx < cfg["num_co_layers"]:
            if v_idx < cfg["v_num_hidden_layers"]:
                v = bert_layer(sd, f"vilbert.encoder.v_layer.{v_idx}", v, v_mask, heads)
                v_idx += 1
            v, t = connection_layer(sd, f"vilbert.encoder.c_layer.{co_idx}", v, t, v_mask, t_mask, heads)
            co_idx += 1
    text_pooled = torch.tanh(linear(sd, "vilbert.t_pooler.0", t[:, 0]))
    visual_pooled = torch.tanh(linear(sd, "vilbert.v_pooler.0", v.mean(dim=1)))
    pooled = torch.cat([text_pooled, visual_pooled], dim=-1)
    logits = linear(sd, "classifier.4", F.relu(linear(sd, "classifier.1", pooled)))
    out = {"logits": logits, "pooled_output": pooled, "text_pooled": text_pooled, "visual_pooled": visual_pooled,
           "text_output": t, "visual_output": v}
    if labels is not None:
        out["loss"] = F.cross_entropy(logits, labels)
    return out


# ------------------------------------------------------------------------------------------------ seeded weights
def used_prefixes() -> tuple:
    """Parameters outside these prefixes exist in the reference state_dict but never reach the output (the unused encoder
    and pooler of the text BertModel)."""
    return ("vilbert.bert.embeddings.", "vilbert.visual_embeddings.", "vilbert.encoder.", "vilbert.t_pooler.",
            "vilbert.v_pooler.", "classifier.")


def seeded_core_state(shapes: Dict[str, tuple], seed: int = 0) -> Dict[str, Tensor]:
    """Deterministic weights for every floating-point tensor of a reference ``state_dict`` (``shapes``: key -> shape, taken
    from the reference module, so that the key set is the reference's by construction): normal(0, 0.05) matrices, LayerNorm
    gains around one, small biases."""
    g = torch.Generator().manual_seed(seed)
    sd = {}
    for k, shape in shapes.items():
        if k.endswith("LayerNorm.weight"):
            sd[k] = 0.75 + 0.5 * torch.rand(shape, generator=g)
        elif k.endswith(".bias"):
            sd[k] = (torch.rand(shape, generator=g) - 0.5) * 0.2
        else:
            sd[k] = torch.randn(shape, generator=g) * (0.05 if len(shape) > 1 else 0.02)
    return sd


def param_shapes(cfg: Dict, num_labels: int = 2) -> Dict[str, tuple]:
    """Shapes of the tensors the forward pass reads (the used subset of the reference state_dict), keyed as the reference."""
    h, inter = cfg["hidden_size"], cfg["intermediate_size"]
    s: Dict[str, tuple] = {}

    def lin(p, o, i):
        s[p + ".weight"], s[p + ".bias"] = (o, i), (o,)

    def ln(p):
        s[p + ".weight"], s[p + ".bias"] = (h,), (h,)

    e = "vilbert.bert.embeddings"
    s[e + ".word_embeddings.weight"] = (cfg["vocab_size"], h)
    s[e + ".position_embeddings.weight"] = (cfg["max_position_embeddings"], h)
    s[e + ".token_type_embeddings.weight"] = (cfg["type_vocab_size"], h)
    ln(e + ".LayerNorm")
    v = "vilbert.visual_embeddings"
    lin(v + ".image_embeddings", h, cfg["v_feature_size"])
    lin(v + ".location_embeddings", h, 5)
    s[v + ".position_embeddings.weight"] = (cfg["max_regions"], h)
    ln(v + ".LayerNorm")

    def attn(p):
        for n in ("query", "key", "value"):
            lin(p + ".self." + n, h, h)
        lin(p + ".output.dense", h, h)
        ln(p + ".output.LayerNorm")

    def ffn(p_int, p_out):
        lin(p_int + ".dense", inter, h)
        lin(p_out + ".dense", h, inter)
        ln(p_out + ".LayerNorm")

    for stream, n in (("v_layer", cfg["v_num_hidden_layers"]), ("t_layer", cfg["t_num_hidden_layers"])):
        for i in range(n):
            p = f"vilbert.encoder.{stream}.{i}"
            attn(p + ".attention")
            ffn(p + ".intermediate", p + ".output")
    for i in range(cfg["num_co_layers"]):
        p = f"vilbert.encoder.c_layer.{i}"
        attn(p + ".biattention_v")
        attn(p + ".biattention_t")
        ffn(p + ".intermediate_v", p + ".output_v")
        ffn(p + ".intermediate_t", p + ".output_t")
    lin("vilbert.t_pooler.0", h, h)
    lin("vilbert.v_pooler.0", h, h)
    lin("classifier.1", h, 2 * h)
    lin("classifier.4", num_labels, h)
    return s


def synthetic_batch(cfg: Dict, batch: int = 4, seq: int = 32, regions: int = 20, seed: int = 1234) -> Dict[str, Tensor]:
    """LMDB-shaped batch as in vilbert_oracle.synthetic_batch, with a ragged visual mask (this surface is fed by the
    on-the-fly extractors, which pad)."""
    g = torch.Generator().manual_seed(seed)
    lengths = torch.randint(4, seq, (batch,), generator=g)
    mask = (torch.arange(seq).unsqueeze(0) < lengths.unsqueeze(1)).long()
    ids = torch.randint(1, cfg["vocab_size"], (batch, seq), generator=g) * mask
    feats = torch.randn(batch, regions, cfg["v_feature_size"], generator=g).abs()
    xy = torch.rand(batch, regions, 2, generator=g) * 0.7
    wh = torch.rand(batch, regions, 2, generator=g) * 0.25 + 0.05
    loc = torch.cat([xy, xy + wh, (wh[..., 0] * wh[..., 1]).unsqueeze(-1)], dim=-1)
    vlen = torch.randint(regions // 2, regions + 1, (batch,), generator=g)
    vmask = (torch.arange(regions).unsqueeze(0) < vlen.unsqueeze(1)).float()
    return {"input_ids": ids, "attention_mask": mask, "token_type_ids": torch.zeros_like(ids), "visual_features": feats,
            "visual_attention_mask": vmask, "spatial_locations": loc, "labels": torch.randint(0, 2, (batch,), generator=g)}


def loss_and_grads(sd: Dict[str, Tensor], cfg: Dict, batch: Dict[str, Tensor]):
    """Forward + autograd of the CE loss over every used parameter: (outputs, {key: grad})."""
    leaves = {k: v.clone().requires_grad_(True) for k, v in sd.items() if v.is_floating_point()}
    out = forward(leaves, cfg, **batch)
    used = [k for k in leaves if k.startswith(used_prefixes())]
    grads = torch.autograd.grad(out["loss"], [leaves[k] for k in used], allow_unused=True)
    return {k: v.detach() for k, v in out.items()}, dict(zip(used, grads))

