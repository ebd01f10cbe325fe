"""ORACLE (test infrastructure, never shipped, never on the product path).

CPU, fp32, plain-PyTorch restatement of the reference's ViLBERT hot path, written as pure functions over a
``state_dict`` so that it travels to the GPU box (the reference itself, /root/reference, does not).  Only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference`` legs may import it.

Pinning: the reference ships NO golden vectors / known-answer tests for this path, so the
oracle is pinned against outputs of the reference itself: ``oracle/make_golden.py`` imports
``/root/reference/src/multimodalclassification/models/vilbert_facebook_arch.py`` in the authoring container,
runs it on seeded inputs/weights and commits logits / loss / per-tensor gradient digests under ``tests/golden/``;
``tests/test_oracle_cpu.py`` checks this file against those fixtures (bit-level agreement is not expected of two
fp32 op orderings; the measured agreement is ~1e-6 and the test tolerance is 1e-5).

Every function cites the reference lines it follows (paths relative to
/root/reference/src/multimodalclassification/models/).
"""
from __future__ import annotations

import math
from typing import Dict, Optional

import torch
import torch.nn.functional as F

Tensor = torch.Tensor


def facebook_config() -> Dict:
    """vilbert_facebook_arch.py:35-60 (the values are the contract, restated)."""
    return {
        "hidden_size": 768, "num_attention_heads": 12, "num_hidden_layers": 12, "intermediate_size": 3072,
        "hidden_dropout_prob": 0.1, "attention_probs_dropout_prob": 0.1, "max_position_embeddings": 512,
        "vocab_size": 30522,
        "v_hidden_size": 1024, "v_num_attention_heads": 8, "v_num_hidden_layers": 6, "v_intermediate_size": 1024,
        "v_hidden_dropout_prob": 0.1, "v_attention_probs_dropout_prob": 0.1,
        "num_co_attention_layers": 6, "bi_hidden_size": 1024,
        "v_feature_size": 2048, "v_loc_size": 5,
    }


def tiny_config() -> Dict:
    """A reduced configuration with the same structure (heads of 64 / 128 wide, every layer kind present) used for
    golden fixtures that are small enough to commit."""
    c = facebook_config()
    c.update({"num_hidden_layers": 4, "v_num_hidden_layers": 2, "num_co_attention_layers": 2,
              "vocab_size": 1000, "max_position_embeddings": 512})
    return c


# ------------------------------------------------------------------------------------------------ building blocks
def layer_norm(x: Tensor, w: Tensor, b: Tensor, eps: float = 1e-12) -> Tensor:
    """BertLayerNorm.forward, vilbert_facebook_arch.py:72-76 (biased variance, eps inside the sqrt)."""
    u = x.mean(-1, keepdim=True)
    s = (x - u).pow(2).mean(-1, keepdim=True)
    return w * ((x - u) / torch.sqrt(s + eps)) + b


def linear(sd, prefix: str, x: Tensor) -> Tensor:
    return F.linear(x, sd[prefix + ".weight"], sd[prefix + ".bias"])


def split_heads(x: Tensor, heads: int) -> Tensor:
    """transpose_for_scores, vilbert_facebook_arch.py:121-124 / 245-248."""
    b, s, h = x.shape
    return x.view(b, s, heads, h // heads).permute(0, 2, 1, 3)


def merge_heads(x: Tensor) -> Tensor:
    b, h, s, d = x.shape
    return x.permute(0, 2, 1, 3).contiguous().view(b, s, h * d)


def softmax_attention(q: Tensor, k: Tensor, v: Tensor, heads: int, add_mask: Optional[Tensor]) -> Tensor:
    """vilbert_facebook_arch.py:131-144 (self) and :272-285 (cross): QK^T / sqrt(d) + mask -> softmax -> PV."""
    qh, kh, vh = split_heads(q, heads), split_heads(k, heads), split_heads(v, heads)
    scores = torch.matmul(qh, kh.transpose(-1, -2)) / math.sqrt(qh.shape[-1])
    if add_mask is not None:
        scores = scores + add_mask
    probs = F.softmax(scores, dim=-1)
    return merge_heads(torch.matmul(probs, vh))


def extended_mask(mask: Optional[Tensor]) -> Optional[Tensor]:
    """vilbert_facebook_arch.py:530-540: (1.0 - m[:,None,None,:]) * -10000.0 (int64 masks promote to fp32)."""
    if mask is None:
        return None
    return (1.0 - mask.unsqueeze(1).unsqueeze(2)) * -10000.0


def bert_layer(sd, p: str, x: Tensor, mask: Optional[Tensor], heads: int) -> Tensor:
    """BertLayer.forward :215-219 = BertAttention :171-174 -> BertIntermediate :184-185 -> BertOutput :197-201."""
    q = linear(sd, p + ".attention.self.query", x)
    k = linear(sd, p + ".attention.self.key", x)
    v = linear(sd, p + ".attention.self.value", x)
    ctx = softmax_attention(q, k, v, heads, mask)
    a = layer_norm(linear(sd, p + ".attention.output.dense", ctx) + x,
                   sd[p + ".attention.output.LayerNorm.weight"], sd[p + ".attention.output.LayerNorm.bias"])
    i = F.gelu(linear(sd, p + ".intermediate.dense", a))
    return layer_norm(linear(sd, p + ".output.dense", i) + a,
                      sd[p + ".output.LayerNorm.weight"], sd[p + ".output.LayerNorm.bias"])


def co_attention_layer(sd, p: str, cfg, v: Tensor, t: Tensor, v_mask, t_mask):
    """CoAttentionLayer.forward :377-394 = BiAttention :253-294 -> BiOutput :324-338 -> the two FFNs."""
    heads = cfg["v_num_attention_heads"]
    b = p + ".biattention"
    v_ctx = softmax_attention(linear(sd, b + ".query1", v), linear(sd, b + ".key2", t), linear(sd, b + ".value2", t),
                              heads, t_mask)
    t_ctx = softmax_attention(linear(sd, b + ".query2", t), linear(sd, b + ".key1", v), linear(sd, b + ".value1", v),
                              heads, v_mask)
    o = p + ".biOutput"
    v_att = layer_norm(linear(sd, o + ".dense1", v_ctx) + v, sd[o + ".LayerNorm1.weight"], sd[o + ".LayerNorm1.bias"])
    t_att = layer_norm(linear(sd, o + ".dense2", t_ctx) + t, sd[o + ".LayerNorm2.weight"], sd[o + ".LayerNorm2.bias"])
    v_int = F.gelu(linear(sd, p + ".v_intermediate.dense", v_att))
    v_out = layer_norm(linear(sd, p + ".v_output.dense", v_int) + v_att,
                       sd[p + ".v_output.LayerNorm.weight"], sd[p + ".v_output.LayerNorm.bias"])
    t_int = F.gelu(linear(sd, p + ".t_intermediate.dense", t_att))
    t_out = layer_norm(linear(sd, p + ".t_output.dense", t_int) + t_att,
                       sd[p + ".t_output.LayerNorm.weight"], sd[p + ".t_output.LayerNorm.bias"])
    return v_out, t_out


def text_embeddings(sd, input_ids: Tensor, token_type_ids: Optional[Tensor]) -> Tensor:
    """transformers BertEmbeddings.forward (called at vilbert_facebook_arch.py:524): word + type + position ->
    nn.LayerNorm(eps=1e-12).  Dropout is the identity in eval mode."""
    p = "bert.embeddings"
    bsz, seq = input_ids.shape
    if token_type_ids is None:
        token_type_ids = torch.zeros_like(input_ids)
    e = sd[p + ".word_embeddings.weight"][input_ids] + sd[p + ".token_type_embeddings.weight"][token_type_ids]
    e = e + sd[p + ".position_embeddings.weight"][:seq].unsqueeze(0)
    return F.layer_norm(e, (e.shape[-1],), sd[p + ".LayerNorm.weight"], sd[p + ".LayerNorm.bias"], 1e-12)


def visual_embeddings(sd, feats: Tensor, locs: Tensor) -> Tensor:
    """VisualEmbeddings.forward :100-104."""
    p = "bert.v_embeddings"
    s = linear(sd, p + ".image_embeddings", feats) + linear(sd, p + ".image_location_embeddings", locs)
    return layer_norm(s, sd[p + ".LayerNorm.weight"], sd[p + ".LayerNorm.bias"])


# ------------------------------------------------------------------------------------------------ the model
def forward(sd: Dict[str, Tensor], cfg: Dict, input_ids: Tensor, attention_mask: Optional[Tensor] = None,
            token_type_ids: Optional[Tensor] = None, visual_features: Optional[Tensor] = None,
            visual_attention_mask: Optional[Tensor] = None, spatial_locations: Optional[Tensor] = None,
            labels: Optional[Tensor] = None, return_hidden: bool = False) -> Dict[str, Tensor]:
    """ViLBERTForClassification.forward :610-641 in eval mode (dropout = identity)."""
    t = text_embeddings(sd, input_ids, token_type_ids)
    v = visual_embeddings(sd, visual_features, spatial_locations)
    t_mask, v_mask = extended_mask(attention_mask), extended_mask(visual_attention_mask)
    # ViLBERTEncoder.forward :459-481: co-attention after text layers 1,3,5,7,9,11 while co-layers remain
    c = 0
    for i in range(cfg["num_hidden_layers"]):
        t = bert_layer(sd, f"bert.encoder.layer.{i}", t, t_mask, cfg["num_attention_heads"])
        if i in (1, 3, 5, 7, 9, 11) and c < cfg["num_co_attention_layers"]:
            v = bert_layer(sd, f"bert.encoder.v_layer.{c}", v, v_mask, cfg["v_num_attention_heads"])
            v, t = co_attention_layer(sd, f"bert.encoder.c_layer.{c}", cfg, v, t, v_mask, t_mask)
            c += 1
    # BertPooler :404-408 (token / region 0), concat, classifier :569-578, CE :637-639
    t_pooled = torch.tanh(linear(sd, "bert.t_pooler.dense", t[:, 0]))
    v_pooled = torch.tanh(linear(sd, "bert.v_pooler.dense", v[:, 0]))
    pooled = torch.cat([t_pooled, v_pooled], dim=-1)
    h = F.relu(linear(sd, "classifier.1", pooled))
    logits = linear(sd, "classifier.4", h)
    out = {"logits": logits}
    if labels is not None:
        out["loss"] = F.cross_entropy(logits, labels)
    if return_hidden:
        out.update({"t_hidden": t, "v_hidden": v, "t_pooled": t_pooled, "v_pooled": v_pooled})
    return out


# ------------------------------------------------------------------------------------------------ seeded data
def param_shapes(cfg: Dict, num_labels: int = 2) -> Dict[str, tuple]:
    """Appendix A of SURVEY.md: every state_dict key with its shape, in the reference's registration order."""
    H, Hv, bi = cfg["hidden_size"], cfg["v_hidden_size"], cfg["bi_hidden_size"]
    I, Iv = cfg["intermediate_size"], cfg["v_intermediate_size"]
    s: Dict[str, tuple] = {}

    def lin(p, o, i):
        s[p + ".weight"] = (o, i)
        s[p + ".bias"] = (o,)

    def ln(p, n):
        s[p + ".weight"] = (n,)
        s[p + ".bias"] = (n,)

    def bert_layer_shapes(p, h, inter):
        for n in ("query", "key", "value"):
            lin(f"{p}.attention.self.{n}", h, h)
        lin(f"{p}.attention.output.dense", h, h)
        ln(f"{p}.attention.output.LayerNorm", h)
        lin(f"{p}.intermediate.dense", inter, h)
        lin(f"{p}.output.dense", h, inter)
        ln(f"{p}.output.LayerNorm", h)

    s["bert.embeddings.word_embeddings.weight"] = (cfg["vocab_size"], H)
    s["bert.embeddings.position_embeddings.weight"] = (cfg["max_position_embeddings"], H)
    s["bert.embeddings.token_type_embeddings.weight"] = (2, H)
    ln("bert.embeddings.LayerNorm", H)
    lin("bert.v_embeddings.image_embeddings", Hv, cfg["v_feature_size"])
    lin("bert.v_embeddings.image_location_embeddings", Hv, cfg["v_loc_size"])
    ln("bert.v_embeddings.LayerNorm", Hv)
    for i in range(cfg["num_hidden_layers"]):
        bert_layer_shapes(f"bert.encoder.layer.{i}", H, I)
    for i in range(cfg["v_num_hidden_layers"]):
        bert_layer_shapes(f"bert.encoder.v_layer.{i}", Hv, Iv)
    for i in range(cfg["num_co_attention_layers"]):
        p = f"bert.encoder.c_layer.{i}"
        for n in ("query1", "key1", "value1"):
            lin(f"{p}.biattention.{n}", bi, Hv)
        for n in ("query2", "key2", "value2"):
            lin(f"{p}.biattention.{n}", bi, H)
        lin(f"{p}.biOutput.dense1", Hv, bi)
        ln(f"{p}.biOutput.LayerNorm1", Hv)
        lin(f"{p}.biOutput.dense2", H, bi)
        ln(f"{p}.biOutput.LayerNorm2", H)
        lin(f"{p}.biOutput.q_dense1", Hv, bi)
        lin(f"{p}.biOutput.q_dense2", H, bi)
        lin(f"{p}.v_intermediate.dense", Iv, Hv)
        lin(f"{p}.v_output.dense", Hv, Iv)
        ln(f"{p}.v_output.LayerNorm", Hv)
        lin(f"{p}.t_intermediate.dense", I, H)
        lin(f"{p}.t_output.dense", H, I)
        ln(f"{p}.t_output.LayerNorm", H)
    lin("bert.t_pooler.dense", bi, H)
    lin("bert.v_pooler.dense", Hv, Hv)
    lin("classifier.1", bi, bi + Hv)
    lin("classifier.4", num_labels, bi)
    return s


def seeded_state_dict(cfg: Dict, seed: int = 0, num_labels: int = 2, gain: float = 1.0) -> Dict[str, Tensor]:
    """Deterministic random-init weights that do not depend on module construction order or on the reference being
    importable: every tensor is drawn from its own CPU generator seeded by (seed, index).  Matrices ~ U(-a, a) with
    a = gain/sqrt(fan_in) (the scale of nn.Linear's default init), embeddings ~ N(0, 0.02), LayerNorm weights
    1 + 0.1 N(0,1), LayerNorm / Linear biases 0.02 N(0,1)  (non-trivial so that every parameter matters)."""
    sd: Dict[str, Tensor] = {}
    for idx, (k, shp) in enumerate(param_shapes(cfg, num_labels).items()):
        g = torch.Generator(device="cpu").manual_seed(seed * 100003 + idx)
        if "embeddings.weight" in k and len(shp) == 2 and "image" not in k:
            t = torch.randn(shp, generator=g) * 0.02
            if "word_embeddings" in k:
                t[0].zero_()  # padding_idx=0 row (HF initialises it to zero)
        elif "LayerNorm" in k and k.endswith(".weight"):
            t = 1.0 + 0.1 * torch.randn(shp, generator=g)
        elif len(shp) == 2:
            a = gain / math.sqrt(shp[1])
            t = (torch.rand(shp, generator=g) * 2 - 1) * a
        else:
            t = 0.02 * torch.randn(shp, generator=g)
        sd[k] = t
    return sd


def synthetic_batch(cfg: Dict, batch: int = 16, seq: int = 128, regions: int = 100, seed: int = 1234,
                    with_visual_mask: bool = False, with_token_types: bool = True) -> Dict[str, Tensor]:
    """SURVEY.md §8d synthetic LMDB-shaped batch (keys per data_processing/lmdb_dataset.py:230-239)."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    lengths = torch.randint(8, seq, (batch,), generator=g)
    ids = torch.randint(1, cfg["vocab_size"], (batch, seq), generator=g)
    mask = (torch.arange(seq).unsqueeze(0) < lengths.unsqueeze(1)).long()
    ids = ids * mask
    feats = torch.randn(batch, regions, cfg["v_feature_size"], generator=g).abs()
    xy = torch.rand(batch, regions, 2, generator=g) * 0.7
    wh = torch.rand(batch, regions, 2, generator=g) * 0.25 + 0.05
    loc = torch.cat([xy, xy + wh, (wh[..., 0] * wh[..., 1]).unsqueeze(-1)], dim=-1)
    labels = torch.randint(0, 2, (batch,), generator=g)
    out = {"input_ids": ids, "attention_mask": mask, "visual_features": feats, "spatial_locations": loc,
           "labels": labels}
    if with_token_types:
        out["token_type_ids"] = torch.zeros_like(ids)
    if with_visual_mask:
        vm = torch.ones(batch, regions)
        if regions > 4:
            vm[:, regions - 3:] = 0.0  # exercise the visual key mask
        out["visual_attention_mask"] = vm
    return out


def loss_and_grads(sd: Dict[str, Tensor], cfg: Dict, batch: Dict[str, Tensor], scalar: str = "loss"):
    """fp32 autograd of the restated forward: returns (outputs, {name: grad or None})."""
    leaves = {k: v.detach().clone().requires_grad_(True) for k, v in sd.items()}
    out = forward(leaves, cfg, **batch)
    obj = out["loss"] if scalar == "loss" else out["logits"][:, 1].sum()
    obj.backward()
    return {k: v.detach() for k, v in out.items()}, {k: v.grad for k, v in leaves.items()}
