"""The reference's second two-stream surface on the same engine: drop-in for
``models/vilbert_core.py::ViLBERTForClassification`` (/root/reference/src/multimodalclassification/models/vilbert_core.py:
593-657) with the reference's parameter tree and ``state_dict`` keys (``vilbert.bert.*`` — a whole ``transformers.BertModel``
of which only the embeddings are used —, ``vilbert.visual_embeddings.*``, ``vilbert.encoder.{v_layer,t_layer,c_layer}.*``,
``vilbert.{t,v}_pooler.0.*``, ``classifier.{1,4}.*``).

Structurally this model IS the Facebook-architecture model with every width set to the text BertModel's (768, 12 heads of
64) plus three differences, which ``vilbert._Engine`` switches on from the configuration / parameter set it is handed:

* a learned region-position table in the visual embedding (``V_POS_KEY``; read through a one-hot GEMM operand);
* the visual stream is mean-pooled before its pooler (``_v_pool = "mean"``);
* classifier dropout 0.5 (``_classifier_dropout``).

``BertConnectionLayer`` (two cross-attentions, each with its own output dense + LayerNorm on the query's residual, then one
FFN per stream; vilbert_core.py:271-330) is the engine's co-attention block under other names: the projections applied to the
visual stream are ``biattention_v.self.query`` and ``biattention_t.self.{key,value}``, those applied to the text stream
``biattention_t.self.query`` and ``biattention_v.self.{key,value}``.  ``_engine_named_parameters`` hands the parameters to
the engine under the Facebook layout's names, so the flat buffers, fused q|k|v GEMMs, graphs, gradient buckets and the fused
optimizer apply unchanged.

STATUS: verified against the reference-made fixture and the pinned oracle on the B200 (tests/test_vilbert_core_gpu.py) and, for
the host schedule, in the GPU-less container over the functional kernel stand-ins (tests/test_vilbert_core_cpu.py);
``dropin.install()`` binds it.  CUDA only, no fall-back.
"""
from __future__ import annotations

from typing import Any, Dict, Iterator, Tuple

import torch
import torch.nn as nn

from ._lib import VbError
from .vilbert import V_POS_KEY, ViLBERTForClassification, _bert_layer


def get_vilbert_config() -> Dict[str, Any]:
    """Reference ``get_vilbert_config`` (vilbert_core.py:668-688)."""
    return {"hidden_size": 768, "num_attention_heads": 12, "intermediate_size": 3072, "hidden_dropout_prob": 0.1,
            "attention_probs_dropout_prob": 0.1, "v_feature_size": 2048, "v_num_hidden_layers": 6, "max_regions": 100,
            "t_num_hidden_layers": 12, "num_co_layers": 6, "classifier_dropout": 0.5, "num_labels": 2}


def _self_output(h: int) -> nn.Module:
    m = nn.Module()
    m.dense = nn.Linear(h, h)
    m.LayerNorm = nn.LayerNorm(h, eps=1e-12)
    return m


def _cross_attention(h: int) -> nn.Module:
    """Parameter container of ``BertCrossAttention`` (vilbert_core.py:223-243)."""
    m = nn.Module()
    m.self = nn.Module()
    m.self.query, m.self.key, m.self.value = nn.Linear(h, h), nn.Linear(h, h), nn.Linear(h, h)
    m.output = _self_output(h)
    return m


def _ffn_pair(m: nn.Module, tag: str, h: int, inter: int) -> None:
    i, o = nn.Module(), nn.Module()
    i.dense = nn.Linear(h, inter)
    o.dense = nn.Linear(inter, h)
    o.LayerNorm = nn.LayerNorm(h, eps=1e-12)
    setattr(m, "intermediate_" + tag, i)
    setattr(m, "output_" + tag, o)


def _connection_layer(h: int, inter: int) -> nn.Module:
    m = nn.Module()
    m.biattention_v, m.biattention_t = _cross_attention(h), _cross_attention(h)
    _ffn_pair(m, "v", h, inter)
    _ffn_pair(m, "t", h, inter)
    return m


class ViLBERTForClassification(ViLBERTForClassification):          # noqa: F811 - the reference class's own name
    """Reference ``vilbert_core.ViLBERTForClassification(config, num_labels=2, bert_model_name="bert-base-uncased")``.
    Extra keyword-only argument ``bert_config``: a ``transformers.BertConfig`` to build the text ``BertModel`` from instead of
    downloading ``bert_model_name`` (no network on the test machines)."""

    def __init__(self, config: Dict[str, Any], num_labels: int = 2, bert_model_name: str = "bert-base-uncased", *,
                 bert_config=None):
        nn.Module.__init__(self)
        from transformers import BertModel
        bert = BertModel(bert_config) if bert_config is not None else BertModel.from_pretrained(bert_model_name)
        bc = bert.config
        config.update(hidden_size=bc.hidden_size, num_attention_heads=bc.num_attention_heads,          # vilbert_core.py:500-507
                      intermediate_size=bc.intermediate_size, hidden_dropout_prob=bc.hidden_dropout_prob,
                      attention_probs_dropout_prob=bc.attention_probs_dropout_prob)
        h, inter = bc.hidden_size, bc.intermediate_size
        if config.get("v_num_hidden_layers", 6) < config.get("num_co_layers", 6):
            raise VbError("fewer visual layers than co-attention connections is not supported")
        self.num_labels = num_labels
        self.core_config = config
        v = self.vilbert = nn.Module()
        v.bert = bert
        ve = v.visual_embeddings = nn.Module()
        ve.image_embeddings = nn.Linear(config.get("v_feature_size", 2048), h)
        ve.location_embeddings = nn.Linear(5, h)
        ve.position_embeddings = nn.Embedding(config.get("max_regions", 100), h)
        ve.LayerNorm = nn.LayerNorm(h, eps=1e-12)
        enc = v.encoder = nn.Module()
        enc.v_layer = nn.ModuleList([_bert_layer(h, inter) for _ in range(config.get("v_num_hidden_layers", 6))])
        enc.t_layer = nn.ModuleList([_bert_layer(h, inter) for _ in range(config.get("t_num_hidden_layers", 12))])
        enc.c_layer = nn.ModuleList([_connection_layer(h, inter) for _ in range(config.get("num_co_layers", 6))])
        v.t_pooler = nn.Sequential(nn.Linear(h, h), nn.Tanh())
        v.v_pooler = nn.Sequential(nn.Linear(h, h), nn.Tanh())
        pc = config.get("classifier_dropout", 0.5)
        self.classifier = nn.Sequential(nn.Dropout(pc), nn.Linear(2 * h, h), nn.ReLU(), nn.Dropout(pc), nn.Linear(h, num_labels))
        # the engine's view of this model: the Facebook-architecture configuration with every width = the text model's
        self.config = {"hidden_size": h, "num_attention_heads": bc.num_attention_heads, "num_hidden_layers": config.get("t_num_hidden_layers", 12),
                       "intermediate_size": inter, "hidden_dropout_prob": bc.hidden_dropout_prob,
                       "attention_probs_dropout_prob": bc.attention_probs_dropout_prob, "vocab_size": bc.vocab_size,
                       "max_position_embeddings": bc.max_position_embeddings, "type_vocab_size": bc.type_vocab_size,
                       "v_hidden_size": h, "v_num_attention_heads": bc.num_attention_heads,
                       "v_num_hidden_layers": config.get("num_co_layers", 6), "v_intermediate_size": inter,
                       "v_hidden_dropout_prob": bc.hidden_dropout_prob, "v_attention_probs_dropout_prob": bc.attention_probs_dropout_prob,
                       "v_feature_size": config.get("v_feature_size", 2048), "v_loc_size": 5, "bi_hidden_size": h,
                       "bi_num_attention_heads": bc.num_attention_heads, "num_co_attention_layers": config.get("num_co_layers", 6),
                       "_v_pool": "mean", "_classifier_dropout": pc}
        self._engine = None
        self._anchor = None
        self._ddp_group = None
        self._ddp_compress = None

    # -- the engine's parameter names ---------------------------------------------------------------------------------
    def _engine_named_parameters(self) -> Iterator[Tuple[str, nn.Parameter]]:
        own = dict(self.named_parameters())
        n_co = self.config["num_co_attention_layers"]
        out: Dict[str, nn.Parameter] = {}

        def put(fb: str, core: str, both=("weight", "bias")):
            for leaf in both:
                out[f"{fb}.{leaf}"] = own.pop(f"{core}.{leaf}")

        def layer(fb: str, core: str):
            for n in ("query", "key", "value"):
                put(f"{fb}.attention.self.{n}", f"{core}.attention.self.{n}")
            put(fb + ".attention.output.dense", core + ".attention.output.dense")
            put(fb + ".attention.output.LayerNorm", core + ".attention.output.LayerNorm")
            put(fb + ".intermediate.dense", core + ".intermediate.dense")
            put(fb + ".output.dense", core + ".output.dense")
            put(fb + ".output.LayerNorm", core + ".output.LayerNorm")

        e = "vilbert.bert.embeddings"
        for n in ("word_embeddings", "position_embeddings", "token_type_embeddings"):
            put("bert.embeddings." + n, f"{e}.{n}", ("weight",))
        put("bert.embeddings.LayerNorm", e + ".LayerNorm")
        ve = "vilbert.visual_embeddings"
        put("bert.v_embeddings.image_embeddings", ve + ".image_embeddings")
        put("bert.v_embeddings.image_location_embeddings", ve + ".location_embeddings")
        put("bert.v_embeddings.LayerNorm", ve + ".LayerNorm")
        out[V_POS_KEY] = own.pop(ve + ".position_embeddings.weight")
        for i in range(self.config["num_hidden_layers"]):
            layer(f"bert.encoder.layer.{i}", f"vilbert.encoder.t_layer.{i}")
        for i in range(n_co):
            layer(f"bert.encoder.v_layer.{i}", f"vilbert.encoder.v_layer.{i}")
            fb, c = f"bert.encoder.c_layer.{i}", f"vilbert.encoder.c_layer.{i}"
            put(fb + ".biattention.query1", c + ".biattention_v.self.query")       # applied to the visual stream
            put(fb + ".biattention.key1", c + ".biattention_t.self.key")
            put(fb + ".biattention.value1", c + ".biattention_t.self.value")
            put(fb + ".biattention.query2", c + ".biattention_t.self.query")       # applied to the text stream
            put(fb + ".biattention.key2", c + ".biattention_v.self.key")
            put(fb + ".biattention.value2", c + ".biattention_v.self.value")
            put(fb + ".biOutput.dense1", c + ".biattention_v.output.dense")
            put(fb + ".biOutput.LayerNorm1", c + ".biattention_v.output.LayerNorm")
            put(fb + ".biOutput.dense2", c + ".biattention_t.output.dense")
            put(fb + ".biOutput.LayerNorm2", c + ".biattention_t.output.LayerNorm")
            put(fb + ".v_intermediate.dense", c + ".intermediate_v.dense")
            put(fb + ".v_output.dense", c + ".output_v.dense")
            put(fb + ".v_output.LayerNorm", c + ".output_v.LayerNorm")
            put(fb + ".t_intermediate.dense", c + ".intermediate_t.dense")
            put(fb + ".t_output.dense", c + ".output_t.dense")
            put(fb + ".t_output.LayerNorm", c + ".output_t.LayerNorm")
        put("bert.t_pooler.dense", "vilbert.t_pooler.0")
        put("bert.v_pooler.dense", "vilbert.v_pooler.0")
        put("classifier.1", "classifier.1")
        put("classifier.4", "classifier.4")
        for k, p in own.items():            # the text BertModel's encoder / pooler and visual layers beyond the connections
            out["unused." + k] = p
        return iter(out.items())

    def freeze_bert_layers(self, num_layers: int = 6) -> None:
        """As ``ViLBERTFacebook.freeze_layers`` does for this surface (vilbert_facebook.py:234-247): the text embeddings and
        the first N text layers of the encoder."""
        if num_layers <= 0:
            return
        for p in self.vilbert.bert.embeddings.parameters():
            p.requires_grad = False
        for i, layer in enumerate(self.vilbert.encoder.t_layer):
            if i < num_layers:
                for p in layer.parameters():
                    p.requires_grad = False

    def predict_proba(self, logits: torch.Tensor) -> torch.Tensor:
        return torch.softmax(logits, dim=-1)

    def predict(self, logits: torch.Tensor) -> torch.Tensor:
        return torch.argmax(logits, dim=-1)
