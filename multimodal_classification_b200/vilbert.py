"""B200-native ``ViLBERTForClassification``: the drop-in for the reference's
``models/vilbert_facebook_arch.py`` (class at :554, forward at :610-641).

The module tree exists only to hold fp32 master ``nn.Parameter``s under the reference's names (Appendix A of SURVEY.md:
523 tensors, identical ``state_dict`` layout).  ``forward`` never calls a torch compute op: it stages the batch into
static device buffers and replays a hand-scheduled sequence of kernels from ``libvilbert_b200.so`` (engine below) —
forward and, through one ``torch.autograd.Function``, backward — captured in CUDA graphs, with the text and visual
streams of the encoder running concurrently between co-attention joins.

Precision: fp32 master weights and gradients, bf16 weight shadows / activations, fp32 accumulation everywhere.
"""
from __future__ import annotations

import contextlib
import operator
import os
import weakref
from typing import Any, Dict, List, Optional, Tuple

import torch
import torch.nn as nn

from . import _lib, ops
from ._lib import VbError

CO_ATTENTION_TEXT_LAYERS = (1, 3, 5, 7, 9, 11)  # reference :457
V_POS_KEY = "bert.v_embeddings.position_embeddings.weight"   # only the vilbert_core surface has it


def get_facebook_vilbert_config() -> Dict[str, Any]:
    """Same dictionary as the reference's ``get_facebook_vilbert_config`` (vilbert_facebook_arch.py:35-60)."""
    return {
        "hidden_size": 768, "num_attention_heads": 12, "num_hidden_layers": 12, "intermediate_size": 3072,
        "hidden_dropout_prob": 0.1, "attention_probs_dropout_prob": 0.1, "max_position_embeddings": 512,
        "vocab_size": 30522,
        "v_hidden_size": 1024, "v_num_attention_heads": 8, "v_num_hidden_layers": 6, "v_intermediate_size": 1024,
        "v_hidden_dropout_prob": 0.1, "v_attention_probs_dropout_prob": 0.1,
        "num_co_attention_layers": 6, "bi_hidden_size": 1024,
        "v_feature_size": 2048, "v_loc_size": 5,
    }


# ---------------------------------------------------------------------------------------------------------------------
# parameter containers (names = the reference's state_dict keys)
# ---------------------------------------------------------------------------------------------------------------------
class _LayerNormParams(nn.Module):
    def __init__(self, n):
        super().__init__()
        self.weight = nn.Parameter(torch.ones(n))
        self.bias = nn.Parameter(torch.zeros(n))


class _Holder(nn.Module):
    pass


def _self_attention(h):
    m = _Holder()
    m.query, m.key, m.value = nn.Linear(h, h), nn.Linear(h, h), nn.Linear(h, h)
    return m


def _dense_ln(i, o):
    m = _Holder()
    m.dense = nn.Linear(i, o)
    m.LayerNorm = _LayerNormParams(o)
    return m


def _dense(i, o):
    m = _Holder()
    m.dense = nn.Linear(i, o)
    return m


def _bert_layer(h, inter):
    m = _Holder()
    m.attention = _Holder()
    m.attention.self = _self_attention(h)
    m.attention.output = _dense_ln(h, h)
    m.intermediate = _dense(h, inter)
    m.output = _dense_ln(inter, h)
    return m


def _co_layer(cfg):
    H, Hv, bi = cfg["hidden_size"], cfg["v_hidden_size"], cfg["bi_hidden_size"]
    m = _Holder()
    b = m.biattention = _Holder()
    b.query1, b.key1, b.value1 = nn.Linear(Hv, bi), nn.Linear(Hv, bi), nn.Linear(Hv, bi)
    b.query2, b.key2, b.value2 = nn.Linear(H, bi), nn.Linear(H, bi), nn.Linear(H, bi)
    o = m.biOutput = _Holder()
    o.dense1, o.LayerNorm1 = nn.Linear(bi, Hv), _LayerNormParams(Hv)
    o.dense2, o.LayerNorm2 = nn.Linear(bi, H), _LayerNormParams(H)
    o.q_dense1, o.q_dense2 = nn.Linear(bi, Hv), nn.Linear(bi, H)  # present, never used (reference :319-320)
    m.v_intermediate = _dense(Hv, cfg["v_intermediate_size"])
    m.v_output = _dense_ln(cfg["v_intermediate_size"], Hv)
    m.t_intermediate = _dense(H, cfg["intermediate_size"])
    m.t_output = _dense_ln(cfg["intermediate_size"], H)
    return m


class _TextEmbeddings(nn.Module):
    """Parameter layout of transformers ``BertEmbeddings`` (the reference reuses ``BertModel(...).embeddings``)."""

    def __init__(self, cfg):
        super().__init__()
        H = cfg["hidden_size"]
        self.word_embeddings = nn.Embedding(cfg["vocab_size"], H, padding_idx=0)
        self.position_embeddings = nn.Embedding(cfg["max_position_embeddings"], H)
        self.token_type_embeddings = nn.Embedding(2, H)
        self.LayerNorm = nn.LayerNorm(H, eps=1e-12)
        for e in (self.word_embeddings, self.position_embeddings, self.token_type_embeddings):
            nn.init.normal_(e.weight, mean=0.0, std=0.02)
        with torch.no_grad():
            self.word_embeddings.weight[0].zero_()


class _Backbone(nn.Module):
    def __init__(self, cfg):
        super().__init__()
        H, Hv = cfg["hidden_size"], cfg["v_hidden_size"]
        self.embeddings = _TextEmbeddings(cfg)
        ve = self.v_embeddings = _Holder()
        ve.image_embeddings = nn.Linear(cfg["v_feature_size"], Hv)
        ve.image_location_embeddings = nn.Linear(cfg["v_loc_size"], Hv)
        ve.LayerNorm = _LayerNormParams(Hv)
        enc = self.encoder = _Holder()
        enc.layer = nn.ModuleList([_bert_layer(H, cfg["intermediate_size"]) for _ in range(cfg["num_hidden_layers"])])
        enc.v_layer = nn.ModuleList([_bert_layer(Hv, cfg["v_intermediate_size"]) for _ in range(cfg["v_num_hidden_layers"])])
        enc.c_layer = nn.ModuleList([_co_layer(cfg) for _ in range(cfg["num_co_attention_layers"])])
        self.t_pooler = _dense(H, cfg["bi_hidden_size"])
        self.v_pooler = _dense(Hv, Hv)


# ---------------------------------------------------------------------------------------------------------------------
# flat parameter storage
# ---------------------------------------------------------------------------------------------------------------------
def _align(n, a=64):
    return (n + a - 1) // a * a


_DATA_PTR = torch.Tensor.data_ptr
_VERSION = operator.attrgetter("_version")
_LIVE_FLATS: "weakref.WeakSet" = weakref.WeakSet()
_HOOK = []


def _install_optimizer_hook() -> None:
    """One process-wide ``torch.optim`` post-step hook: after ANY optimizer step, every live engine whose parameters moved
    records the "parameters updated" event (``_FlatParams.note_updated``).  Costs one version sum per engine per step."""
    if _HOOK:
        return
    from torch.optim.optimizer import register_optimizer_step_post_hook

    def hook(optimizer, args, kwargs):
        for flat in list(_LIVE_FLATS):
            if flat.device.type != "cuda":
                continue
            ver = flat.versions()
            if ver != flat._version:
                with torch.cuda.device(flat.device):
                    flat.note_updated(ver)
    _HOOK.append(register_optimizer_step_post_hook(hook))


class _FlatParams:
    """All parameters live in ONE fp32 device buffer (and gradients in a second one with the same layout) so that
    fused operands (q|k|v weights and biases) are contiguous without copies, the bf16 shadow refresh and the gradient
    zeroing are single launches, and the data-parallel all-reduce works on contiguous byte ranges.

    Layout: [W: GEMM weight matrices, shadowed in bf16][S: biases / LayerNorm / location weights / embedding tables
    (gradients accumulated with atomics -> zeroed every backward)][U: parameters the forward never reads]."""

    def __init__(self, model: "ViLBERTForClassification", device):
        cfg = model.config
        # a sibling surface (vilbert_core.py) hands its parameters over under THIS layout's names
        named = dict(getattr(model, "_engine_named_parameters", model.named_parameters)())
        order_w: List[str] = []
        order_s: List[str] = []

        def lin(prefix, fused=None):
            order_w.append(prefix + ".weight")
            order_s.append(prefix + ".bias")

        def ln(prefix):
            order_s.extend([prefix + ".weight", prefix + ".bias"])

        def bert_layer(p):
            for n in ("query", "key", "value"):
                order_w.append(f"{p}.attention.self.{n}.weight")
            for n in ("query", "key", "value"):
                order_s.append(f"{p}.attention.self.{n}.bias")
            lin(p + ".attention.output.dense"); ln(p + ".attention.output.LayerNorm")
            lin(p + ".intermediate.dense"); lin(p + ".output.dense"); ln(p + ".output.LayerNorm")

        for i in range(cfg["num_hidden_layers"]):
            bert_layer(f"bert.encoder.layer.{i}")
        for i in range(cfg["v_num_hidden_layers"]):
            bert_layer(f"bert.encoder.v_layer.{i}")
        for i in range(cfg["num_co_attention_layers"]):
            p = f"bert.encoder.c_layer.{i}"
            for side in ("1", "2"):
                for n in ("query", "key", "value"):
                    order_w.append(f"{p}.biattention.{n}{side}.weight")
                for n in ("query", "key", "value"):
                    order_s.append(f"{p}.biattention.{n}{side}.bias")
            lin(p + ".biOutput.dense1"); ln(p + ".biOutput.LayerNorm1")
            lin(p + ".biOutput.dense2"); ln(p + ".biOutput.LayerNorm2")
            lin(p + ".v_intermediate.dense"); lin(p + ".v_output.dense"); ln(p + ".v_output.LayerNorm")
            lin(p + ".t_intermediate.dense"); lin(p + ".t_output.dense"); ln(p + ".t_output.LayerNorm")
        lin("bert.v_embeddings.image_embeddings")
        lin("bert.t_pooler.dense"); lin("bert.v_pooler.dense"); lin("classifier.1")
        if V_POS_KEY in named:      # learned region-position table (vilbert_core.py:436): read through a one-hot GEMM operand
            order_w.append(V_POS_KEY)
        order_s += ["bert.v_embeddings.image_location_embeddings.weight", "bert.v_embeddings.image_location_embeddings.bias",
                    "bert.v_embeddings.LayerNorm.weight", "bert.v_embeddings.LayerNorm.bias",
                    "classifier.4.weight", "classifier.4.bias",
                    "bert.embeddings.LayerNorm.weight", "bert.embeddings.LayerNorm.bias",
                    "bert.embeddings.token_type_embeddings.weight", "bert.embeddings.position_embeddings.weight",
                    "bert.embeddings.word_embeddings.weight"]
        used = set(order_w) | set(order_s)
        self.order_w = order_w
        order_u = [k for k in named if k not in used]
        assert all("q_dense" in k or k.startswith("unused.") for k in order_u), order_u

        self.offsets: Dict[str, int] = {}
        off = 0
        for k in order_w:
            n = named[k].numel()
            assert n % 64 == 0, k
            self.offsets[k] = off
            off += n
        self.w_end = off
        for k in order_s:
            self.offsets[k] = off
            off += _align(named[k].numel())
        self.s_end = off
        for k in order_u:
            self.offsets[k] = off
            off += _align(named[k].numel())
        self.total = off
        self.used = used
        self.device = device
        self.master = torch.empty(self.total, dtype=torch.float32, device=device)
        self.master.zero_()
        self.grad = torch.zeros(self.s_end, dtype=torch.float32, device=device)
        self.shadow = torch.empty(self.w_end, dtype=torch.bfloat16, device=device)
        self.named = named
        for k, p in named.items():
            o, n = self.offsets[k], p.numel()
            view = self.master[o:o + n].view(p.shape)
            view.copy_(p.data)
            p.data = view
        self._ptrs = {k: p.data_ptr() for k, p in named.items()}
        self._plist = list(named.values())
        self._ptr_list = [p.data_ptr() for p in self._plist]
        self._version = -1
        self._updated = None            # (event, parameter versions) of the last note_updated()
        _LIVE_FLATS.add(self)
        _install_optimizer_hook()
        # gradient buckets for data-parallel training: one per encoder block, in flat-buffer order
        from . import ddp
        groups: Dict[str, List[str]] = {}
        self.block_of: Dict[str, str] = {}      # GEMM weight -> the block (= gradient bucket = shadow-refresh unit) holding it
        for k in order_w:
            parts = k.split(".")
            if parts[1] == "encoder":
                name = {"layer": "t", "v_layer": "v", "c_layer": "c"}[parts[2]] + parts[3]
            else:
                name = "tail"
            groups.setdefault(name, []).append(k)
            self.block_of[k] = name
        self.buckets = ddp.block_ranges(self.offsets, {k: named[k].numel() for k in named}, list(groups.items()), self.s_end)

    # fused q|k|v biases are laid out back to back: check rather than assume
    def check_contiguous(self, keys: List[str]):
        o = self.offsets[keys[0]]
        for k in keys:
            assert self.offsets[k] == o, (k, self.offsets[k], o)
            o += self.named[k].numel()

    # both run on every forward, between two graph launches: list comprehensions over a cached parameter list (a generator
    # with a dict lookup per element took 0.1 ms of the 0.4 ms the host needs before it can launch the forward graph)
    def intact(self) -> bool:
        return list(map(_DATA_PTR, self._plist)) == self._ptr_list          # unbound method through map: half the cost of p.data_ptr()

    def versions(self) -> int:
        return sum(map(_VERSION, self._plist))

    def w(self, key: str, rows: Optional[int] = None) -> torch.Tensor:
        """bf16 shadow of a GEMM weight (optionally `rows` rows starting at this key: fused q|k|v)."""
        p = self.named[key]
        r = p.shape[0] if rows is None else rows
        o = self.offsets[key]
        return self.shadow[o:o + r * p.shape[1]].view(r, p.shape[1])

    def m(self, key: str, numel: Optional[int] = None) -> torch.Tensor:
        """fp32 master view (1-D, optionally spanning fused neighbours)."""
        n = self.named[key].numel() if numel is None else numel
        o = self.offsets[key]
        return self.master[o:o + n]

    def g(self, key: str, shape=None, numel: Optional[int] = None) -> torch.Tensor:
        """fp32 gradient view."""
        p = self.named[key]
        n = p.numel() if numel is None else numel
        o = self.offsets[key]
        t = self.grad[o:o + n]
        return t.view(shape) if shape is not None else (t.view(p.shape) if numel is None else t)

    def refresh_shadow(self):
        ops.cast_bf16(self.master[:self.w_end], self.shadow)

    def note_updated(self, versions: Optional[int] = None) -> None:
        """Mark "every parameter write so far sits before this point of the current stream" (called from the optimizer
        post-step hook installed below, or by the owner of a hand-written update loop).  The next forward may then refresh the
        bf16 shadows on a side stream that waits for THIS event instead of for whatever the caller queued afterwards -- in the
        reference's loop (nodes.py:784-799) that is the host-to-device copy of the next batch, which the 0.2 ms refresh then
        overlaps instead of following.  Only honoured while no parameter version moved after the event."""
        ev = torch.cuda.Event()
        ev.record()
        self._updated = (ev, self.versions() if versions is None else versions)


# ---------------------------------------------------------------------------------------------------------------------
# the engine: static buffers + hand-scheduled kernel sequence
# ---------------------------------------------------------------------------------------------------------------------
class _Plan:
    """Everything that depends on the batch geometry: buffers, and the captured forward / backward graphs."""

    def __init__(self, eng: "_Engine", key):
        self.eng = eng
        (self.B, self.T, self.R, self.C, self.has_tmask, self.has_vmask, self.has_types, self.has_labels,
         self.dropout, self.need_grad) = key
        self.bufs: Dict[str, torch.Tensor] = {}
        self.twins: Dict[int, torch.Tensor] = {}     # bf16 activation (data_ptr) -> its fp32 twin (residual stream)
        self.ring: Dict[Any, int] = {}               # LayerNorms issued so far per activation shape (fp32 ring index)
        self.fwd_graph = None
        self.fwd_graph_r = None        # the forward graph that carries the bf16 shadow refresh
        self.bwd_graph = None
        self.fwd_runs = 0
        self.bwd_runs = 0
        self.fwd_launches = 0
        self.bwd_launches = 0
        self.fwd_id = 0
        dev = eng.flat.device
        # chain streams (text = s_main, visual = s_v) run the forward and the dgrad chain at high priority; the weight
        # gradients (off the critical path) go to low-priority side streams and fill the SMs the chain leaves idle
        self.s_main = torch.cuda.Stream(device=dev, priority=-1)
        self.s_v = torch.cuda.Stream(device=dev, priority=-1)
        self.s_tw = torch.cuda.Stream(device=dev)
        self.s_vw = torch.cuda.Stream(device=dev)
        self.side_events: Dict[str, torch.cuda.Event] = {}
        cfg = eng.cfg
        B, T, R = self.B, self.T, self.R
        self.Mt, self.Mv = B * T, B * R
        i32, f32 = torch.int32, torch.float32
        self.ids = self.buf("in.ids", (self.Mt,), i32)
        self.types = self.buf("in.types", (self.Mt,), i32) if self.has_types else None
        self.labels = self.buf("in.labels", (B,), i32) if self.has_labels else None
        self.t_bias = self.buf("in.t_bias", (B, T), f32) if self.has_tmask else None
        self.v_bias = self.buf("in.v_bias", (B, R), f32) if self.has_vmask else None
        self.feat = self.buf("in.feat", (self.Mv, cfg["v_feature_size"]))
        self.loc = self.buf("in.loc", (self.Mv, cfg["v_loc_size"]), f32)
        self.logits = self.buf("out.logits", (B, self.C), f32)
        self.probs = self.buf("out.probs", (B, self.C), f32)
        self.loss = self.buf("out.loss", (1,), f32)
        self.dloss = self.buf("in.dloss", (1,), f32)
        # the dropout seed THIS plan's last forward drew: its backward regenerates the masks from it, whatever other
        # geometries ran forward in between (the engine-wide seed has moved on by then)
        self.seed = self.buf("in.seed", (1,), torch.int64)
        self.dlogits = self.buf("in.dlogits", (B, self.C), f32)
        # constant one-hot GEMM operands of the vilbert_core surface: region index (position table) and sample index (mean pool)
        self.onehot_r = self.onehot_b = self.inv_r = None
        if V_POS_KEY in eng.flat.named:
            if R > eng.flat.named[V_POS_KEY].shape[0]:
                raise VbError(f"{R} regions exceed the region-position table ({eng.flat.named[V_POS_KEY].shape[0]} rows)")
            rp = (R + 7) // 8 * 8
            self.onehot_r = self.buf("const.onehot_r", (self.Mv, rp))
            self.onehot_r.view(B, R, rp)[:, torch.arange(R), torch.arange(R)] = 1.0
        if cfg.get("_v_pool", "first") == "mean":
            bp = (B + 7) // 8 * 8
            self.onehot_b = self.buf("const.onehot_b", (self.Mv, bp))
            self.onehot_b.view(B, R, bp)[torch.arange(B), :, torch.arange(B)] = 1.0
            self.inv_r = self.buf("const.inv_r", (cfg["v_hidden_size"],), f32)
            self.inv_r.fill_(1.0 / R)

    def buf(self, name, shape, dtype=torch.bfloat16) -> torch.Tensor:
        t = self.bufs.get(name)
        if t is None:
            t = torch.zeros(shape, dtype=dtype, device=self.eng.flat.device)
            self.bufs[name] = t
        assert tuple(t.shape) == tuple(shape) and t.dtype == dtype, (name, t.shape, shape)
        return t


class _MappedFlag:
    """Device-side address of a pinned host tensor (cudaHostGetDevicePointer; with unified addressing it is the host
    address itself).  Quacks like a tensor for ops._ptr / ops._need_cuda."""
    is_cuda = True

    def __init__(self, host: torch.Tensor):
        self._host = host
        self._ptr = host.data_ptr()

    def data_ptr(self) -> int:
        return self._ptr


class _Engine:
    def __init__(self, model: "ViLBERTForClassification", device):
        self.model = model
        self.cfg = model.config
        self.flat = _FlatParams(model, device)
        self.plans: Dict[Any, _Plan] = {}
        self.seed = torch.tensor([int(os.environ.get("VB_SEED", "20260101"))], dtype=torch.int64, device=device)
        self.use_graphs = os.environ.get("VB_NO_GRAPH", "0") != "1"
        self.two_streams = os.environ.get("VB_ONE_STREAM", "0") != "1"
        self.side_streams = self.two_streams and os.environ.get("VB_NO_SIDE", "0") != "1"
        self.wgrad_split = os.environ.get("VB_WGRAD_SPLIT", "0") == "1"   # measured slower in the full step (6.41 vs 5.91 ms): the 1 GB zero-fill evicts L2-resident activations
        # SM partitioning: weight-gradient GEMMs (side streams, off the critical path) are capped to a slice of the SMs so
        # that a dgrad-chain kernel never has to wait for a chip-wide weight-gradient grid to drain
        self.fp32_residual = os.environ.get("VB_BF16_RESIDUAL", "0") != "1"
        self.wgrad_ctas = int(os.environ.get("VB_WGRAD_CTAS", "74"))     # measured (round 2 kernels): 0 -> 4.93 ms, 32 -> 5.00, 48 -> 4.83, 74 -> 4.78, 100 -> 4.78 per step
        self._pl = None
        self.launches = 0
        # range-check verdict of the staging kernel: one int32 in mapped pinned host memory (see _raise_on_bad_indices)
        self.err_host = torch.zeros(1, dtype=torch.int32)
        if torch.device(device).type == "cuda":
            self.err_host = self.err_host.pin_memory()
        self.err_flag = _MappedFlag(self.err_host)
        self.err_np = self.err_host.numpy()          # the same word without building a tensor view per read
        self.strict_inputs = os.environ.get("VB_STRICT_INPUTS", "0") == "1"
        self.refresh_stream = (torch.cuda.Stream(device=device)
                               if torch.device(device).type == "cuda" and os.environ.get("VB_SYNC_REFRESH", "0") != "1" else None)
        # The bf16 shadow refresh as a branch of the FORWARD graph: block by block in execution order on the refresh stream, every
        # GEMM waiting only for its own block's cast -- 1.5 GB of HBM streaming beside the first layers' compute instead of 0.19 ms
        # in front of them.  VB_REFRESH_IN_GRAPH=0: one cast before the forward (round-2 baseline, A/B runs).
        self.refresh_in_graph = self.refresh_stream is not None and os.environ.get("VB_REFRESH_IN_GRAPH", "1") != "0"
        # the FFN-1 forward stores GELU'(pre-activation) instead of the pre-activation, so that the FFN-2 dgrad epilogue only
        # multiplies (it was MUFU-bound: 14.7 us against 11.8 for the same GEMM without it); VB_GELU_GRAD_FWD=0: as before
        self.gelu_grad_fwd = os.environ.get("VB_GELU_GRAD_FWD", "1") != "0"
        self.refresh_pending = False          # set by _ensure_engine: the next forward carries the refresh
        self._refresh_events: Optional[Dict[str, Any]] = None
        self._refresh_waited = set()
        self._refresh_now = False
        self.grads_clean = False     # set by ViLBERTForClassification.zero_grad(set_to_none=False)
        self.comm_group = getattr(model, "_ddp_group", None)   # data-parallel: see ddp.attach()
        self.comm_stream = torch.cuda.Stream(device=device, priority=-1) if self.comm_group is not None else None   # as urgent as the chain: NCCL CTAs must get SM slots while GEMMs are running
        self.comm_compress = getattr(model, "_ddp_compress", None) if self.comm_group is not None else None
        self.comm_switch = None
        self.widen_stream = None
        if self.comm_group is not None and getattr(model, "_ddp_transport", "nccl") == "switch":
            self.widen_stream = torch.cuda.Stream(device=device)      # bf16 -> fp32 of reduced ranges: HBM work, off the link stream
            from . import ddp
            # VB_DDP_FP32_BCAST=1: the fp32 gradient buffer becomes a symmetric buffer too and the reduce kernel broadcasts the
            # mean into it already widened.  Measured SLOWER at N = 2 (5.97 vs 5.55 ms/step): twice the bytes on the links,
            # own copy included (multicast stores loop back through the switch); the default broadcasts bf16 in place and
            # widens locally.
            fp32_bcast = os.environ.get("VB_DDP_FP32_BCAST", "0") == "1"
            self.comm_switch = ddp.SwitchExchange(self.flat.s_end, device, self.comm_group, fp32_out=fp32_bcast)
            if fp32_bcast:
                self.flat.grad = self.comm_switch.grad[:self.flat.s_end]
        self.comm_staging = (torch.empty(self.flat.s_end, dtype=torch.bfloat16, device=device)
                             if self.comm_compress == "bf16" and self.comm_switch is None else None)
        self._pending, self._pending_streams = [], set()
        self._site = 0
        f = self.flat
        for p in [f"bert.encoder.layer.{i}.attention.self" for i in range(self.cfg["num_hidden_layers"])] + \
                 [f"bert.encoder.v_layer.{i}.attention.self" for i in range(self.cfg["v_num_hidden_layers"])]:
            f.check_contiguous([p + ".query.weight", p + ".key.weight", p + ".value.weight"])
            f.check_contiguous([p + ".query.bias", p + ".key.bias", p + ".value.bias"])
        for i in range(self.cfg["num_co_attention_layers"]):
            p = f"bert.encoder.c_layer.{i}.biattention"
            for s in ("1", "2"):
                f.check_contiguous([f"{p}.query{s}.weight", f"{p}.key{s}.weight", f"{p}.value{s}.weight"])
                f.check_contiguous([f"{p}.query{s}.bias", f"{p}.key{s}.bias", f"{p}.value{s}.bias"])

    # ------------------------------------------------------------------------------------------------ primitives
    def _gelu_bwd_mode(self):
        """What the dgrad of a GELU layer does with the tensor its forward saved: multiply by it (it IS GELU') or by GELU' of it."""
        return ops.AUX_MUL if self.gelu_grad_fwd else ops.AUX_MUL_GELU_GRAD

    def _shadow(self, wkey, rows=None):
        """bf16 shadow of a GEMM weight for a FORWARD kernel on the current stream; during a refreshing forward the stream first
        waits (once per block) for the cast of the block that holds it."""
        if self._refresh_events is not None:
            block = self.flat.block_of[wkey]
            cur = torch.cuda.current_stream()
            mark = (cur.cuda_stream, block)
            if mark not in self._refresh_waited:
                cur.wait_event(self._refresh_events[block])
                self._refresh_waited.add(mark)
        return self.flat.w(wkey, rows)

    def _refresh_order(self) -> List[str]:
        """Blocks in the order the forward first touches them (run_forward)."""
        cfg = self.cfg
        order, c = ["tail", "t0"], 0
        if cfg["v_num_hidden_layers"] > 0:
            order.append("v0")              # the visual stream starts its first layer beside the first text layer
        for i in range(cfg["num_hidden_layers"]):
            order.append(f"t{i}")
            if i in CO_ATTENTION_TEXT_LAYERS and c < cfg["num_co_attention_layers"]:
                order += [f"v{c}", f"c{c}", f"v{c + 1}"]
                c += 1
        seen, out = set(), []
        for n in order + list(self.flat.buckets):
            if n in self.flat.buckets and n not in seen:
                seen.add(n)
                out.append(n)
        return out

    def _begin_refresh(self, s_t) -> None:
        f, rs = self.flat, self.refresh_stream
        rs.wait_stream(s_t)
        self._refresh_events, self._refresh_waited = {}, set()
        with torch.cuda.stream(rs):
            for name in self._refresh_order():
                lo, hi = f.buckets[name]
                hi = min(hi, f.w_end)
                if hi > lo:
                    ops.cast_bf16(f.master[lo:hi], f.shadow[lo:hi])
                ev = torch.cuda.Event()
                ev.record(rs)
                self._refresh_events[name] = ev

    def _linear(self, x, wkey, out, *, rows=None, act=ops.ACT_NONE, preact=None):
        f = self.flat
        w = self._shadow(wkey, rows)
        bkey = wkey[:-len("weight")] + "bias"
        # GELU layers keep GELU'(pre-activation) for the backward (self.gelu_grad_fwd): the forward epilogue has the tanh anyway
        ops.gemm(x, w, out, bias=f.m(bkey, w.shape[0]), act=act, preact=preact, b_streamed=True,
                 preact_grad=self.gelu_grad_fwd and preact is not None and act == ops.ACT_GELU)

    def _linear_bwd(self, dy, x, wkey, *, rows=None, dx=None, aux=None, aux_mode=ops.AUX_NONE, bias_grad=True):
        """dW = dy^T x (fp32, straight into the flat gradient buffer), db = colsum(dy) unless already produced by the
        kernel that made dy, dx = dy W (+ aux | * gelu'(aux))."""
        f = self.flat
        w = f.w(wkey, rows)
        cur = torch.cuda.current_stream()
        side = self._side_stream(cur)
        if side is not None:
            side.wait_stream(cur)     # dy is complete on the chain stream
        with torch.cuda.stream(side if side is not None else cur):
            # fp32 accumulate into the (pre-zeroed) flat gradient buffer: lets the GEMM split the long token dimension
            # over more CTAs (TMA reduce-add), see run_backward
            ops.gemm(dy, x, self._wgrad_out(wkey, tuple(w.shape), w.numel()), a_mn_major=True, b_mn_major=True,
                     accumulate=self.wgrad_split, d_streamed=True, max_ctas=self.wgrad_ctas)
            if bias_grad:
                bkey = wkey[:-len("weight")] + "bias"
                ops.colsum(dy, f.g(bkey, numel=w.shape[0]))
        if dx is not None:
            ops.gemm(dy, w, dx, b_mn_major=True, aux=aux, aux_mode=aux_mode, b_streamed=True)

    def _wgrad_out(self, wkey, shape, numel):
        """Where a weight gradient is written: the fp32 flat gradient buffer, or -- with the switch transport -- the bf16
        symmetric exchange buffer at the same offset (widened into the fp32 views after the all-reduce)."""
        if self.comm_switch is None:
            return self.flat.g(wkey, shape=shape, numel=numel)
        o = self.flat.offsets[wkey]
        return self.comm_switch.buf[o:o + numel].view(shape)

    def _side_stream(self, cur):
        pl = self._pl
        if pl is None or not self.side_streams:
            return None
        return pl.s_vw if cur == pl.s_v else pl.s_tw

    def _scratch_begin(self, pl, prefix):
        """The chain is about to overwrite scratch buffers `prefix.*`: wait for the weight-gradient GEMMs that read their
        previous contents (recorded by _scratch_end two layers ago: scratch is double-buffered by layer parity)."""
        ev = pl.side_events.get(prefix)
        if ev is not None:
            torch.cuda.current_stream().wait_event(ev)

    def _scratch_end(self, pl, prefix):
        side = self._side_stream(torch.cuda.current_stream())
        if side is not None:
            ev = torch.cuda.Event()
            ev.record(side)
            pl.side_events[prefix] = ev

    def _ln(self, pl, x, res, lnkey, y, tag, p_in=0.0, p_out=0.0):
        """LayerNorm block.  The residual stream is carried in fp32 next to the bf16 activations: every LayerNorm also
        writes its output unrounded (`tag.y32`), and reads its residual from the fp32 twin of `res` when there is one, so
        that the only bf16 roundings left are those of GEMM / attention operands -- the precision model of PyTorch's own
        bf16 autocast, whose error the golden fixtures carry as the yardstick."""
        f = self.flat
        mean, rstd = pl.buf(tag + ".mean", (x.shape[0],), torch.float32), pl.buf(tag + ".rstd", (x.shape[0],), torch.float32)
        site = self._next_site()
        drop = pl.dropout
        # the fp32 copy lives only until the next LayerNorm of the same chain (text / visual) has consumed it: a ring of
        # three buffers per shape instead of one per site keeps the fp32 stream L2-resident (62 distinct 6 MB buffers per
        # step evicted the activations: +0.24 ms); backward recomputes from the bf16 residual, whose 2^-9 relative difference
        # only perturbs the gradient at rounding level
        y32 = None
        if self.fp32_residual:
            k = pl.ring.get(tuple(y.shape), 0)
            pl.ring[tuple(y.shape)] = k + 1
            y32 = pl.buf(f"y32.{y.shape[0]}x{y.shape[1]}.{k % 3}", tuple(y.shape), torch.float32)
        res32 = pl.twins.get(res.data_ptr()) if (res is not None and self.fp32_residual) else None
        ops.layernorm_fwd(x, res, f.m(lnkey + ".weight"), f.m(lnkey + ".bias"), y, mean, rstd,
                          p_in=p_in if drop else 0.0, site_in=site, p_out=p_out if drop else 0.0, site_out=site + 1,
                          seed=pl.seed if drop else None, res32=res32, y32=y32)
        if y32 is not None:
            pl.twins[y.data_ptr()] = y32
        return (mean, rstd, site)

    def _ln_bwd(self, pl, dy, x, res, lnkey, saved, *, dx, dres, bias_key=None, p_in=0.0, p_out=0.0):
        f = self.flat
        mean, rstd, site = saved
        drop = pl.dropout
        res32 = None      # see _ln: the fp32 residual ring has been overwritten by now
        ops.layernorm_bwd(dy, x, res, f.m(lnkey + ".weight"), mean, rstd, dx=dx, dres=dres,
                          dgamma=f.g(lnkey + ".weight"), dbeta=f.g(lnkey + ".bias"),
                          dbias=f.g(bias_key) if bias_key else None,
                          p_in=p_in if drop else 0.0, site_in=site, p_out=p_out if drop else 0.0, site_out=site + 1,
                          seed=pl.seed if drop else None, res32=res32)

    def _bucket_ready(self, name, producers, flush=False):
        """Data-parallel: a gradient bucket has been written by `producers`.  Finished buckets are queued and exchanged in
        groups of >= ddp.FLUSH_BYTES with ONE coalesced NCCL launch per group, on the communication stream, overlapping the
        rest of the backward pass (27 separate all-reduces cost 2.6 ms of launch-bound NCCL time at N = 2, five grouped ones
        ~1.9 ms; tools/nccl_probe.py)."""
        if self.comm_group is None:
            return
        from . import ddp
        if name is not None:
            self._pending.append(self.flat.buckets[name])
            self._pending_streams.update(producers)
        nbytes = sum(hi - lo for lo, hi in self._pending) * 4
        threshold = ddp.SWITCH_FLUSH_BYTES if self.comm_switch is not None else ddp.FLUSH_BYTES
        if not self._pending or (not flush and nbytes < threshold):
            return
        for s in self._pending_streams:
            self.comm_stream.wait_stream(s)
        ranges = ddp.merge_ranges(self._pending)
        with torch.cuda.stream(self.comm_stream):
            if self.comm_switch is not None:
                sw, f = self.comm_switch, self.flat
                for lo, hi in ranges:
                    if hi > f.w_end:      # the small parameters were accumulated in fp32 (atomics): narrow that part once
                        s0 = max(lo, f.w_end)
                        ops.cast_bf16(f.grad[s0:hi], sw.buf[s0:hi])
                    sw.all_reduce_mean(lo, hi)
                    if sw.grad is None:                 # (else the fp32 mean has already landed in f.grad[lo:hi] on every rank)
                        # widen on its own stream: the next range's reduction (NVLink-bound) overlaps this HBM-bound pass
                        self.widen_stream.wait_stream(self.comm_stream)
                        with torch.cuda.stream(self.widen_stream):
                            ops.cast_f32(sw.buf[lo:hi], f.grad[lo:hi])
            else:
                ddp.all_reduce_mean_ranges(self.flat.grad, ranges, self.comm_group, self.comm_staging)
        self._pending, self._pending_streams = [], set()

    def _next_site(self):
        self._site += 2
        return self._site

    # ------------------------------------------------------------------------------------------------ layers
    def _bert_layer_fwd(self, pl, p, tag, x, M, S, H, heads, inter, bias, p_hidden, p_attn):
        """BertLayer.forward (reference :215-219)."""
        sv = {}
        qkv = pl.buf(tag + ".qkv", (M, 3 * H))
        self._linear(x, p + ".attention.self.query.weight", qkv, rows=3 * H)
        ctx = pl.buf(tag + ".ctx", (M, H))
        lse = pl.buf(tag + ".lse", (len(ops.attn_blocks(S)) * pl.B, heads, 128), torch.float32)
        sv["attn_site"] = self._next_site()
        ops.attention_fwd(qkv[:, :H], qkv[:, H:2 * H], qkv[:, 2 * H:], ctx, lse, batch=pl.B, heads=heads, sq=S, sk=S,
                          d=H // heads, mask_bias=bias, p_drop=p_attn if pl.dropout else 0.0, site=sv["attn_site"],
                          seed=pl.seed if pl.dropout else None)
        ao = pl.buf(tag + ".ao", (M, H))
        self._linear(ctx, p + ".attention.output.dense.weight", ao)
        a = pl.buf(tag + ".a", (M, H))
        sv["ln1"] = self._ln(pl, ao, x, p + ".attention.output.LayerNorm", a, tag + ".ln1", p_in=p_hidden)
        pre = pl.buf(tag + ".pre", (M, inter)) if pl.need_grad else None
        it = pl.buf(tag + ".int", (M, inter))
        self._linear(a, p + ".intermediate.dense.weight", it, act=ops.ACT_GELU, preact=pre)
        fo = pl.buf(tag + ".fo", (M, H))
        self._linear(it, p + ".output.dense.weight", fo)
        y = pl.buf(tag + ".y", (M, H))
        sv["ln2"] = self._ln(pl, fo, a, p + ".output.LayerNorm", y, tag + ".ln2", p_in=p_hidden)
        sv.update(x=x, qkv=qkv, ctx=ctx, lse=lse, ao=ao, a=a, pre=pre, it=it, fo=fo, y=y)
        return y, sv

    def _ffn_bwd(self, pl, pfx_int, pfx_out, lnkey, sv_ln, dy, fo, a, pre, it, M, H, inter, sc, p_hidden):
        """Backward of  y = LN(drop(dense_out(gelu(dense_int(a)))) + a);  returns grad wrt a."""
        g_fo = pl.buf(sc + ".g_fo", (M, H))
        g_ares = pl.buf(sc + ".g_ares", (M, H)) if pl.dropout else None
        self._ln_bwd(pl, dy, fo, a, lnkey, sv_ln, dx=g_fo, dres=g_ares, bias_key=pfx_out + ".bias", p_in=p_hidden)
        g_pre = pl.buf(sc + ".g_pre", (M, inter))
        self._linear_bwd(g_fo, it, pfx_out + ".weight", dx=g_pre, aux=pre, aux_mode=self._gelu_bwd_mode(), bias_grad=False)
        g_a = pl.buf(sc + ".g_a", (M, H))
        self._linear_bwd(g_pre, a, pfx_int + ".weight", dx=g_a, aux=g_ares if g_ares is not None else g_fo, aux_mode=ops.AUX_ADD)
        return g_a

    def _bert_layer_bwd(self, pl, p, sv, dy, dx_out, M, S, H, heads, inter, bias, p_hidden, p_attn, sc):
        self._scratch_begin(pl, sc)
        g_a = self._ffn_bwd(pl, p + ".intermediate.dense", p + ".output.dense", p + ".output.LayerNorm", sv["ln2"], dy,
                            sv["fo"], sv["a"], sv["pre"], sv["it"], M, H, inter, sc, p_hidden)
        g_ao = pl.buf(sc + ".g_ao", (M, H))
        g_xres = pl.buf(sc + ".g_xres", (M, H)) if pl.dropout else None
        self._ln_bwd(pl, g_a, sv["ao"], sv["x"], p + ".attention.output.LayerNorm", sv["ln1"], dx=g_ao, dres=g_xres,
                     bias_key=p + ".attention.output.dense.bias", p_in=p_hidden)
        g_ctx = pl.buf(sc + ".g_ctx", (M, H))
        self._linear_bwd(g_ao, sv["ctx"], p + ".attention.output.dense.weight", dx=g_ctx, bias_grad=False)
        g_qkv = pl.buf(sc + ".g_qkv", (M, 3 * H))
        qkv = sv["qkv"]
        ops.attention_bwd(g_ctx, qkv[:, :H], qkv[:, H:2 * H], qkv[:, 2 * H:], sv["lse"], g_qkv[:, :H], g_qkv[:, H:2 * H],
                          g_qkv[:, 2 * H:], batch=pl.B, heads=heads, sq=S, sk=S, d=H // heads, mask_bias=bias,
                          p_drop=p_attn if pl.dropout else 0.0, site=sv["attn_site"], seed=pl.seed if pl.dropout else None,
                          out=sv["ctx"])
        self._linear_bwd(g_qkv, sv["x"], p + ".attention.self.query.weight", rows=3 * H, dx=dx_out,
                         aux=g_xres if g_xres is not None else g_ao, aux_mode=ops.AUX_ADD)
        self._scratch_end(pl, sc)

    # ------------------------------------------------------------------------------------------------ forward
    def run_forward(self, pl: _Plan):
        cfg, f = self.cfg, self.flat
        self._site = 0
        self._pl = pl
        H, Hv, bi = cfg["hidden_size"], cfg["v_hidden_size"], cfg["bi_hidden_size"]
        I, Iv = cfg["intermediate_size"], cfg["v_intermediate_size"]
        nh, nhv = cfg["num_attention_heads"], cfg["v_num_attention_heads"]
        ph, pa = cfg["hidden_dropout_prob"], cfg["attention_probs_dropout_prob"]
        pvh = cfg["v_hidden_dropout_prob"]
        B, T, R, Mt, Mv = pl.B, pl.T, pl.R, pl.Mt, pl.Mv
        s_t = torch.cuda.current_stream()
        s_v = pl.s_v if self.two_streams else s_t
        sv = pl.saved = {}
        pl.ring.clear()
        pl.twins.clear()
        if pl.dropout:
            ops.seed_advance(self.seed, pl.seed)
        self._refresh_events = None          # (a forward that raised half-way must not leave its events to the next one)
        if self._refresh_now:
            self._begin_refresh(s_t)
        s_v.wait_stream(s_t)

        # text embeddings (transformers BertEmbeddings) | visual embeddings (reference :100-104)
        t = pl.buf("emb.t", (Mt, H))
        e = "bert.embeddings"
        mean, rstd = pl.buf("emb.mean", (Mt,), torch.float32), pl.buf("emb.rstd", (Mt,), torch.float32)
        sv["emb_site"] = self._next_site()
        ops.embed_text_fwd(pl.ids, pl.types, f.m(e + ".word_embeddings.weight").view(-1, H),
                           f.m(e + ".position_embeddings.weight").view(-1, H),
                           f.m(e + ".token_type_embeddings.weight").view(-1, H), f.m(e + ".LayerNorm.weight"),
                           f.m(e + ".LayerNorm.bias"), t, mean, rstd, B, T, p_out=ph if pl.dropout else 0.0,
                           site_out=sv["emb_site"], seed=pl.seed if pl.dropout else None,
                           y32=pl.buf("emb.t32", (Mt, H), torch.float32) if self.fp32_residual else None)
        if self.fp32_residual:
            pl.twins[t.data_ptr()] = pl.bufs["emb.t32"]      # consumed by the first text LayerNorm only
        with torch.cuda.stream(s_v):
            ve = "bert.v_embeddings"
            img = pl.buf("vemb.img", (Mv, Hv))
            self._linear(pl.feat, ve + ".image_embeddings.weight", img)
            loc = pl.buf("vemb.loc", (Mv, Hv))
            ops.loc_embed_fwd(pl.loc, f.m(ve + ".image_location_embeddings.weight").view(Hv, -1),
                              f.m(ve + ".image_location_embeddings.bias"), loc)
            if pl.onehot_r is not None:     # + position_embeddings[region]  (vilbert_core.py:470-476)
                res = pl.buf("vemb.res", (Mv, Hv))
                ops.gemm(pl.onehot_r[:, :R], self._shadow(V_POS_KEY)[:R], res, b_mn_major=True, aux=loc, aux_mode=ops.AUX_ADD)
                loc = res
            v = pl.buf("vemb.v", (Mv, Hv))
            sv["vemb_ln"] = self._ln(pl, img, loc, ve + ".LayerNorm", v, "vemb.ln", p_out=pvh)
            sv["vemb_res"] = loc

        c = 0
        for i in range(cfg["num_hidden_layers"]):
            t, sv[f"t{i}"] = self._bert_layer_fwd(pl, f"bert.encoder.layer.{i}", f"t{i}", t, Mt, T, H, nh, I, pl.t_bias, ph, ph)
            if i in CO_ATTENTION_TEXT_LAYERS and c < cfg["num_co_attention_layers"]:
                with torch.cuda.stream(s_v):
                    # the reference hands ONE dropout_prob (v_hidden_dropout_prob) to the visual BertLayer (:441-446)
                    v, sv[f"v{c}"] = self._bert_layer_fwd(pl, f"bert.encoder.v_layer.{c}", f"v{c}", v, Mv, R, Hv, nhv, Iv,
                                                         pl.v_bias, pvh, pvh)
                v, t, sv[f"c{c}"] = self._co_layer_fwd(pl, c, v, t, s_t, s_v)
                c += 1
        s_t.wait_stream(s_v)

        # poolers (:404-408), concat, classifier (:569-578), CE (:637-639)
        pooled = pl.buf("head.pooled", (B, bi + Hv))
        ops.gemm(t.view(B, T * H)[:, :H], self._shadow("bert.t_pooler.dense.weight"), pooled[:, :bi],
                 bias=f.m("bert.t_pooler.dense.bias"), act=ops.ACT_TANH)
        if pl.onehot_b is not None:         # mean over the regions (vilbert_core.py:581) instead of the first region (:404-408)
            vmean32 = pl.buf("head.vmean32", (B, Hv), torch.float32)
            ops.avgpool_nhwc(v.view(B, R, Hv), vmean32)
            v_in = ops.cast_bf16(vmean32, pl.buf("head.vmean", (B, Hv)))
        else:
            v_in = v.view(B, R * Hv)[:, :Hv]
        ops.gemm(v_in, self._shadow("bert.v_pooler.dense.weight"), pooled[:, bi:],
                 bias=f.m("bert.v_pooler.dense.bias"), act=ops.ACT_TANH)
        sv["head_site"] = self._next_site()
        pc = cfg.get("_classifier_dropout", 0.1)
        pooled_d = pooled
        if pl.dropout:
            pooled_d = ops.dropout(pooled, pl.buf("head.pooled_d", (B, bi + Hv)), pc, sv["head_site"], pl.seed)
        hid = pl.buf("head.hid", (B, bi))
        ops.gemm(pooled_d, self._shadow("classifier.1.weight"), hid, bias=f.m("classifier.1.bias"), act=ops.ACT_RELU)
        hid_d = hid
        if pl.dropout:
            hid_d = ops.dropout(hid, pl.buf("head.hid_d", (B, bi)), pc, sv["head_site"] + 1, pl.seed)
        ops.cls_ce_fwd(hid_d, f.m("classifier.4.weight").view(pl.C, bi), f.m("classifier.4.bias"), pl.labels, pl.logits,
                       pl.probs, pl.loss)
        sv.update(t_final=t, v_final=v, v_pool_in=v_in, pooled=pooled, pooled_d=pooled_d, hid=hid, hid_d=hid_d)
        if self._refresh_events is not None:
            s_t.wait_stream(self.refresh_stream)        # join the branch (every block was waited for long ago)
            self._refresh_events = None

    def _co_layer_fwd(self, pl, c, v, t, s_t, s_v):
        """CoAttentionLayer.forward (reference :377-394): BiAttention :253-294, BiOutput :324-338, two FFNs."""
        cfg = self.cfg
        H, Hv, bi = cfg["hidden_size"], cfg["v_hidden_size"], cfg["bi_hidden_size"]
        I, Iv = cfg["intermediate_size"], cfg["v_intermediate_size"]
        heads = cfg["v_num_attention_heads"]
        d = bi // heads
        ph, pa, pvh = cfg["hidden_dropout_prob"], cfg["attention_probs_dropout_prob"], cfg["v_hidden_dropout_prob"]
        B, T, R, Mt, Mv = pl.B, pl.T, pl.R, pl.Mt, pl.Mv
        p, tag = f"bert.encoder.c_layer.{c}", f"c{c}"
        drop = pl.dropout
        sv = {"v_in": v, "t_in": t}
        tqkv = pl.buf(tag + ".tqkv", (Mt, 3 * bi))
        self._linear(t, p + ".biattention.query2.weight", tqkv, rows=3 * bi)
        with torch.cuda.stream(s_v):
            vqkv = pl.buf(tag + ".vqkv", (Mv, 3 * bi))
            self._linear(v, p + ".biattention.query1.weight", vqkv, rows=3 * bi)
        s_t.wait_stream(s_v)
        s_v.wait_stream(s_t)
        sv["site_v"], sv["site_t"] = self._next_site(), self._next_site()
        # text FFN half on the text stream
        t_ctx = pl.buf(tag + ".t_ctx", (Mt, bi))
        t_lse = pl.buf(tag + ".t_lse", (len(ops.attn_blocks(T)) * B, heads, 128), torch.float32)
        ops.attention_fwd(tqkv[:, :bi], vqkv[:, bi:2 * bi], vqkv[:, 2 * bi:], t_ctx, t_lse, batch=B, heads=heads, sq=T,
                          sk=R, d=d, mask_bias=pl.v_bias, p_drop=pa if drop else 0.0, site=sv["site_t"],
                          seed=pl.seed if drop else None)
        t_bo = pl.buf(tag + ".t_bo", (Mt, H))
        self._linear(t_ctx, p + ".biOutput.dense2.weight", t_bo)
        t_att = pl.buf(tag + ".t_att", (Mt, H))
        sv["t_ln1"] = self._ln(pl, t_bo, t, p + ".biOutput.LayerNorm2", t_att, tag + ".t_ln1", p_in=ph)
        t_pre = pl.buf(tag + ".t_pre", (Mt, I)) if pl.need_grad else None
        t_int = pl.buf(tag + ".t_int", (Mt, I))
        self._linear(t_att, p + ".t_intermediate.dense.weight", t_int, act=ops.ACT_GELU, preact=t_pre)
        t_fo = pl.buf(tag + ".t_fo", (Mt, H))
        self._linear(t_int, p + ".t_output.dense.weight", t_fo)
        t_out = pl.buf(tag + ".t_out", (Mt, H))
        sv["t_ln2"] = self._ln(pl, t_fo, t_att, p + ".t_output.LayerNorm", t_out, tag + ".t_ln2", p_in=ph)
        with torch.cuda.stream(s_v):
            v_ctx = pl.buf(tag + ".v_ctx", (Mv, bi))
            v_lse = pl.buf(tag + ".v_lse", (len(ops.attn_blocks(R)) * B, heads, 128), torch.float32)
            ops.attention_fwd(vqkv[:, :bi], tqkv[:, bi:2 * bi], tqkv[:, 2 * bi:], v_ctx, v_lse, batch=B, heads=heads, sq=R,
                              sk=T, d=d, mask_bias=pl.t_bias, p_drop=pa if drop else 0.0, site=sv["site_v"],
                              seed=pl.seed if drop else None)
            v_bo = pl.buf(tag + ".v_bo", (Mv, Hv))
            self._linear(v_ctx, p + ".biOutput.dense1.weight", v_bo)
            v_att = pl.buf(tag + ".v_att", (Mv, Hv))
            sv["v_ln1"] = self._ln(pl, v_bo, v, p + ".biOutput.LayerNorm1", v_att, tag + ".v_ln1", p_in=ph)
            v_pre = pl.buf(tag + ".v_pre", (Mv, Iv)) if pl.need_grad else None
            v_int = pl.buf(tag + ".v_int", (Mv, Iv))
            self._linear(v_att, p + ".v_intermediate.dense.weight", v_int, act=ops.ACT_GELU, preact=v_pre)
            v_fo = pl.buf(tag + ".v_fo", (Mv, Hv))
            self._linear(v_int, p + ".v_output.dense.weight", v_fo)
            v_out = pl.buf(tag + ".v_out", (Mv, Hv))
            sv["v_ln2"] = self._ln(pl, v_fo, v_att, p + ".v_output.LayerNorm", v_out, tag + ".v_ln2", p_in=pvh)
        sv.update(tqkv=tqkv, vqkv=vqkv, t_ctx=t_ctx, t_lse=t_lse, t_bo=t_bo, t_att=t_att, t_pre=t_pre, t_int=t_int,
                  t_fo=t_fo, t_out=t_out, v_ctx=v_ctx, v_lse=v_lse, v_bo=v_bo, v_att=v_att, v_pre=v_pre, v_int=v_int,
                  v_fo=v_fo, v_out=v_out)
        return v_out, t_out, sv

    # ------------------------------------------------------------------------------------------------ backward
    def run_backward(self, pl: _Plan):
        cfg, f = self.cfg, self.flat
        H, Hv, bi = cfg["hidden_size"], cfg["v_hidden_size"], cfg["bi_hidden_size"]
        I, Iv = cfg["intermediate_size"], cfg["v_intermediate_size"]
        nh, nhv = cfg["num_attention_heads"], cfg["v_num_attention_heads"]
        ph, pa, pvh = cfg["hidden_dropout_prob"], cfg["attention_probs_dropout_prob"], cfg["v_hidden_dropout_prob"]
        B, T, R, Mt, Mv = pl.B, pl.T, pl.R, pl.Mt, pl.Mv
        sv = pl.saved
        self._pl = pl
        pl.side_events.clear()
        s_t = torch.cuda.current_stream()
        s_v = pl.s_v if self.two_streams else s_t
        sides_t = [pl.s_tw] if self.side_streams else []
        sides_v = [pl.s_vw] if self.side_streams else []
        # gradients that are accumulated with atomics start from zero (biases, LayerNorm, location weights, tables).  The word
        # embedding table is 94 of the 100 MB and is only touched by the LAST kernel of the pass: its fill runs on a side stream
        # (16 us off the head of the critical path), the small rest here
        o_word = f.offsets.get("bert.embeddings.word_embeddings.weight", f.s_end)
        word_zeroed = None
        if self.side_streams and f.w_end < o_word < f.s_end:
            f.grad[f.w_end:o_word].zero_()
            pl.s_tw.wait_stream(s_t)
            with torch.cuda.stream(pl.s_tw):
                f.grad[o_word:f.s_end].zero_()
                word_zeroed = torch.cuda.Event()
                word_zeroed.record(pl.s_tw)
        else:
            f.grad[f.w_end:f.s_end].zero_()
        if self.wgrad_split:
            # weight gradients are produced by split-K GEMMs that reduce-add into the buffer: zero it once, on the side
            # stream(s) that own it, while the head of the backward chain runs
            if self.side_streams:
                pl.s_tw.wait_stream(s_t)
                with torch.cuda.stream(pl.s_tw):
                    f.grad[:f.w_end].zero_()
                pl.s_vw.wait_stream(pl.s_tw)
            else:
                f.grad[:f.w_end].zero_()
        dy_t = [pl.buf("g.t_ping", (Mt, H)), pl.buf("g.t_pong", (Mt, H))]
        dy_v = [pl.buf("g.v_ping", (Mv, Hv)), pl.buf("g.v_pong", (Mv, Hv))]
        dy_t[0].zero_()
        dy_v[0].zero_()

        # head
        g_hid_d = pl.buf("g.hid_d", (B, bi))
        ops.cls_ce_bwd(sv["hid_d"], f.m("classifier.4.weight").view(pl.C, bi), pl.labels, pl.probs, pl.dloss, pl.dlogits,
                       f.g("classifier.4.weight"), f.g("classifier.4.bias"), g_hid_d)
        g_hid = g_hid_d
        if pl.dropout:
            g_hid = ops.dropout(g_hid_d, pl.buf("g.hid", (B, bi)), cfg.get("_classifier_dropout", 0.1), sv["head_site"] + 1, pl.seed)
        g_hid_pre = ops.act_bwd(g_hid, sv["hid"], pl.buf("g.hid_pre", (B, bi)), ops.ACT_RELU)
        g_pooled_d = pl.buf("g.pooled_d", (B, bi + Hv))
        self._linear_bwd(g_hid_pre, sv["pooled_d"], "classifier.1.weight", dx=g_pooled_d)
        g_pooled = g_pooled_d
        if pl.dropout:
            g_pooled = ops.dropout(g_pooled_d, pl.buf("g.pooled", (B, bi + Hv)), cfg.get("_classifier_dropout", 0.1), sv["head_site"], pl.seed)
        g_pool_pre = ops.act_bwd(g_pooled, sv["pooled"], pl.buf("g.pool_pre", (B, bi + Hv)), ops.ACT_TANH)
        s_v.wait_stream(s_t)
        self._linear_bwd(g_pool_pre[:, :bi], sv["t_final"].view(B, T * H)[:, :H], "bert.t_pooler.dense.weight",
                         dx=dy_t[0].view(B, T * H)[:, :H])
        with torch.cuda.stream(s_v):
            if pl.onehot_b is not None:     # mean pool: every region receives dpooled / R (one-hot GEMM with a column scale)
                g_vmean = pl.buf("g.vmean", (B, Hv))
                self._linear_bwd(g_pool_pre[:, bi:], sv["v_pool_in"], "bert.v_pooler.dense.weight", dx=g_vmean)
                ops.gemm(pl.onehot_b[:, :B], g_vmean, dy_v[0], b_mn_major=True, scale=pl.inv_r)
            else:
                self._linear_bwd(g_pool_pre[:, bi:], sv["v_pool_in"], "bert.v_pooler.dense.weight",
                                 dx=dy_v[0].view(B, R * Hv)[:, :Hv])

        # encoder, reversed
        it_, iv_ = 0, 0   # which ping/pong buffer currently holds the incoming gradient
        n_co = min(cfg["num_co_attention_layers"], sum(1 for i in range(cfg["num_hidden_layers"]) if i in CO_ATTENTION_TEXT_LAYERS))
        c = n_co
        for i in reversed(range(cfg["num_hidden_layers"])):
            if i in CO_ATTENTION_TEXT_LAYERS and c > 0 and self._co_index(i) < n_co:
                c -= 1
                self._co_layer_bwd(pl, c, sv[f"c{c}"], dy_v[iv_], dy_t[it_], dy_v[1 - iv_], dy_t[1 - it_], s_t, s_v)
                iv_, it_ = 1 - iv_, 1 - it_
                self._bucket_ready(f"c{c}", [s_t, s_v] + sides_t + sides_v)
                with torch.cuda.stream(s_v):
                    self._bert_layer_bwd(pl, f"bert.encoder.v_layer.{c}", sv[f"v{c}"], dy_v[iv_], dy_v[1 - iv_], Mv, R, Hv,
                                         nhv, Iv, pl.v_bias, pvh, pvh, f"gv{c & 1}")
                iv_ = 1 - iv_
                self._bucket_ready(f"v{c}", [s_v] + sides_v)
            self._bert_layer_bwd(pl, f"bert.encoder.layer.{i}", sv[f"t{i}"], dy_t[it_], dy_t[1 - it_], Mt, T, H, nh, I,
                                 pl.t_bias, ph, ph, f"gt{i & 1}")
            it_ = 1 - it_
            self._bucket_ready(f"t{i}", [s_t] + sides_t)

        # embeddings
        with torch.cuda.stream(s_v):
            ve = "bert.v_embeddings"
            g_s = pl.buf("g.vemb_s", (Mv, Hv))
            self._ln_bwd(pl, dy_v[iv_], pl.bufs["vemb.img"], sv["vemb_res"], ve + ".LayerNorm", sv["vemb_ln"], dx=g_s,
                         dres=None, bias_key=ve + ".image_embeddings.bias", p_out=pvh)
            self._linear_bwd(g_s, pl.feat, ve + ".image_embeddings.weight", bias_grad=False)
            if pl.onehot_r is not None:     # d position table = onehot^T g_s: the ordinary weight-gradient contraction
                n_pos = f.named[V_POS_KEY].shape[0]
                ops.gemm(pl.onehot_r[:, :R], g_s, self._wgrad_out(V_POS_KEY, (n_pos, Hv), n_pos * Hv)[:R], a_mn_major=True,
                         b_mn_major=True, d_streamed=True)
            ops.loc_embed_bwd(g_s, pl.loc, f.g(ve + ".image_location_embeddings.weight"),
                              f.g(ve + ".image_location_embeddings.bias"))
        e = "bert.embeddings"
        if word_zeroed is not None:
            s_t.wait_event(word_zeroed)
        ops.embed_text_bwd(dy_t[it_], pl.ids, pl.types, f.m(e + ".word_embeddings.weight").view(-1, H),
                           f.m(e + ".position_embeddings.weight").view(-1, H),
                           f.m(e + ".token_type_embeddings.weight").view(-1, H), f.m(e + ".LayerNorm.weight"),
                           pl.bufs["emb.mean"], pl.bufs["emb.rstd"], B, T,
                           dword=f.g(e + ".word_embeddings.weight"), dpos=f.g(e + ".position_embeddings.weight"),
                           dtype=f.g(e + ".token_type_embeddings.weight"), dgamma=f.g(e + ".LayerNorm.weight"),
                           dbeta=f.g(e + ".LayerNorm.bias"), p_out=ph if pl.dropout else 0.0, site_out=sv["emb_site"],
                           seed=pl.seed if pl.dropout else None)
        s_t.wait_stream(s_v)
        for side in sides_t + sides_v:
            s_t.wait_stream(side)
        self._bucket_ready("tail", [s_t], flush=True)
        if self.comm_stream is not None:
            s_t.wait_stream(self.comm_stream)
        if self.widen_stream is not None:
            s_t.wait_stream(self.widen_stream)

    @staticmethod
    def _co_index(text_layer):
        return CO_ATTENTION_TEXT_LAYERS.index(text_layer)

    def _co_layer_bwd(self, pl, c, sv, dv, dt, dv_out, dt_out, s_t, s_v):
        cfg = self.cfg
        H, Hv, bi = cfg["hidden_size"], cfg["v_hidden_size"], cfg["bi_hidden_size"]
        I, Iv = cfg["intermediate_size"], cfg["v_intermediate_size"]
        heads = cfg["v_num_attention_heads"]
        d = bi // heads
        ph, pa, pvh = cfg["hidden_dropout_prob"], cfg["attention_probs_dropout_prob"], cfg["v_hidden_dropout_prob"]
        B, T, R, Mt, Mv = pl.B, pl.T, pl.R, pl.Mt, pl.Mv
        p = f"bert.encoder.c_layer.{c}"
        drop = pl.dropout
        seed = pl.seed if drop else None
        sct, scv = f"gct{c & 1}", f"gcv{c & 1}"              # scratch double-buffered by layer parity
        self._scratch_begin(pl, sct)                         # both chains write both qkv-gradient buffers
        self._scratch_begin(pl, scv)
        with torch.cuda.stream(s_v):
            self._scratch_begin(pl, sct)
            self._scratch_begin(pl, scv)
        g_tqkv = pl.buf(sct + ".tqkv", (Mt, 3 * bi))
        g_vqkv = pl.buf(scv + ".vqkv", (Mv, 3 * bi))
        tqkv, vqkv = sv["tqkv"], sv["vqkv"]
        # text half down to the gradient of t_ctx
        g_tatt = self._ffn_bwd(pl, p + ".t_intermediate.dense", p + ".t_output.dense", p + ".t_output.LayerNorm", sv["t_ln2"],
                               dt, sv["t_fo"], sv["t_att"], sv["t_pre"], sv["t_int"], Mt, H, I, sct, ph)
        g_tbo = pl.buf(sct + ".g_bo", (Mt, H))
        g_tres = pl.buf(sct + ".g_res", (Mt, H)) if drop else None
        self._ln_bwd(pl, g_tatt, sv["t_bo"], sv["t_in"], p + ".biOutput.LayerNorm2", sv["t_ln1"], dx=g_tbo, dres=g_tres,
                     bias_key=p + ".biOutput.dense2.bias", p_in=ph)
        g_tctx = pl.buf(sct + ".g_ctx", (Mt, bi))
        self._linear_bwd(g_tbo, sv["t_ctx"], p + ".biOutput.dense2.weight", dx=g_tctx, bias_grad=False)
        with torch.cuda.stream(s_v):
            g_vatt = self._ffn_bwd(pl, p + ".v_intermediate.dense", p + ".v_output.dense", p + ".v_output.LayerNorm",
                                   sv["v_ln2"], dv, sv["v_fo"], sv["v_att"], sv["v_pre"], sv["v_int"], Mv, Hv, Iv, scv, pvh)
            g_vbo = pl.buf(scv + ".g_bo", (Mv, Hv))
            g_vres = pl.buf(scv + ".g_res", (Mv, Hv)) if drop else None
            self._ln_bwd(pl, g_vatt, sv["v_bo"], sv["v_in"], p + ".biOutput.LayerNorm1", sv["v_ln1"], dx=g_vbo, dres=g_vres,
                         bias_key=p + ".biOutput.dense1.bias", p_in=ph)
            g_vctx = pl.buf(scv + ".g_ctx", (Mv, bi))
            self._linear_bwd(g_vbo, sv["v_ctx"], p + ".biOutput.dense1.weight", dx=g_vctx, bias_grad=False)
            # regions attend to tokens: dq -> visual q1, dk/dv -> text k2/v2
            ops.attention_bwd(g_vctx, vqkv[:, :bi], tqkv[:, bi:2 * bi], tqkv[:, 2 * bi:], sv["v_lse"], g_vqkv[:, :bi],
                              g_tqkv[:, bi:2 * bi], g_tqkv[:, 2 * bi:], batch=B, heads=heads, sq=R, sk=T, d=d,
                              mask_bias=pl.t_bias, p_drop=pa if drop else 0.0, site=sv["site_v"], seed=seed, out=sv["v_ctx"])
        # tokens attend to regions: dq -> text q2, dk/dv -> visual k1/v1
        ops.attention_bwd(g_tctx, tqkv[:, :bi], vqkv[:, bi:2 * bi], vqkv[:, 2 * bi:], sv["t_lse"], g_tqkv[:, :bi],
                          g_vqkv[:, bi:2 * bi], g_vqkv[:, 2 * bi:], batch=B, heads=heads, sq=T, sk=R, d=d,
                          mask_bias=pl.v_bias, p_drop=pa if drop else 0.0, site=sv["site_t"], seed=seed, out=sv["t_ctx"])
        s_t.wait_stream(s_v)
        s_v.wait_stream(s_t)
        self._linear_bwd(g_tqkv, sv["t_in"], p + ".biattention.query2.weight", rows=3 * bi, dx=dt_out,
                         aux=g_tres if g_tres is not None else g_tbo, aux_mode=ops.AUX_ADD)
        self._scratch_end(pl, sct)
        with torch.cuda.stream(s_v):
            self._linear_bwd(g_vqkv, sv["v_in"], p + ".biattention.query1.weight", rows=3 * bi, dx=dv_out,
                             aux=g_vres if g_vres is not None else g_vbo, aux_mode=ops.AUX_ADD)
            self._scratch_end(pl, scv)

    # ------------------------------------------------------------------------------------------------ execution
    def _execute(self, pl: _Plan, which: str):
        fn = self.run_forward if which == "fwd" else self.run_backward
        runs = pl.fwd_runs if which == "fwd" else pl.bwd_runs
        if which == "fwd":
            # a forward that carries the shadow refresh is a graph of its own (the refresh branch + one event wait per block)
            self._refresh_now, self.refresh_pending = self.refresh_pending, False
            slot = "fwd_graph_r" if self._refresh_now else "fwd_graph"
            pl.fwd_runs += 1
        else:
            slot = "bwd_graph"
            pl.bwd_runs += 1
        graph = getattr(pl, slot)
        caller = torch.cuda.current_stream()
        pl.s_main.wait_stream(caller)
        if not self.use_graphs or runs == 0:
            lc = _lib.launch_count()
            with torch.cuda.stream(pl.s_main):
                fn(pl)  # eager (also the warm-up that loads modules / sets function attributes before any capture)
            setattr(pl, which + "_launches", _lib.launch_count() - lc)
            caller.wait_stream(pl.s_main)
            return
        if graph is None:
            torch.cuda.synchronize()
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph, stream=pl.s_main, capture_error_mode="thread_local"):
                fn(pl)
            setattr(pl, slot, graph)
        with torch.cuda.stream(pl.s_main):
            graph.replay()
        caller.wait_stream(pl.s_main)


class _Step(torch.autograd.Function):
    """One node in the autograd graph for the whole model: forward returns (logits, loss); backward runs the
    hand-written backward pass and deposits fp32 gradients on the parameters (views of the flat gradient buffer)."""

    @staticmethod
    def forward(ctx, anchor, module, plan):
        eng = module._engine
        eng._execute(plan, "fwd")
        plan.fwd_id += 1
        ctx.module, ctx.plan, ctx.fwd_id = module, plan, plan.fwd_id
        return plan.logits.clone(), plan.loss[0].clone()

    @staticmethod
    def backward(ctx, g_logits, g_loss):
        module, plan = ctx.module, ctx.plan
        eng = module._engine
        if plan.fwd_id != ctx.fwd_id:
            raise VbError("backward() of a forward whose saved activations were overwritten by a later forward of the "
                          "same shape; call backward before the next forward")
        if g_logits is None:
            plan.dlogits.zero_()
        else:
            plan.dlogits.copy_(g_logits)
        if g_loss is None:
            plan.dloss.zero_()
        else:
            plan.dloss.copy_(g_loss.reshape(1))
        module._raise_on_bad_indices(eng)
        carry = None
        params = module._grad_bindings()                     # [(key, parameter, view into the flat gradient buffer)], cached
        if not eng.grads_clean and any(p.grad is not None for _, p, _ in params):
            carry = {k: p.grad.clone() for k, p, _ in params if p.grad is not None}   # gradient accumulation (rare path)
        eng.grads_clean = False
        eng._execute(plan, "bwd")
        for k, p, g in params:
            if carry is not None and k in carry:
                g.add_(carry[k])
            p.grad = g
        return None, None, None


class ViLBERTForClassification(nn.Module):
    """Drop-in for the reference class of the same name (vilbert_facebook_arch.py:554-641): same constructor, same
    parameter names / shapes / ``state_dict``, same keyword ``forward`` returning ``{"logits", "loss"?}``, same
    ``get_num_parameters`` / ``freeze_bert_layers``.  CUDA only: a forward on CPU tensors raises."""

    def __init__(self, config: Dict[str, Any], num_labels: int = 2):
        super().__init__()
        self.config = config
        self.num_labels = num_labels
        self.bert = _Backbone(config)
        cls_in = config["bi_hidden_size"] + config["v_hidden_size"]
        self.classifier = nn.Sequential(nn.Dropout(0.1), nn.Linear(cls_in, config["bi_hidden_size"]), nn.ReLU(),
                                        nn.Dropout(0.1), nn.Linear(config["bi_hidden_size"], num_labels))
        self._engine: Optional[_Engine] = None
        self._anchor = None
        self._ddp_group = None
        self._ddp_compress = None
        if config["bi_hidden_size"] != config["v_hidden_size"]:
            raise VbError("bi_hidden_size must equal v_hidden_size (as in the reference's v_pooler / BiOutput)")

    # -- reference surface ------------------------------------------------------------------------------------------
    def get_num_parameters(self) -> Tuple[int, int]:
        total = sum(p.numel() for p in self.parameters())
        trainable = sum(p.numel() for p in self.parameters() if p.requires_grad)
        return total, trainable

    def freeze_bert_layers(self, num_layers: int = 6) -> None:
        """Reference :586-608: freeze text embeddings and the first N text layers."""
        if num_layers <= 0:
            return
        for p in self.bert.embeddings.parameters():
            p.requires_grad = False
        for i, layer in enumerate(self.bert.encoder.layer):
            if i < num_layers:
                for p in layer.parameters():
                    p.requires_grad = False

    def _apply(self, fn, *a, **k):
        out = super()._apply(fn, *a, **k)
        self._engine = None   # parameters were re-created (device / dtype move): re-flatten lazily
        return out

    def _trainable_used(self):
        eng = self._engine
        return [(k, p) for k, p in eng.flat.named.items() if p.requires_grad and k in eng.flat.used]

    def _grad_bindings(self):
        """(key, parameter, gradient view) for every trainable parameter the forward uses; rebuilt only when the engine or a
        requires_grad flag changes (building 469 views per backward costs more host time than the whole forward launch)."""
        eng = self._engine
        sig = (id(eng.flat), tuple(p.requires_grad for p in eng.flat.named.values()))
        cache = getattr(eng, "_grad_cache", None)
        if cache is None or cache[0] != sig:
            cache = (sig, [(k, p, eng.flat.g(k)) for k, p in self._trainable_used()])
            eng._grad_cache = cache
        return cache[1]

    def _raise_on_bad_indices(self, eng) -> None:
        """The staging kernel range-checks ids / token types / labels the way nn.Embedding and nn.CrossEntropyLoss do and
        ORs a bit into a flag that lives in mapped pinned host memory: reading it costs nothing and needs no stream
        synchronisation, so the verdict on batch k is raised at the latest when batch k+1 arrives or batch k's backward
        starts (VB_STRICT_INPUTS=1: synchronise and raise inside the same forward).  Offending values were clamped, so no
        kernel indexed out of bounds in the meantime."""
        bits = int(eng.err_np[0])
        if bits:
            eng.err_np[0] = 0
            what = [n for b, n in ((_lib.STAGE_ERR_ID, f"input_ids outside [0, {self.config['vocab_size']})"),
                                   (_lib.STAGE_ERR_TYPE, "token_type_ids outside [0, 2)"),
                                   (_lib.STAGE_ERR_LABEL, f"labels outside [0, {self.num_labels}) and != -100")) if bits & b]
            raise VbError("index out of range in the staged batch: " + "; ".join(what))

    def zero_grad(self, set_to_none: bool = True) -> None:
        """nn.Module.zero_grad; with set_to_none=False the gradients that are views of the engine's flat buffer are cleared with
        one fill, and the next backward knows it has nothing to carry over (no per-parameter clone + add)."""
        eng = self._engine
        if set_to_none or eng is None:
            return super().zero_grad(set_to_none=set_to_none)
        own = {g.data_ptr() for _, _, g in self._grad_bindings()}
        foreign = [p for p in self.parameters() if p.grad is not None and p.grad.data_ptr() not in own]
        eng.flat.grad.zero_()
        for p in foreign:
            p.grad.zero_()
        eng.grads_clean = not foreign

    def _ensure_engine(self, device) -> _Engine:
        eng = self._engine
        if eng is not None and (not eng.flat.intact() or eng.flat.device != device):
            eng = None
        if eng is None:
            for p in self.parameters():
                if p.device != device:
                    raise VbError(f"parameter on {p.device}, inputs on {device}: call model.to(device) first")
                if p.dtype != torch.float32:
                    raise VbError("parameters must stay fp32 (bf16 shadows are kept internally)")
            eng = self._engine = _Engine(self, device)
            self._anchor = torch.zeros(1, device=device, requires_grad=True)
        ver = eng.flat.versions()
        if ver != eng.flat._version:
            upd = eng.flat._updated
            # Two regimes.  The GPU is still busy with the previous step (the host runs ahead): the forward about to run carries
            # the refresh as a branch of its graph, block by block beside its first layers.  The GPU has already drained up to
            # the optimizer's update (the reference's loop reads loss.item() every step, so the host is the late one): cast now,
            # on the side stream, beside the batch copy and the rest of this call's host work -- the window is free anyway.
            idle = upd is not None and upd[1] == ver and eng.refresh_stream is not None and upd[0].query()
            if eng.refresh_in_graph and not idle:
                eng.refresh_pending = True
            elif upd is not None and upd[1] == ver and eng.refresh_stream is not None:
                # nothing touched the parameters after the optimizer's hook: refresh beside whatever was queued since
                # (the next batch's host-to-device copy), the forward waits for it below
                eng.refresh_stream.wait_event(upd[0])
                with torch.cuda.stream(eng.refresh_stream):
                    eng.flat.refresh_shadow()
                torch.cuda.current_stream().wait_stream(eng.refresh_stream)
            else:
                eng.flat.refresh_shadow()
            eng.flat._updated = None
            eng.flat._version = ver
        return eng

    def parameters_updated(self) -> None:
        """For training loops that update the parameters with their own kernels (raw pointers, no version bump, no torch
        optimizer): call right after the update.  The next forward refreshes the bf16 weight shadows, ordered after this
        point of the current stream.  ``torch.optim`` optimizers need no call (post-step hook), nor does FusedAdamW."""
        eng = self._engine
        if eng is not None:
            eng.flat.note_updated()
            eng.flat._version = -1

    def forward(self, input_ids, attention_mask=None, token_type_ids=None, visual_features=None,
                visual_attention_mask=None, spatial_locations=None, labels=None, *, image_feat=None, image_loc=None,
                image_attention_mask=None):
        # aliases named by the north-star signature
        visual_features = image_feat if visual_features is None else visual_features
        spatial_locations = image_loc if spatial_locations is None else spatial_locations
        visual_attention_mask = image_attention_mask if visual_attention_mask is None else visual_attention_mask
        if visual_features is None or spatial_locations is None:
            raise VbError("visual_features and spatial_locations are required")
        if not input_ids.is_cuda:
            raise VbError("ViLBERTForClassification (B200) runs on CUDA tensors only; there is no CPU fallback")
        device = input_ids.device
        cfg = self.config
        on_current = device.type == "cuda" and torch.cuda.current_device() == device.index
        with (contextlib.nullcontext() if on_current else torch.cuda.device(device)):       # (the context manager costs ~15 us)
            eng = self._ensure_engine(device)
            B, T = input_ids.shape
            R = visual_features.shape[1]
            if T > cfg["max_position_embeddings"]:
                raise VbError(f"{T} tokens exceed max_position_embeddings = {cfg['max_position_embeddings']} "
                              "(the position table has no such row)")
            if T > ops.ATTN_MAX_SEQ or R > ops.ATTN_MAX_SEQ:
                raise VbError(f"sequence lengths above {ops.ATTN_MAX_SEQ} are not supported by the blocked attention (T={T}, R={R})")
            if visual_features.shape[2] != cfg["v_feature_size"] or spatial_locations.shape[-1] != cfg["v_loc_size"]:
                raise VbError("visual feature / location width does not match the configuration")
            need_grad = torch.is_grad_enabled() and any(p.requires_grad for p in eng.flat._plist)
            dropout = bool(self.training)
            key = (B, T, R, self.num_labels, attention_mask is not None, visual_attention_mask is not None,
                   token_type_ids is not None, labels is not None, dropout, need_grad)
            pl = eng.plans.get(key)
            if pl is None:
                pl = eng.plans[key] = _Plan(eng, key)
            # stage the batch into the plan's static buffers (the graphs read these addresses): ONE launch
            self._raise_on_bad_indices(eng)          # verdict of the PREVIOUS batch's range checks (no sync on this one)
            L = _lib
            # (the staging launch reads pointer / element count / dtype of contiguous tensors: no flattened views are built)
            segs = [(L.STAGE_INDEX, input_ids, pl.ids, 0, cfg["vocab_size"], L.STAGE_ERR_ID)]
            if token_type_ids is not None:
                segs.append((L.STAGE_INDEX, token_type_ids, pl.types, 0, 2, L.STAGE_ERR_TYPE))
            if labels is not None:
                segs.append((L.STAGE_INDEX, labels, pl.labels, 0, self.num_labels, L.STAGE_ERR_LABEL))
            if attention_mask is not None:
                segs.append((L.STAGE_MASK, attention_mask, pl.t_bias, 0, 0, 0))
            if visual_attention_mask is not None:
                segs.append((L.STAGE_MASK, visual_attention_mask, pl.v_bias, 0, 0, 0))
            segs.append((L.STAGE_FEAT, visual_features, pl.feat, 0, 0, 0))
            segs.append((L.STAGE_COPY_F32, spatial_locations, pl.loc, 0, 0, 0))
            ops.stage_batch([(k, src if src.is_contiguous() else src.contiguous(), dst, lo, hi, bit)
                             for k, src, dst, lo, hi, bit in segs], eng.err_flag)
            if eng.strict_inputs:
                torch.cuda.current_stream().synchronize()
                self._raise_on_bad_indices(eng)
            if need_grad:
                logits, loss = _Step.apply(self._anchor, self, pl)
            else:
                eng._execute(pl, "fwd")
                logits, loss = pl.logits.clone(), pl.loss[0].clone()
        out = {"logits": logits}
        if labels is not None:
            out["loss"] = loss
        return out


def load_facebook_weights(model: ViLBERTForClassification, checkpoint_path: str) -> int:
    """Same contract as the reference's loader (vilbert_facebook_arch.py:644-683): copy every checkpoint tensor whose key
    and shape match, ignore the rest, return the number loaded."""
    state = torch.load(checkpoint_path, map_location="cpu")
    own = model.state_dict()
    picked = {k: v for k, v in state.items() if k in own and own[k].shape == v.shape}
    model.load_state_dict(picked, strict=False)
    return len(picked)
