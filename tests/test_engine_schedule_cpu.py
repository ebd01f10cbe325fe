"""The ViLBERT engine's HOST schedule (multimodal_classification_b200/vilbert.py: forward and backward kernel sequence, flat
parameter / gradient layout, fp32 residual ring, bucket order) run in the GPU-less container over the functional stand-ins of
tests/ops_sim.py and compared with the fp32 oracle: logits, loss and every parameter gradient.  The kernels themselves are
tested on the B200 (tests/test_*_gpu.py); this covers what sits above them."""
import pytest
import torch

from oracle import vilbert_oracle as vo


@pytest.fixture
def simulated(monkeypatch):
    if torch.cuda.is_available():
        pytest.skip("the stand-ins are for the GPU-less container")
    import ops_sim
    ops_sim.install(monkeypatch)


def _run(cfg, batch_kw, train=False):
    from multimodal_classification_b200.vilbert import ViLBERTForClassification
    sd = vo.seeded_state_dict(cfg)
    batch = vo.synthetic_batch(cfg, **batch_kw)
    model = ViLBERTForClassification(cfg, num_labels=2)
    model.load_state_dict(sd, strict=True)
    model.eval()
    out = model(**batch)
    out["loss"].backward()
    ref_out, ref_grads = vo.loss_and_grads(sd, cfg, batch)
    return model, out, ref_out, ref_grads


@pytest.mark.parametrize("batch_kw", [dict(batch=2, seq=24, regions=12, seed=3),
                                      dict(batch=3, seq=16, regions=9, seed=5, with_visual_mask=True, with_token_types=False)])
def test_engine_schedule_matches_oracle(simulated, batch_kw):
    cfg = vo.tiny_config()
    model, out, ref_out, ref_grads = _run(cfg, batch_kw)
    scale = ref_out["logits"].abs().max().item()
    assert (out["logits"].float() - ref_out["logits"]).abs().max().item() <= 2e-2 * scale
    assert abs(out["loss"].item() - ref_out["loss"].item()) <= 1e-3
    worst = (1.0, "")
    for k, p in model.named_parameters():
        g_ref = ref_grads[k]
        if g_ref is None:
            assert p.grad is None, k
            continue
        assert p.grad is not None, k
        if ".key" in k and k.endswith(".bias"):
            assert p.grad.abs().max().item() < 1e-3, k
            continue
        g, r = p.grad.flatten().double(), g_ref.flatten().double()
        cos = float((g @ r) / (g.norm() * r.norm() + 1e-30))
        worst = min(worst, (cos, k))
    assert worst[0] >= 0.97, worst


def _close_grads(model, ref_grads, factor=1.0):
    for k, p in model.named_parameters():
        r = ref_grads[k]
        if r is None or (".key" in k and k.endswith(".bias")):
            continue
        g, r = p.grad.flatten().double(), r.flatten().double() * factor
        assert float((g @ r) / (g.norm() * r.norm() + 1e-30)) >= 0.97, k
        assert abs(float(g.norm()) - float(r.norm())) <= 0.15 * float(r.norm()) + 1e-6, k


def test_no_grad_frozen_layers_and_gradient_accumulation(simulated):
    """Host paths around the schedule: the no_grad forward equals the training-graph forward; frozen text layers receive no
    gradient while the rest is unchanged; a second backward without zeroing accumulates (2x the oracle gradient)."""
    from multimodal_classification_b200.vilbert import ViLBERTForClassification
    cfg = vo.tiny_config()
    sd = vo.seeded_state_dict(cfg)
    batch = vo.synthetic_batch(cfg, batch=2, seq=16, regions=8, seed=21)
    _, ref_grads = vo.loss_and_grads(sd, cfg, batch)
    model = ViLBERTForClassification(cfg, num_labels=2)
    model.load_state_dict(sd, strict=True)
    model.eval()
    with torch.no_grad():
        plain = model(**batch)
    out = model(**batch)
    assert torch.equal(plain["logits"], out["logits"]) and not plain["logits"].requires_grad and out["loss"].requires_grad
    out["loss"].backward()
    _close_grads(model, ref_grads)
    model(**batch)["loss"].backward()                    # no zero_grad in between: gradients add up
    _close_grads(model, ref_grads, factor=2.0)
    model.zero_grad(set_to_none=True)
    model.freeze_bert_layers(2)
    model(**batch)["loss"].backward()
    frozen = [k for k, p in model.named_parameters() if not p.requires_grad]
    assert frozen and all(k.startswith(("bert.embeddings.", "bert.encoder.layer.0.", "bert.encoder.layer.1.")) for k in frozen)
    assert all(dict(model.named_parameters())[k].grad is None for k in frozen)
    for p in model.parameters():
        p.requires_grad_(True)


@pytest.mark.parametrize("env", [{"VB_ONE_STREAM": "1"}, {"VB_BF16_RESIDUAL": "1"}, {"VB_NO_SIDE": "1"}])
def test_engine_switches_keep_the_result(simulated, monkeypatch, env):
    """The engine's scheduling switches (one stream, no side streams, bf16 residual stream) change where and in which
    precision kernels run, not what is computed."""
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    cfg = vo.tiny_config()
    model, out, ref_out, ref_grads = _run(cfg, dict(batch=2, seq=16, regions=8, seed=4))
    assert (out["logits"].float() - ref_out["logits"]).abs().max().item() <= 2e-2 * ref_out["logits"].abs().max().item()
    _close_grads(model, ref_grads)


def test_more_than_128_regions_and_mixed_input_dtypes(simulated):
    """130 regions (the log-sum-exp buffers grow to two query blocks), int32 ids, float masks, bf16 features."""
    from multimodal_classification_b200.vilbert import ViLBERTForClassification
    cfg = vo.tiny_config()
    sd = vo.seeded_state_dict(cfg)
    batch = vo.synthetic_batch(cfg, batch=2, seq=12, regions=130, seed=8, with_visual_mask=True)
    ref_out, _ = vo.loss_and_grads(sd, cfg, batch)
    model = ViLBERTForClassification(cfg, num_labels=2)
    model.load_state_dict(sd, strict=True)
    model.eval()
    mixed = dict(batch)
    mixed["input_ids"] = batch["input_ids"].int()
    mixed["attention_mask"] = batch["attention_mask"].float()
    mixed["visual_features"] = batch["visual_features"].to(torch.bfloat16)
    with torch.no_grad():
        out = model(**mixed)
    assert (out["logits"].float() - ref_out["logits"]).abs().max().item() <= 2e-2 * ref_out["logits"].abs().max().item()


def test_training_mode_code_paths_run(simulated, monkeypatch):
    """model.train(): every dropout branch of the forward / backward schedule executes (dropout itself acts as the identity in
    the stand-ins, so the result must still be the oracle's) — for both surfaces that share the engine."""
    import ops_sim
    from oracle import vilbert_core_oracle as co
    from test_vilbert_core_cpu import _model as core_model, _seeded_state
    monkeypatch.setattr(ops_sim, "IDENTITY_DROPOUT", True)
    cfg = vo.tiny_config()
    from multimodal_classification_b200.vilbert import ViLBERTForClassification
    sd = vo.seeded_state_dict(cfg)
    batch = vo.synthetic_batch(cfg, batch=2, seq=16, regions=8, seed=41)
    model = ViLBERTForClassification(cfg, num_labels=2)
    model.load_state_dict(sd, strict=True)
    model.train()
    out = model(**batch)
    out["loss"].backward()
    ref_out, ref_grads = vo.loss_and_grads(sd, cfg, batch)
    assert abs(out["loss"].item() - ref_out["loss"].item()) <= 1e-3
    _close_grads(model, ref_grads)
    ccfg = co.tiny_core_config()
    core = core_model(ccfg)
    csd = _seeded_state(core)
    core.load_state_dict(csd, strict=False)
    core.train()
    cbatch = co.synthetic_batch(ccfg, batch=2, seq=16, regions=10, seed=42)
    cout = core(**cbatch)
    cout["loss"].backward()
    want, _ = co.loss_and_grads(csd, ccfg, cbatch)
    # random-init logits of the 256-wide tiny core are ~3 units apart (loss 1.56): the bf16 operand roundings the stand-ins
    # reproduce move the loss by ~1.2e-3 relative; the fixture tests hold the 1e-3 bar on the pinned batch
    assert abs(cout["loss"].item() - want["loss"].item()) <= 3e-3


# ------------------------------------------------------------------------------------------------ data parallel (gloo)
def _dp_worker(rank, world, port, q):
    import os
    import sys
    import torch.distributed as dist
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import ops_sim
        from multimodal_classification_b200 import ddp
        from multimodal_classification_b200.vilbert import ViLBERTForClassification
        ops_sim.install(ops_sim.PlainPatch)
        ddp.FLUSH_BYTES = 1 << 20                       # several grouped exchanges inside one backward pass
        cfg = vo.tiny_config()
        sd = vo.seeded_state_dict(cfg)
        batch = vo.synthetic_batch(cfg, batch=2, seq=16, regions=8, seed=10 + rank)
        want = {}
        for k, g in vo.loss_and_grads(sd, cfg, batch)[1].items():          # mean over ranks of the single-rank oracle gradients
            if g is not None and not (".key" in k and k.endswith(".bias")):
                g = g.clone()
                dist.all_reduce(g)
                want[k] = (g / world).flatten().double()
        results = []
        for compress in (None, "bf16"):
            model = ViLBERTForClassification(cfg, num_labels=2)
            model.load_state_dict(sd, strict=True)
            model.eval()
            ddp.attach(model, dist.group.WORLD, compress=compress)
            model(**batch)["loss"].backward()
            worst, worst_rel = (1.0, ""), 0.0
            for k, p in model.named_parameters():
                if k not in want:
                    continue
                g = p.grad.flatten().double()
                worst = min(worst, (float((g @ want[k]) / (g.norm() * want[k].norm() + 1e-30)), k))
                worst_rel = max(worst_rel, float(abs(g.norm() - want[k].norm()) / (want[k].norm() + 1e-30)))
            digest = float(sum(p.grad.double().sum() for p in model.parameters() if p.grad is not None))
            results.append((str(compress), worst, worst_rel, digest))
        q.put((rank, results))
    finally:
        dist.destroy_process_group()


def test_data_parallel_backward_averages_gradients_over_gloo():
    """Two ranks, different batches, the engine's own bucket exchange inside backward (grouped all-reduces over the flat
    gradient buffer, fp32 and bf16-compressed): every rank ends with the mean of the two single-rank oracle gradients."""
    import socket
    import torch.multiprocessing as mp
    if torch.cuda.is_available():
        pytest.skip("the stand-ins are for the GPU-less container")
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_dp_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=300) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank in (0, 1):
        for compress, worst, worst_rel, digest in res[rank]:
            assert worst[0] >= 0.97, (rank, compress, worst)
            assert worst_rel <= 0.2, (rank, compress, worst_rel)
    for a, b in zip(res[0], res[1]):                                    # both ranks hold the same averaged gradients
        assert abs(a[3] - b[3]) <= 1e-6 * max(1.0, abs(a[3])), (a, b)


def test_shadow_refresh_inside_the_forward_orders_blocks_as_the_forward_touches_them(simulated, monkeypatch):
    """The bf16 shadow refresh as a branch of the forward (vilbert._Engine._begin_refresh / _shadow): every GEMM weight belongs
    to exactly one block, blocks are cast in the order the forward first touches them, every stream waits for a block's event
    before its first GEMM on that block, and after the forward every shadow is the bf16 rounding of its (updated) master."""
    import ops_sim
    from multimodal_classification_b200 import ops
    from multimodal_classification_b200.vilbert import ViLBERTForClassification
    cfg = vo.tiny_config()
    sd = vo.seeded_state_dict(cfg)
    batch = vo.synthetic_batch(cfg, batch=2, seq=16, regions=8, seed=4)
    model = ViLBERTForClassification(cfg, num_labels=2)
    model.load_state_dict(sd, strict=True)
    model.eval()
    with torch.no_grad():
        ref0 = model(**batch)["logits"].clone()
    eng = model._engine
    flat = eng.flat
    assert set(flat.block_of) == set(flat.order_w) and set(flat.block_of.values()) <= set(flat.buckets)
    order = eng._refresh_order()
    assert sorted(order) == sorted(flat.buckets) and order[0] == "tail" and order[1] == "t0"
    # run the refreshing forward over the stand-ins with a recording stream / event pair
    log = []

    class Ev(ops_sim._Event):
        def record(self, stream=None):
            log.append(("cast_done", id(self)))

    class St(ops_sim._Stream):
        def wait_event(self, event):
            log.append(("wait", id(event)))
    main = St()
    monkeypatch.setattr(torch.cuda, "Event", Ev)
    monkeypatch.setattr(torch.cuda, "current_stream", lambda d=None: main)
    casts = []
    real_cast = ops.cast_bf16
    monkeypatch.setattr(ops, "cast_bf16", lambda src, dst: (casts.append(src.data_ptr()), real_cast(src, dst))[1])
    eng.refresh_stream, eng.refresh_in_graph = ops_sim._Stream(), True
    with torch.no_grad():
        for p in model.parameters():
            p.mul_(1.01)                                   # an optimizer step: versions move
        out = model(**batch)["logits"]
    assert not torch.equal(out, ref0)
    # casts were issued block by block in execution order ...
    starts = [flat.master[flat.buckets[n][0]:].data_ptr() for n in order if min(flat.buckets[n][1], flat.w_end) > flat.buckets[n][0]]
    assert [c for c in casts if c in starts] == starts
    # ... every wait names an event that had been recorded before it (no use before the cast) ...
    done = set()
    for kind, ident in log:
        if kind == "cast_done":
            done.add(ident)
        else:
            assert ident in done
    assert sum(1 for k, _ in log if k == "wait") >= len(order)
    # ... and every shadow is fresh
    for key in flat.order_w:
        assert torch.equal(flat.w(key), flat.named[key].detach().to(torch.bfloat16)), key
    ref = vo.forward({k: v.detach() for k, v in model.state_dict().items()}, cfg, **batch)["logits"]
    assert (out.float() - ref).abs().max().item() <= 2e-2 * ref.abs().max().item()
