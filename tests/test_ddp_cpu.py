"""Data-parallel host logic on the CPU (gloo, world_size 2): gradient-bucket layout and the bucket all-reduce that the
backward pass issues per finished block (multimodal_classification_b200/ddp.py).  The kernels themselves need a GPU; the
N>1 path that does not (bucket ranges, mean all-reduce over contiguous slices of the flat gradient buffer, parameter
broadcast) is exercised here."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _tiny_model():
    from multimodal_classification_b200.vilbert import ViLBERTForClassification
    from oracle import vilbert_oracle as vo
    torch.manual_seed(0)
    return ViLBERTForClassification(vo.tiny_config(), num_labels=2)


def test_bucket_ranges_partition_the_gradient_buffer():
    from multimodal_classification_b200.vilbert import _FlatParams
    m = _tiny_model()
    flat = _FlatParams(m, torch.device("cpu"))
    spans = sorted(flat.buckets.values())
    assert spans[0][0] == 0 and spans[-1][1] == flat.s_end
    for (lo, hi), (lo2, _) in zip(spans, spans[1:]):
        assert lo < hi == lo2                      # contiguous, disjoint, nothing skipped
    # every used parameter's gradient lives in exactly one bucket; the unused q_dense* tensors in none
    for k, p in m.named_parameters():
        o = flat.offsets[k]
        inside = [n for n, (lo, hi) in flat.buckets.items() if lo <= o and o + p.numel() <= hi]
        assert len(inside) == (0 if "q_dense" in k else 1), (k, inside)
    # parameters are views of the flat master buffer (state_dict layout unchanged)
    assert all(p.data_ptr() >= flat.master.data_ptr() for p in m.parameters())
    cfg = m.config
    expected = {f"t{i}" for i in range(cfg["num_hidden_layers"])} | {f"v{i}" for i in range(cfg["v_num_hidden_layers"])} | \
               {f"c{i}" for i in range(cfg["num_co_attention_layers"])} | {"tail"}
    assert set(flat.buckets) == expected


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from multimodal_classification_b200 import ddp
        from multimodal_classification_b200.vilbert import _FlatParams
        torch.manual_seed(100 + rank)                 # different init per rank on purpose
        from multimodal_classification_b200.vilbert import ViLBERTForClassification
        from oracle import vilbert_oracle as vo
        m = ViLBERTForClassification(vo.tiny_config(), num_labels=2)
        ddp.broadcast_parameters(m)                   # now identical to rank 0
        flat = _FlatParams(m, torch.device("cpu"))
        digest = flat.master.double().sum().item()
        # per-rank "gradients", reduced bucket by bucket in the order the backward pass finishes them
        g = torch.Generator().manual_seed(7 + rank)
        flat.grad.copy_(torch.randn(flat.s_end, generator=g))
        mine = flat.grad.clone()
        order = ["tail"] + [n for n in flat.buckets if n != "tail"]
        pending = []
        for i, name in enumerate(reversed(order)):        # the engine queues finished buckets and exchanges them in groups
            pending.append(flat.buckets[name])
            if len(pending) == 3 or i == len(order) - 1:
                ranges = ddp.merge_ranges(pending)
                assert sum(hi - lo for lo, hi in ranges) == sum(hi - lo for lo, hi in pending)
                ddp.all_reduce_mean_ranges(flat.grad, ranges, dist.group.WORLD)
                pending = []
        others = [torch.randn(flat.s_end, generator=torch.Generator().manual_seed(7 + r)) for r in range(world)]
        want = torch.stack(others).mean(0)
        ok = torch.allclose(flat.grad, want, atol=1e-6) and torch.equal(others[rank], mine)
        ddp.attach(m, dist.group.WORLD)
        q.put((rank, digest, bool(ok), m._ddp_group is not None and m._engine is None and m._ddp_compress is None))   # fp32 exchange unless asked
    finally:
        dist.destroy_process_group()


def test_bucket_all_reduce_world_size_2():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=240) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res[0][1] == res[1][1]                     # broadcast made the replicas identical
    assert all(r[2] and r[3] for r in res)
