"""The loader's producer/consumer ring (multimodal_classification_b200/ingest.py) exercised on the CPU: CUDA streams, events
and pinned memory are replaced by inert stand-ins and the one kernel by the oracle's arithmetic, so that slot hand-over,
epoch restarts, short last batches, abandoned epochs and producer errors are covered without a GPU.  The real device path
is tests/test_ingest_gpu.py."""
import pickle

import numpy as np
import pytest
import torch

from ingest_fixture import BS, F, G, KEYS, R, T, frame, golden_batches, store, tokenizer  # noqa: E402


@pytest.fixture
def ingest_on_cpu(monkeypatch):
    import cuda_standins
    from multimodal_classification_b200 import ingest
    if torch.cuda.is_available():
        pytest.skip("stand-ins are for the GPU-less container")
    cuda_standins.apply(monkeypatch.setattr)
    return ingest


def check(batch, want, dt):
    assert list(batch.keys()) == KEYS
    for k in KEYS:
        got = batch[k]
        assert tuple(got.shape) == want[k].shape, k
        if k == "visual_features" and dt == torch.bfloat16:
            assert torch.equal(got.view(torch.int16), torch.from_numpy(want[k]).to(torch.bfloat16).view(torch.int16))
        else:
            assert got.numpy().dtype == want[k].dtype and np.array_equal(got.numpy(), want[k]), k


@pytest.mark.parametrize("dt", [torch.bfloat16, torch.float32])
@pytest.mark.parametrize("depth", [2, 3, 5])
def test_ring_yields_reference_batches_over_epochs(ingest_on_cpu, dt, depth):
    ingest = ingest_on_cpu
    rec = ingest.LMDBRecords(store().get, R, F)
    seq = ingest.FeatureStoreLoader(frame(), rec, tokenizer(), T, BS, depth=depth, feature_dtype=dt, device="cpu")
    assert len(seq) == 3 and len(seq.dataset) == len(G["ids"])
    for _ in range(2):
        assert sum(check(b, w, dt) is None for b, w in zip(seq, golden_batches("lmdb_seq"))) == 3
    shuf = ingest.FeatureStoreLoader(frame(), rec, tokenizer(), T, BS, shuffle=True, drop_last=True, depth=depth,
                                     feature_dtype=dt, device="cpu")
    torch.manual_seed(2024)
    assert sum(check(b, w, dt) is None for b, w in zip(shuf, golden_batches("lmdb_shuf"))) == len(shuf) == 2
    id_map = {str(k): int(v) for k, v in zip(G["h5_ids"], G["h5_rows"])}
    h5 = ingest.FeatureStoreLoader(frame(), ingest.ArrayRecords(G["h5_visual"], G["h5_spatial"], id_map, R, F), tokenizer(), T,
                                   BS, depth=depth, feature_dtype=dt, device="cpu")
    assert sum(check(b, w, dt) is None for b, w in zip(h5, golden_batches("h5_seq"))) == 3


def test_ring_errors_early_exit_and_empty_split(ingest_on_cpu):
    ingest = ingest_on_cpu
    st = store()
    st[b"1006"] = pickle.dumps({"features": np.zeros((R + 2, F), np.float32)})
    with pytest.raises(ingest.VbError, match="features of shape"):
        list(ingest.FeatureStoreLoader(frame(), ingest.LMDBRecords(st.get, R, F), tokenizer(), T, BS, device="cpu"))
    good = ingest.FeatureStoreLoader(frame(), ingest.LMDBRecords(store().get, R, F), tokenizer(), T, 2, device="cpu")
    for i, _ in enumerate(good):
        if i == 1:
            break
    assert len(list(good)) == len(good) == 6
    empty = ingest.FeatureStoreLoader(frame().iloc[:0], ingest.LMDBRecords(store().get, R, F), tokenizer(), T, 2, device="cpu")
    assert len(empty) == 0 and list(empty) == []
    with pytest.raises(ingest.VbError):
        ingest.FeatureStoreLoader(frame(), ingest.LMDBRecords(store().get, R, F), tokenizer(), T, 2, device="cpu", depth=1)


def test_two_ranks_get_disjoint_equal_shards(ingest_on_cpu):
    """Two loaders as two data-parallel ranks would build them: same epoch permutation, disjoint halves, equal batch counts;
    set_epoch changes the permutation for both alike."""
    ingest = ingest_on_cpu
    rec = ingest.LMDBRecords(store().get, R, F)
    loaders = [ingest.FeatureStoreLoader(frame(), rec, tokenizer(), T, 2, shuffle=True, drop_last=False, device="cpu", rank=r,
                                         world_size=2, seed=5) for r in range(2)]
    assert len(loaders[0]) == len(loaders[1]) == 3          # 11 samples -> 6 per rank (one wrapped) -> 3 batches of 2
    seen = []
    for epoch in (0, 1):
        per_rank = []
        for ld in loaders:
            ld.set_epoch(epoch)
            per_rank.append(np.concatenate([b["input_ids"].numpy() for b in ld]))
        assert per_rank[0].shape == per_rank[1].shape == (6, T)
        seen.append(np.concatenate(per_rank))
    all_ids = np.concatenate([b["input_ids"] for b in golden_batches("lmdb_seq")])
    for rows in seen:                                        # 12 rows = the 11 distinct samples + one wrapped duplicate
        assert {r.tobytes() for r in rows} == {r.tobytes() for r in all_ids}
    assert not np.array_equal(seen[0], seen[1])
    with pytest.raises(ingest.VbError):
        ingest.FeatureStoreLoader(frame(), rec, tokenizer(), T, 2, device="cpu", rank=2, world_size=2)


# ------------------------------------------------------------------------------------------------ two processes (gloo)
def _rank_worker(rank, world, port, q):
    import os
    import sys
    import torch.distributed as dist
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import cuda_standins
        from multimodal_classification_b200 import ingest
        cuda_standins.apply(setattr)
        rec = ingest.LMDBRecords(store().get, R, F)
        train, val, _ = ingest.create_lmdb_dataloaders(frame(), frame().iloc[:5], frame().iloc[:2], batch_size=2, max_seq_length=T,
                                                        device="cpu", tokenizer=tokenizer(), records=rec)
        mine = np.concatenate([b["input_ids"].numpy() for b in train])
        val_rows = sum(b["labels"].shape[0] for b in val)
        gathered = [None] * world
        dist.all_gather_object(gathered, mine)
        q.put((rank, (train.rank, train.world_size, val.world_size), len(train), mine.shape, val_rows,
               [g.tobytes() for g in gathered]))
    finally:
        dist.destroy_process_group()


def test_factories_shard_the_training_split_over_a_gloo_group():
    """create_lmdb_dataloaders inside a world-size-2 process group: the training loaders take their rank from the group and
    read disjoint, equally long shards of one permutation; validation stays whole on every rank."""
    import socket
    import torch.multiprocessing as mp
    if torch.cuda.is_available():
        pytest.skip("stand-ins are for the GPU-less container")
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_rank_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=240) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert [r[1] for r in res] == [(0, 2, 1), (1, 2, 1)]
    assert res[0][2] == res[1][2] == 3 and res[0][3] == res[1][3] == (6, T)      # 11 samples -> 6 per rank, drop_last keeps 3 x 2
    assert res[0][4] == res[1][4] == 5
    assert res[0][5] == res[1][5]                                                 # both ranks gathered the same shards
    rows = np.concatenate([np.frombuffer(b, np.int64).reshape(-1, T) for b in res[0][5]])
    all_ids = np.concatenate([b["input_ids"] for b in golden_batches("lmdb_seq")])
    assert {r.tobytes() for r in rows} == {r.tobytes() for r in all_ids}          # together: every sample (one wrapped twice)
