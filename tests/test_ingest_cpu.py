"""Feature-store ingest on the CPU: the oracle restatement and the product's host-side logic (record
decoding, bulk tokenisation, blob layout, epoch order) against tests/golden/ingest.npz, the batches the reference's own
LMDBFeaturesDataset / PrecomputedFeaturesDataset + DataLoader produced (oracle/make_golden_ingest.py).  No kernel runs here:
the device half (bf16 rounding, box normalisation) is covered by tests/test_ingest_gpu.py."""

import numpy as np
import pytest
import torch

from oracle import ingest_oracle as io

from ingest_fixture import BS, F, G, R, T, assert_batch_equal, frame, golden_batches, store, tokenizer  # noqa: E402


def test_fixture_store_is_the_seeded_store():
    rows, st = io.seeded_store(R, F)
    assert st == store() and [r[0] for r in rows] == [str(i) for i in G["ids"]]


def test_oracle_reproduces_reference_batches_bit_exact():
    st, tok, df = store(), tokenizer(), frame()
    samples = [io.lmdb_sample(str(df.iloc[i]["id"]), str(df.iloc[i]["text"]), int(df.iloc[i]["label"]), st.get, tok, T, R, F)
               for i in range(len(df))]
    want = golden_batches("lmdb_seq")
    plan = io.batch_indices(len(df), BS, drop_last=False)
    assert len(plan) == len(want) == 3 and len(plan[-1]) == 3
    for idx, w in zip(plan, want):
        assert_batch_equal(io.collate([samples[i] for i in idx]), w)


def test_oracle_box_arithmetic_edge_rows():
    b = np.array([[999.9, 0.1, 1000.1, 7.0], [500, 400, 300, 200]], np.float32)
    s = io.process_boxes(b, 2)
    assert s.dtype == np.float32 and s[1, 4] == np.float32(0.04) and s[0, 0] == np.float32(999.9) / np.float32(1000.0)
    assert not io.process_boxes(None, 3).any() and not io.process_boxes(b[0], 3).any() and not io.process_boxes(b[:, :3], 3).any()


def test_product_host_packing_matches_reference_batches():
    """TextTable + LMDBRecords + pack_batch fill the blob with exactly the reference's ids/mask/types/labels/features; the raw
    boxes they leave for the device normalise (oracle arithmetic) to the reference's spatial rows."""
    from multimodal_classification_b200 import ingest
    table = ingest.TextTable(frame(), tokenizer(), T)
    rec = ingest.LMDBRecords(store().get, R, F)
    want = golden_batches("lmdb_seq")
    for idx, w in zip(ingest.batch_plan(ingest.epoch_order(len(table), False), BS, False), want):
        lay = ingest.BatchLayout(len(idx), T, R, F, rec.box_width)
        blob = torch.zeros(lay.nbytes, dtype=torch.uint8)
        views = lay.views(blob)
        ingest.pack_batch(table, rec, idx, {k: v.numpy() for k, v in views.items()})
        spatial = np.stack([io.process_boxes(b, R) for b in views["boxes"].numpy()])
        assert_batch_equal({"input_ids": views["input_ids"], "attention_mask": views["attention_mask"],
                            "token_type_ids": views["token_type_ids"], "visual_features": views["features"],
                            "spatial_locations": spatial, "labels": views["labels"]}, w)
        assert all(o % 256 == 0 for o in lay.offsets.values())


def test_product_array_records_match_reference_hdf5_batches():
    from multimodal_classification_b200 import ingest
    table = ingest.TextTable(frame(), tokenizer(), T)
    id_map = {str(k): int(v) for k, v in zip(G["h5_ids"], G["h5_rows"])}
    rec = ingest.ArrayRecords(G["h5_visual"], G["h5_spatial"], id_map, R, F)
    for idx, w in zip(ingest.batch_plan(range(len(table)), BS, False), golden_batches("h5_seq")):
        lay = ingest.BatchLayout(len(idx), T, R, F, rec.box_width)
        views = lay.views(torch.zeros(lay.nbytes, dtype=torch.uint8))
        ingest.pack_batch(table, rec, idx, {k: v.numpy() for k, v in views.items()})
        assert_batch_equal({"input_ids": views["input_ids"], "attention_mask": views["attention_mask"],
                            "token_type_ids": views["token_type_ids"], "visual_features": views["features"],
                            "spatial_locations": views["boxes"], "labels": views["labels"]}, w)


def test_shuffled_epoch_order_is_the_dataloaders():
    """Under the same global seed the product's epoch order equals the reference loader's (the shuffled fixture batches), and
    equals torch's DataLoader on an index dataset for two consecutive epochs."""
    from multimodal_classification_b200 import ingest
    torch.manual_seed(2024)
    plan = ingest.batch_plan(ingest.epoch_order(len(G["ids"]), True), BS, True)
    want = golden_batches("lmdb_shuf")
    labels = np.array([int(v) for v in G["labels"]])
    seq_ids = np.concatenate([b["input_ids"] for b in golden_batches("lmdb_seq")])
    assert len(plan) == len(want) == 2
    for idx, w in zip(plan, want):
        assert np.array_equal(seq_ids[idx], w["input_ids"]) and np.array_equal(labels[idx], w["labels"])
    from torch.utils.data import DataLoader
    torch.manual_seed(5)
    dl = DataLoader(list(range(37)), batch_size=5, shuffle=True, drop_last=True)
    ref = [[b.tolist() for b in dl] for _ in range(2)]
    torch.manual_seed(5)
    got = [ingest.batch_plan(ingest.epoch_order(37, True), 5, True) for _ in range(2)]
    assert got == ref


def test_ragged_record_is_an_error_and_loader_needs_cuda():
    import pickle
    from multimodal_classification_b200 import ingest
    from multimodal_classification_b200._lib import VbError
    rec = ingest.LMDBRecords({b"7": pickle.dumps({"features": np.zeros((R + 1, F), np.float32)})}.get, R, F)
    with pytest.raises(VbError, match="features of shape"):
        rec.fetch("7", np.zeros((R, F), np.float32), np.zeros((R, 4), np.float32))
    if not torch.cuda.is_available():
        with pytest.raises(VbError, match="no CPU fall-back"):
            ingest.FeatureStoreLoader(frame(), rec, tokenizer(), T)


def test_shard_order_is_distributed_samplers():
    """Per-rank order = torch's DistributedSampler (same seed / epoch), for even and ragged splits, shuffled or not; the
    ranks' shards have equal length and together cover every sample."""
    from torch.utils.data.distributed import DistributedSampler
    from multimodal_classification_b200 import ingest
    for n, world in [(37, 2), (40, 4), (5, 8), (1, 2), (64, 8)]:
        data = list(range(n))
        for shuffle in (False, True):
            for epoch in (0, 3):
                shards = []
                for rank in range(world):
                    ref = DistributedSampler(data, num_replicas=world, rank=rank, shuffle=shuffle, seed=11)
                    ref.set_epoch(epoch)
                    got = ingest.shard_order(n, shuffle, rank, world, seed=11, epoch=epoch)
                    assert got == list(ref), (n, world, rank, shuffle, epoch)
                    shards.append(got)
                assert len({len(s) for s in shards}) == 1 and set().union(*shards) == set(data)
    assert ingest.shard_order(0, True, 1, 2) == []
