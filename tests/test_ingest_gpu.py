"""Feature-store ingest on the GPU, through the C ABI (vb_lmdb_regions) and the loader built on it:
bit-exact against the oracle and against the batches the reference's own Dataset + DataLoader produced
(tests/golden/ingest.npz); the encoder gives the same logits from a loader batch as from the reference-format batch."""
import numpy as np
import pytest
import torch

from oracle import ingest_oracle as io
from oracle import vilbert_oracle as vo
from ingest_fixture import BS, F, G, KEYS, R, T, frame, golden_batches, store, tokenizer

pytestmark = pytest.mark.gpu


def bf16_rne(x: np.ndarray) -> torch.Tensor:
    return torch.from_numpy(x).to(torch.bfloat16)


@pytest.mark.parametrize("rows,feat,stride", [(1, 8, 4), (600, 16, 4), (1600, 2048, 4), (37, 24, 6), (51200, 2048, 5)])
def test_lmdb_regions_kernel_bit_exact(rows, feat, stride):
    from multimodal_classification_b200 import ops
    rng = np.random.default_rng(rows + feat)
    f = np.abs(rng.standard_normal((rows, feat))).astype(np.float32)
    f.reshape(-1)[:8] = [0.0, 1.00390625, 1.01171875, 3.4e38, 1e-40, -1.00390625, 65504.0, 2.0 ** -133]   # ties, inf, denormals
    b = rng.uniform(-50, 1100, (rows, stride)).astype(np.float32)
    b[0, :4] = [999.9, 0.1, 1000.1, 7.0]
    fd, bd = torch.from_numpy(f).cuda(), torch.from_numpy(b).cuda()
    f16 = torch.empty(rows, feat, dtype=torch.bfloat16, device="cuda")
    sp = torch.empty(rows, 5, dtype=torch.float32, device="cuda")
    ops.lmdb_regions(fd, f16, bd, sp)
    assert torch.equal(f16.cpu().view(torch.int16), bf16_rne(f).view(torch.int16))
    assert np.array_equal(sp.cpu().numpy().view(np.uint32), io.process_boxes(b, rows).view(np.uint32))
    # either half alone
    f16.zero_(); sp.zero_()
    ops.lmdb_regions(fd, f16)
    ops.lmdb_regions(boxes=bd, spatial=sp)
    assert torch.equal(f16.cpu().view(torch.int16), bf16_rne(f).view(torch.int16))
    assert np.array_equal(sp.cpu().numpy(), io.process_boxes(b, rows))
    # idempotence: values already representable in bf16 pass through unchanged
    again = torch.empty_like(f16)
    ops.lmdb_regions(f16.float(), again)
    assert torch.equal(again.view(torch.int16), f16.view(torch.int16))


def check(batch, want, feature_dtype):
    assert list(batch.keys()) == KEYS
    for k in KEYS:
        got = batch[k]
        assert got.is_cuda and tuple(got.shape) == want[k].shape, k
        if k == "visual_features" and feature_dtype == torch.bfloat16:
            assert got.dtype == torch.bfloat16 and torch.equal(got.cpu().view(torch.int16), bf16_rne(want[k]).view(torch.int16))
        else:
            assert got.cpu().numpy().dtype == want[k].dtype, k
            assert np.array_equal(got.cpu().numpy(), want[k]), k


@pytest.mark.parametrize("feature_dtype", [torch.bfloat16, torch.float32])
@pytest.mark.parametrize("depth", [2, 3, 5])
def test_loader_yields_the_reference_batches(feature_dtype, depth):
    from multimodal_classification_b200 import ingest
    rec = ingest.LMDBRecords(store().get, R, F)
    seq = ingest.FeatureStoreLoader(frame(), rec, tokenizer(), T, BS, depth=depth, feature_dtype=feature_dtype)
    want = golden_batches("lmdb_seq")
    assert len(seq) == len(want) == 3 and len(seq.dataset) == len(G["ids"])
    for _ in range(2):                                  # two epochs over the same ring
        n = 0
        for batch, w in zip(seq, want):
            check(batch, w, feature_dtype)
            n += 1
        assert n == 3
    shuf = ingest.FeatureStoreLoader(frame(), rec, tokenizer(), T, BS, shuffle=True, drop_last=True, depth=depth,
                                     feature_dtype=feature_dtype)
    torch.manual_seed(2024)
    got = list(zip(shuf, golden_batches("lmdb_shuf")))
    assert len(got) == len(shuf) == 2
    # only the newest batch of a ring is guaranteed live: re-run and check as we go
    torch.manual_seed(2024)
    for batch, w in zip(shuf, golden_batches("lmdb_shuf")):
        check(batch, w, feature_dtype)


def test_loader_over_hdf5_layout_arrays():
    from multimodal_classification_b200 import ingest
    id_map = {str(k): int(v) for k, v in zip(G["h5_ids"], G["h5_rows"])}
    rec = ingest.ArrayRecords(G["h5_visual"], G["h5_spatial"], id_map, R, F)
    for dt in (torch.bfloat16, torch.float32):
        loader = ingest.FeatureStoreLoader(frame(), rec, tokenizer(), T, BS, feature_dtype=dt)
        for batch, w in zip(loader, golden_batches("h5_seq")):
            check(batch, w, dt)


def test_producer_errors_surface_and_early_exit_is_clean():
    import pickle
    from multimodal_classification_b200 import ingest
    from multimodal_classification_b200._lib import VbError
    st = store()
    st[b"1006"] = pickle.dumps({"features": np.zeros((R + 2, F), np.float32)})
    bad = ingest.FeatureStoreLoader(frame(), ingest.LMDBRecords(st.get, R, F), tokenizer(), T, BS)
    with pytest.raises(VbError, match="features of shape"):
        list(bad)
    good = ingest.FeatureStoreLoader(frame(), ingest.LMDBRecords(store().get, R, F), tokenizer(), T, 2)
    for i, _ in enumerate(good):
        if i == 1:
            break                                        # abandon the epoch with batches still in flight
    for batch, w in zip(ingest.FeatureStoreLoader(frame(), ingest.LMDBRecords(store().get, R, F), tokenizer(), T, BS),
                        golden_batches("lmdb_seq")):
        check(batch, w, torch.bfloat16)
    assert len(list(good)) == len(good) == 6


def test_encoder_consumes_loader_batches_like_reference_batches():
    """Same logits and loss whether the encoder is fed the loader's HBM-resident batch or the reference-format batch
    (fp32 features and oracle-normalised boxes moved to the device by the caller, nodes.py:784)."""
    from multimodal_classification_b200 import ingest
    from multimodal_classification_b200.vilbert import ViLBERTForClassification
    rows, st = io.seeded_store(20, 2048, seed=3)
    import pandas as pd
    df = pd.DataFrame({"id": [int(r[0]) for r in rows], "text": [r[1] for r in rows], "label": [r[2] for r in rows]})
    tok = tokenizer()
    torch.manual_seed(0)
    model = ViLBERTForClassification(vo.tiny_config(), num_labels=2).cuda().eval()
    loader = ingest.FeatureStoreLoader(df, ingest.LMDBRecords(st.get, 20, 2048), tok, 32, 4)
    start = 0
    for batch in loader:
        with torch.no_grad():
            out = model(**batch)
        n = batch["labels"].shape[0]
        samples = [io.lmdb_sample(str(df.iloc[i]["id"]), str(df.iloc[i]["text"]), int(df.iloc[i]["label"]), st.get, tok, 32, 20,
                                  2048) for i in range(start, start + n)]
        ref_batch = {k: torch.from_numpy(v).cuda() for k, v in io.collate(samples).items()}
        with torch.no_grad():
            want = model(**ref_batch)
        assert torch.equal(out["logits"], want["logits"]) and torch.equal(out["loss"], want["loss"])
        start += n
    assert start == len(df)
