"""Shared helpers of the ingest tests: the committed fixture tests/golden/ingest.npz (written by
oracle/make_golden_ingest.py from the reference's own Dataset + DataLoader) unpacked into stores, frames and batches."""
import os

import numpy as np
import pandas as pd
import torch

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "ingest.npz"))
R, F, T, BS = int(G["R"]), int(G["F"]), int(G["T"]), int(G["BS"])
KEYS = ["input_ids", "attention_mask", "token_type_ids", "visual_features", "spatial_locations", "labels"]


def golden_batches(tag):
    return [{k: G[f"{tag}_{i}_{k}"] for k in KEYS} for i in range(int(G[tag + "_n"]))]


def store():
    blob, sizes, out, off = G["store_blob"].tobytes(), G["store_sizes"], {}, 0
    for key, n in zip(G["store_keys"], sizes):
        out[str(key).encode()] = blob[off:off + int(n)]
        off += int(n)
    return out


def frame():
    return pd.DataFrame({"id": [int(i) for i in G["ids"]], "text": [str(t) for t in G["texts"]],
                         "label": [int(v) for v in G["labels"]]})


def tokenizer():
    from transformers import BertTokenizer
    return BertTokenizer(vocab={str(w): i for i, w in enumerate(G["vocab"])})


def assert_batch_equal(got, want):
    assert list(got.keys()) == KEYS
    for k in KEYS:
        g = got[k].numpy() if isinstance(got[k], torch.Tensor) else got[k]
        assert g.dtype == want[k].dtype and g.shape == want[k].shape, k
        assert np.array_equal(g, want[k]), k
