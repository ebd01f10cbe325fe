"""ORACLE (test infrastructure, never shipped, never on the product path).

CPU restatement of the reference's feature-store ingest: what one sample of
``LMDBFeaturesDataset`` (/root/reference/src/multimodalclassification/pipelines/data_processing/lmdb_dataset.py:126-239)
and of ``PrecomputedFeaturesDataset`` (.../precomputed_dataset.py:78-123) is, and how ``torch.utils.data.DataLoader``
(default collate, ``RandomSampler`` when ``shuffle=True``; lmdb_dataset.py:289-312) turns samples into batches.  Plain numpy,
float32 arithmetic in the reference's operation order.  Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU
legs may import it.

Third-party pieces on this path (absent from /root/reference, unpinned by it): ``lmdb`` (a byte-key -> byte-value store: the
reference only calls ``txn.get(key)``), ``h5py`` (the reference only indexes two datasets by row), ``pickle`` and the
``transformers`` tokenizer (called once per sample with ``max_length / padding="max_length" / truncation=True``).

Pinning: the reference holds no test for this path, so ``oracle/make_golden_ingest.py`` runs the reference's OWN dataset
classes in the authoring container over an in-memory stand-in for ``lmdb`` / ``h5py`` (neither is installed here) and commits
the stored records together with the batches the reference produced as ``tests/golden/ingest.npz``;
``tests/test_ingest_cpu.py`` checks this file against them bit for bit (every key, dtype and value).
"""
from __future__ import annotations

import pickle
from typing import Callable, Dict, List, Optional, Sequence, Tuple

import numpy as np

BOX_DIV = 1000.0          # lmdb_dataset.py:193-203 "assumed 1000x1000 image size"
AREA_DIV = 1000000.0      # lmdb_dataset.py:196


def query_keys(img_id: str) -> List[bytes]:
    """Keys tried in order for one image id (lmdb_dataset.py:130-137)."""
    return [img_id.encode(), img_id.encode(), f"{img_id}.png".encode(), img_id.zfill(5).encode()]


def query(get: Callable[[bytes], Optional[bytes]], img_id: str):
    """lmdb_dataset.py:126-141: the first key that is present wins; the value is a pickle."""
    for key in query_keys(img_id):
        item = get(key)
        if item is not None:
            return pickle.loads(item)
    return None


def process_boxes(boxes, num_regions: int) -> np.ndarray:
    """lmdb_dataset.py:181-208: [x1, y1, x2, y2] -> [x1/1000, y1/1000, x2/1000, y2/1000, (w*h)/1e6], all float32."""
    if boxes is None:
        return np.zeros((num_regions, 5), np.float32)
    boxes = np.array(boxes, dtype=np.float32)
    if boxes.ndim != 2 or boxes.shape[1] < 4:
        return np.zeros((num_regions, 5), np.float32)
    w = boxes[:, 2] - boxes[:, 0]
    h = boxes[:, 3] - boxes[:, 1]
    area = (w * h) / np.float32(AREA_DIV)
    d = np.float32(BOX_DIV)
    return np.stack([boxes[:, 0] / d, boxes[:, 1] / d, boxes[:, 2] / d, boxes[:, 3] / d, area], axis=1).astype(np.float32)


def extract_features(record, num_regions: int, feature_dim: int) -> Tuple[np.ndarray, np.ndarray]:
    """lmdb_dataset.py:143-179: key fall-backs ``features / feature / fc6`` and ``boxes / bbox``; a record that is not a dict
    is the feature array itself; anything missing becomes zeros."""
    if record is None:
        return np.zeros((num_regions, feature_dim), np.float32), np.zeros((num_regions, 5), np.float32)
    if isinstance(record, dict):
        features = record.get("features")
        if features is None:
            features = record.get("feature")
        if features is None:
            features = record.get("fc6")
        boxes = record.get("boxes")
        if boxes is None:
            boxes = record.get("bbox")
    else:
        features, boxes = record, None
    if features is not None:
        visual = np.array(features, dtype=np.float32)
    else:
        visual = np.zeros((num_regions, feature_dim), np.float32)
    return visual, process_boxes(boxes, num_regions)


def precomputed_features(visual_store, spatial_store, id_map: Dict[str, int], img_id: str, num_regions: int,
                         feature_dim: int) -> Tuple[np.ndarray, np.ndarray]:
    """precomputed_dataset.py:84-99: row ``id_map[img_id]`` of the two stored arrays, zeros when the id is unknown."""
    if img_id in id_map:
        i = id_map[img_id]
        return np.asarray(visual_store[i], np.float32), np.asarray(spatial_store[i], np.float32)
    return np.zeros((num_regions, feature_dim), np.float32), np.zeros((num_regions, 5), np.float32)


def tokenize(tokenizer, text: str, max_seq_length: int) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    """lmdb_dataset.py:221-235: one padded, truncated encoding per sample; token types default to zeros."""
    enc = tokenizer(text, max_length=max_seq_length, padding="max_length", truncation=True, return_tensors="np")
    ids = np.asarray(enc["input_ids"]).reshape(-1).astype(np.int64)
    mask = np.asarray(enc["attention_mask"]).reshape(-1).astype(np.int64)
    types = np.asarray(enc["token_type_ids"]).reshape(-1).astype(np.int64) if "token_type_ids" in enc else np.zeros_like(ids)
    return ids, mask, types


def collate(samples: Sequence[Dict[str, np.ndarray]]) -> Dict[str, np.ndarray]:
    """torch default_collate on dicts of equal-shaped tensors: stack along a new leading axis, key order kept."""
    return {k: np.stack([s[k] for s in samples], axis=0) for k in samples[0]}


def lmdb_sample(row_id: str, text: str, label: int, get, tokenizer, max_seq_length: int, num_regions: int,
                feature_dim: int) -> Dict[str, np.ndarray]:
    """lmdb_dataset.py:210-239."""
    visual, spatial = extract_features(query(get, row_id), num_regions, feature_dim)
    ids, mask, types = tokenize(tokenizer, text, max_seq_length)
    return {"input_ids": ids, "attention_mask": mask, "token_type_ids": types, "visual_features": visual,
            "spatial_locations": spatial, "labels": np.asarray(label, np.int64)}


def batch_indices(n: int, batch_size: int, drop_last: bool, order: Optional[Sequence[int]] = None) -> List[List[int]]:
    """BatchSampler over ``order`` (sequential when None)."""
    order = list(range(n)) if order is None else list(order)
    out = [order[i:i + batch_size] for i in range(0, n, batch_size)]
    if drop_last and out and len(out[-1]) < batch_size:
        out.pop()
    return out


# ------------------------------------------------------------------------------------------------ seeded stand-in store
def seeded_store(num_regions: int = 6, feature_dim: int = 16, seed: int = 7):
    """A small in-memory LMDB image: (rows, store).  ``rows`` = list of (id, text, label); ``store`` = {key bytes: pickle}.
    Covers every branch of lmdb_dataset.py:126-208: each key spelling, a missing id, the three feature key names, both box
    key names, a bare-array record, boxes that are absent / 1-D / too narrow / wider than four columns / float64 lists,
    coordinates whose quotients by 1000 are not exactly representable."""
    rng = np.random.default_rng(seed)

    def feats():
        return np.abs(rng.standard_normal((num_regions, feature_dim))).astype(np.float32)

    def boxes(cols=4, dtype=np.float32):
        x1 = rng.uniform(0, 700, num_regions)
        y1 = rng.uniform(0, 700, num_regions)
        b = np.stack([x1, y1, x1 + rng.uniform(5, 333.3, num_regions), y1 + rng.uniform(5, 333.3, num_regions)], 1)
        if cols > 4:
            b = np.concatenate([b, rng.uniform(0, 1, (num_regions, cols - 4))], 1)
        return b.astype(dtype)

    rows, store = [], {}

    def add(img_id, key, record, text, label):
        rows.append((img_id, text, label))
        if key is not None:
            store[key] = pickle.dumps(record, protocol=4)

    add("1001", b"1001", {"features": feats(), "boxes": boxes()}, "hello world", 1)
    add("1002", b"1002.png", {"feature": feats(), "bbox": boxes()}, "the cat sat on the mat", 0)
    add("37", b"00037", {"fc6": feats(), "boxes": boxes(6)}, "meme", 1)
    add("4242", None, None, "this id is not in the store", 0)
    add("1005", b"1005", feats(), "bare array record", 1)
    add("1006", b"1006", {"features": feats()}, "", 0)
    add("1007", b"1007", {"features": feats(), "boxes": boxes()[:, :3]}, "boxes too narrow", 1)
    add("1008", b"1008", {"features": feats(), "boxes": boxes()[0]}, "boxes one-dimensional", 0)
    add("1009", b"1009", {"features": feats().astype(np.float64).tolist(), "boxes": boxes(4, np.float64).tolist()},
        "lists of python floats " * 8, 1)
    add("1010", b"1010", {"boxes": boxes()}, "no features at all!", 0)
    hand = np.array([[0, 0, 1000, 1000], [1, 2, 3, 4], [999.9, 0.1, 1000.1, 7], [333, 333, 666, 667], [10, 10, 10, 10],
                     [500, 400, 300, 200]], np.float32)
    add("1011", b"1011", {"features": feats(), "boxes": np.resize(hand, (num_regions, 4))}, "hand-picked boxes", 1)
    return rows, store


VOCAB = ["[PAD]", "[UNK]", "[CLS]", "[SEP]", "[MASK]", "hello", "world", "the", "cat", "sat", "on", "mat", "me", "##me",
         "this", "id", "is", "not", "in", "store", "bare", "array", "record", "boxes", "too", "narrow", "one", "-",
         "dimensional", "lists", "of", "python", "float", "##s", "no", "features", "at", "all", "!", "hand", "picked"]
