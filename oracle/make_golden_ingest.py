"""Generate tests/golden/ingest.npz by running the reference's OWN dataset classes
(pipelines/data_processing/lmdb_dataset.py, precomputed_dataset.py) and ``torch.utils.data.DataLoader`` exactly as
``create_lmdb_dataloaders`` / ``create_precomputed_dataloaders`` configure it.

``lmdb`` and ``h5py`` are not installed in the authoring container and the real feature stores cannot be downloaded, so both
modules are replaced by in-memory stand-ins that offer only what the reference calls (``lmdb.open(...).begin().get / stat``,
``h5py.File(path)[name][row]``); the two reference files are loaded by path (their package ``__init__`` pulls in kedro) and
run unmodified.  ``FIXED_NUM_REGIONS / FIXED_FEATURE_DIM`` are overridden in a subclass so that the committed fixture stays
small; they only size the zero tensors of the fall-back branches."""
import importlib.util
import os
import sys
import types

import numpy as np
import pandas as pd
import torch
from torch.utils.data import DataLoader

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ingest_oracle as io  # noqa: E402

REF = "/root/reference/src/multimodalclassification/pipelines/data_processing"
R, F, T, BS = 6, 16, 24, 4


class _Txn:
    def __init__(self, store):
        self.store = store

    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False

    def get(self, key):
        return self.store.get(bytes(key))

    def stat(self):
        return {"entries": len(self.store)}


class _Env:
    def __init__(self, store):
        self.store = store

    def begin(self, write=False):
        return _Txn(self.store)

    def close(self):
        pass


def _load(name):
    spec = importlib.util.spec_from_file_location("ref_" + name, os.path.join(REF, name + ".py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def _batches(loader):
    out = []
    for b in loader:
        out.append({k: v.numpy() for k, v in b.items()})
    return out


def main():
    from transformers import BertTokenizer
    rows, store = io.seeded_store(R, F)
    stores = {"mem://detectron.lmdb": store}
    lmdb = types.ModuleType("lmdb")
    lmdb.open = lambda path, **kw: _Env(stores[path])
    h5py = types.ModuleType("h5py")
    h5 = {}
    h5py.File = lambda path, mode="r": h5[path]
    sys.modules["lmdb"], sys.modules["h5py"] = lmdb, h5py
    ref_lmdb, ref_pre = _load("lmdb_dataset"), _load("precomputed_dataset")

    tok = BertTokenizer(vocab={w: i for i, w in enumerate(io.VOCAB)})
    df = pd.DataFrame({"id": [int(r[0]) for r in rows], "text": [r[1] for r in rows], "label": [r[2] for r in rows]})

    class Small(ref_lmdb.LMDBFeaturesDataset):
        FIXED_NUM_REGIONS, FIXED_FEATURE_DIM = R, F

    ds = Small(df, "mem://detectron.lmdb", tok, max_seq_length=T)
    out = {"R": R, "F": F, "T": T, "BS": BS, "ids": np.array([r[0] for r in rows]), "texts": np.array([r[1] for r in rows]),
           "labels": np.array([r[2] for r in rows]), "vocab": np.array(io.VOCAB),
           "store_keys": np.array([k.decode() for k in store]),
           "store_blob": np.frombuffer(b"".join(store.values()), np.uint8),
           "store_sizes": np.array([len(v) for v in store.values()])}

    def put(tag, batches):
        out[tag + "_n"] = len(batches)
        for i, b in enumerate(batches):
            for k, v in b.items():
                out[f"{tag}_{i}_{k}"] = v

    # evaluation-style loader (lmdb_dataset.py:298-304) and training-style loader (:289-296) under a known global seed
    put("lmdb_seq", _batches(DataLoader(ds, batch_size=BS, shuffle=False, num_workers=0)))
    torch.manual_seed(2024)
    put("lmdb_shuf", _batches(DataLoader(ds, batch_size=BS, shuffle=True, num_workers=0, drop_last=True)))

    # HDF5-layout store (precomputed_dataset.py): two row-indexed arrays and an id -> row map; one id is left out
    rng = np.random.default_rng(11)
    known = [r[0] for r in rows if r[0] != "1007"]
    id_map = {k: i for i, k in enumerate(reversed(known))}
    vis = np.abs(rng.standard_normal((len(known), R, F))).astype(np.float32)
    spa = rng.uniform(0, 1, (len(known), R, 5)).astype(np.float32)
    h5["mem://features.h5"] = {"visual_features": vis, "spatial_features": spa}
    map_path = "/tmp/_ingest_id_map.npy"
    np.save(map_path, id_map, allow_pickle=True)
    ds2 = ref_pre.PrecomputedFeaturesDataset(df, "mem://features.h5", map_path, tok, max_seq_length=T, num_regions=R,
                                             visual_feature_dim=F)
    put("h5_seq", _batches(DataLoader(ds2, batch_size=BS, shuffle=False, num_workers=0)))
    out.update(h5_visual=vis, h5_spatial=spa, h5_ids=np.array(list(id_map.keys())), h5_rows=np.array(list(id_map.values())))

    path = os.path.join(ROOT, "tests", "golden", "ingest.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path) // 1024, "KiB;", out["lmdb_seq_n"], "+", out["lmdb_shuf_n"], "+", out["h5_seq_n"],
          "batches")


if __name__ == "__main__":
    main()
