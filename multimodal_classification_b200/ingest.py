"""Feature-store ingest for the ViLBERT hot path.

Replaces the reference's per-sample ``Dataset`` + ``DataLoader(num_workers=0)`` in front of the encoder
(/root/reference/src/multimodalclassification/pipelines/data_processing/lmdb_dataset.py:61-319 for Facebook's
``detectron.lmdb``, precomputed_dataset.py:21-228 for the HDF5 layout) by a loader that yields the SAME batch dicts — same
keys, shapes, index dtypes and values — already resident in HBM:

* the text columns are tokenised once, in bulk, when the loader is built (the reference tokenises every sample again in every
  epoch, lmdb_dataset.py:221-228);
* a producer thread decodes the records of the next batches straight into one pinned host blob per batch (features, raw
  boxes, ids, mask, token types, labels), ships the blob with ONE host->device copy on its own stream and runs ONE kernel
  (``vb_lmdb_regions``) that rounds the features to bf16 and normalises the boxes exactly as ``_process_boxes`` (:181-208);
* the consumer only waits on a CUDA event: copy and unpack of batch k+1 overlap the training step of batch k.

``train_model`` / ``_evaluate`` (pipelines/model_training/nodes.py:784, 914) consume the loader unchanged: ``len(loader)``,
``len(loader.dataset)``, iteration, ``{k: v.to(device)}`` (a no-op here).  A yielded batch stays valid until ``max(1, depth - 2)``
further batches have been requested (the buffers are a ring; one iterator at a time).  There is no CPU fall-back: building a loader needs CUDA.
"""
from __future__ import annotations

import pickle
import queue
import threading
import time
from typing import Callable, Dict, Iterator, List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import ops
from ._lib import VbError

BOX_DIV, AREA_DIV = 1000.0, 1000000.0          # lmdb_dataset.py:193-203


# ------------------------------------------------------------------------------------------------ record sources
class LMDBRecords:
    """Region records of Facebook's detectron.lmdb: ``get(key) -> pickle bytes | None`` (``txn.get`` of an open environment).
    Key spellings, record key names and the zero fall-backs follow lmdb_dataset.py:126-179."""
    box_width, boxes_are_raw = 4, True

    def __init__(self, get: Callable[[bytes], Optional[bytes]], num_regions: int = 100, feature_dim: int = 2048):
        self.get, self.num_regions, self.feature_dim = get, num_regions, feature_dim

    @classmethod
    def open(cls, lmdb_path: str, num_regions: int = 100, feature_dim: int = 2048) -> "LMDBRecords":
        import lmdb                                                    # same flags as lmdb_dataset.py:113-120
        env = lmdb.open(lmdb_path, readonly=True, max_readers=1, lock=False, readahead=False, meminit=False)
        txn = env.begin(write=False)
        self = cls(txn.get, num_regions, feature_dim)
        self._env, self._txn = env, txn
        return self

    def record(self, img_id: str):
        for key in (img_id, f"{img_id}.png", img_id.zfill(5)):         # :130-135 (the first two spellings encode equal)
            item = self.get(key.encode())
            if item is not None:
                return pickle.loads(item)
        return None

    def fetch(self, img_id: str, feat_out: np.ndarray, box_out: np.ndarray) -> None:
        rec = self.record(img_id)
        features = boxes = None
        if isinstance(rec, dict):
            for k in ("features", "feature", "fc6"):
                features = rec.get(k)
                if features is not None:
                    break
            boxes = rec.get("boxes")
            if boxes is None:
                boxes = rec.get("bbox")
        elif rec is not None:
            features = rec
        if features is None:
            feat_out.fill(0)
        else:
            f = np.asarray(features, dtype=np.float32)
            if f.shape != feat_out.shape:
                raise VbError(f"record {img_id!r}: features of shape {f.shape}, expected {feat_out.shape} "
                              "(the reference's collate cannot stack ragged records either)")
            np.copyto(feat_out, f)
        b = None if boxes is None else np.asarray(boxes, dtype=np.float32)
        if b is None or b.ndim != 2 or b.shape[1] < 4:                 # :186-191 -> zero spatial rows
            box_out.fill(0)
        else:
            if b.shape[0] != box_out.shape[0]:
                raise VbError(f"record {img_id!r}: {b.shape[0]} boxes, expected {box_out.shape[0]}")
            np.copyto(box_out, b[:, :4])


class ArrayRecords:
    """HDF5-layout store (precomputed_dataset.py:78-99): ``visual[row]`` fp32 [R, F], ``spatial[row]`` fp32 [R, 5] (already
    normalised) and ``id_map`` image id -> row.  Anything indexable by row works (h5py datasets, ``np.memmap``, arrays)."""
    box_width, boxes_are_raw = 5, False

    def __init__(self, visual, spatial, id_map: Dict[str, int], num_regions: int = 100, feature_dim: int = 2048):
        self.visual, self.spatial, self.id_map = visual, spatial, id_map
        self.num_regions, self.feature_dim = num_regions, feature_dim

    @classmethod
    def open(cls, features_path: str, id_map_path: str, num_regions: int = 100, feature_dim: int = 2048) -> "ArrayRecords":
        import h5py
        f = h5py.File(features_path, "r")
        self = cls(f["visual_features"], f["spatial_features"], np.load(id_map_path, allow_pickle=True).item(), num_regions,
                   feature_dim)
        self._file = f
        return self

    def fetch(self, img_id: str, feat_out: np.ndarray, box_out: np.ndarray) -> None:
        if img_id in self.id_map:
            i = self.id_map[img_id]
            np.copyto(feat_out, np.asarray(self.visual[i], dtype=np.float32))
            np.copyto(box_out, np.asarray(self.spatial[i], dtype=np.float32))
        else:
            feat_out.fill(0)
            box_out.fill(0)


# ------------------------------------------------------------------------------------------------ text columns
class TextTable:
    """``id / text / label`` columns of the split, tokenised once.  Per-row semantics of lmdb_dataset.py:211-238:
    ``str(row["id"])``, ``str(row.get("text", ""))``, ``int(row.get("label", 0))``, padding to ``max_seq_length`` with
    truncation, token types defaulting to zeros."""

    def __init__(self, data, tokenizer, max_seq_length: int = 128, chunk: int = 2048):
        data = data.reset_index(drop=True)
        rows = [data.iloc[i] for i in range(len(data))]
        self.img_ids: List[str] = [str(r["id"]) for r in rows]
        texts = [str(r.get("text", "")) for r in rows]
        self.labels = np.array([int(r.get("label", 0)) for r in rows], dtype=np.int64).reshape(len(rows))
        n, t = len(rows), max_seq_length
        self.input_ids = np.zeros((n, t), np.int64)
        self.attention_mask = np.zeros((n, t), np.int64)
        self.token_type_ids = np.zeros((n, t), np.int64)
        for s in range(0, n, chunk):
            enc = tokenizer(texts[s:s + chunk], max_length=t, padding="max_length", truncation=True, return_tensors="np")
            self.input_ids[s:s + chunk] = enc["input_ids"]
            self.attention_mask[s:s + chunk] = enc["attention_mask"]
            if "token_type_ids" in enc:
                self.token_type_ids[s:s + chunk] = enc["token_type_ids"]
        self.max_seq_length = t

    def __len__(self) -> int:
        return len(self.img_ids)


# ------------------------------------------------------------------------------------------------ batch blob
class BatchLayout:
    """Byte layout of one batch blob (identical on the host and in HBM).  Sections start on 256-byte boundaries."""
    KEYS = ("features", "boxes", "input_ids", "attention_mask", "token_type_ids", "labels")

    def __init__(self, batch: int, seq_len: int, regions: int, feature_dim: int, box_width: int):
        self.shapes = {"features": (batch, regions, feature_dim), "boxes": (batch, regions, box_width),
                       "input_ids": (batch, seq_len), "attention_mask": (batch, seq_len),
                       "token_type_ids": (batch, seq_len), "labels": (batch,)}
        self.dtypes = {k: torch.float32 if k in ("features", "boxes") else torch.int64 for k in self.KEYS}
        self.offsets, off = {}, 0
        for k in self.KEYS:
            self.offsets[k] = off
            nbytes = int(np.prod(self.shapes[k])) * (4 if self.dtypes[k] == torch.float32 else 8)
            off += (nbytes + 255) // 256 * 256
        self.nbytes = off
        self.batch = batch

    def views(self, blob: torch.Tensor) -> Dict[str, torch.Tensor]:
        """Typed views of a uint8 blob (host or device)."""
        out = {}
        for k in self.KEYS:
            shape, dt = self.shapes[k], self.dtypes[k]
            nbytes = int(np.prod(shape)) * dt.itemsize
            out[k] = blob[self.offsets[k]:self.offsets[k] + nbytes].view(dt).view(shape)
        return out


def pack_batch(table: TextTable, records, indices: Sequence[int], host: Dict[str, np.ndarray]) -> None:
    """Fill the numpy views of one host blob with the samples ``indices`` (pure host work, no CUDA).  One thread: spreading
    the per-sample decodes over a thread pool gave no gain (2.8 ms per 16 x 100 x 2048 batch either way on the authoring host)."""
    idx = np.asarray(indices, dtype=np.int64)
    host["input_ids"][...] = table.input_ids[idx]
    host["attention_mask"][...] = table.attention_mask[idx]
    host["token_type_ids"][...] = table.token_type_ids[idx]
    host["labels"][...] = table.labels[idx]
    for j, i in enumerate(indices):
        records.fetch(table.img_ids[i], host["features"][j], host["boxes"][j])


def epoch_order(n: int, shuffle: bool) -> List[int]:
    """Sample order of one epoch.  With ``shuffle`` it consumes the global torch RNG exactly as
    ``iter(DataLoader(..., shuffle=True))`` does (one draw for the iterator's base seed, one for ``RandomSampler``'s private
    generator, then ``randperm``), so that under the same ``torch.manual_seed`` the batches are the reference's batches."""
    if not shuffle:
        return list(range(n))
    torch.empty((), dtype=torch.int64).random_()                                   # _BaseDataLoaderIter._base_seed
    seed = int(torch.empty((), dtype=torch.int64).random_().item())                # RandomSampler.__iter__
    g = torch.Generator()
    g.manual_seed(seed)
    return torch.randperm(n, generator=g).tolist()


def shard_order(n: int, shuffle: bool, rank: int, world_size: int, seed: int = 0, epoch: int = 0) -> List[int]:
    """This rank's sample order for data-parallel training (disjoint shards, one process per GPU): the index
    arithmetic of ``torch.utils.data.distributed.DistributedSampler(drop_last=False)`` — a permutation seeded by
    ``seed + epoch`` shared by all ranks, padded by wrapping to a multiple of ``world_size``, then every ``world_size``-th
    index starting at ``rank``.  Every rank gets the same number of samples, so the ranks' collectives stay in step."""
    if shuffle:
        g = torch.Generator()
        g.manual_seed(seed + epoch)
        order = torch.randperm(n, generator=g).tolist()
    else:
        order = list(range(n))
    if n == 0:
        return []
    total = (n + world_size - 1) // world_size * world_size
    pad = total - n
    order += (order * ((pad + n - 1) // n))[:pad]
    return order[rank:total:world_size]


def batch_plan(order: Sequence[int], batch_size: int, drop_last: bool) -> List[List[int]]:
    out = [list(order[i:i + batch_size]) for i in range(0, len(order), batch_size)]
    if drop_last and out and len(out[-1]) < batch_size:
        out.pop()
    return out


# ------------------------------------------------------------------------------------------------ the loader
class _Slot:
    def __init__(self, layout: BatchLayout, device, feature_dtype):
        self.host = torch.empty(layout.nbytes, dtype=torch.uint8).pin_memory()
        self.dev = torch.empty(layout.nbytes, dtype=torch.uint8, device=device)
        b, r, f = layout.shapes["features"]
        self.feat16 = torch.empty(b, r, f, dtype=torch.bfloat16, device=device) if feature_dtype == torch.bfloat16 else None
        self.spatial = torch.empty(b, r, 5, dtype=torch.float32, device=device)
        self.ready = torch.cuda.Event()
        self.released = torch.cuda.Event()
        self.released.record()
        self._views: Dict[int, tuple] = {}

    def views(self, lay: BatchLayout, boxes_are_raw: bool):
        """(numpy views of the pinned blob, torch views of the HBM blob, the batch dict handed to the consumer), cached per
        batch size so that neither thread builds tensor views in steady state."""
        got = self._views.get(lay.batch)
        if got is None:
            host = {k: v.numpy() for k, v in lay.views(self.host).items()}
            dev, b = lay.views(self.dev), lay.batch
            batch = {"input_ids": dev["input_ids"], "attention_mask": dev["attention_mask"],
                     "token_type_ids": dev["token_type_ids"],
                     "visual_features": self.feat16[:b] if self.feat16 is not None else dev["features"],
                     "spatial_locations": self.spatial[:b] if boxes_are_raw else dev["boxes"], "labels": dev["labels"]}
            got = self._views[lay.batch] = (host, dev, batch)
        return got


class FeatureStoreLoader:
    """Iterable of reference-shaped batch dicts resident on ``device`` (see the module docstring).

    ``rank`` / ``world_size`` > 1 shard every epoch like ``DistributedSampler`` (one loader per process / GPU, call
    ``set_epoch`` each epoch); with one process the order is the reference ``DataLoader``'s.
    ``feature_dtype=torch.bfloat16`` (default) hands the encoder its GEMM operand directly; ``torch.float32`` yields batches
    bit-identical to the reference loader's (the encoder then rounds them itself — same result)."""

    def __init__(self, data, records, tokenizer, max_seq_length: int = 128, batch_size: int = 32, shuffle: bool = False,
                 drop_last: bool = False, device="cuda", depth: int = 3, feature_dtype: torch.dtype = torch.bfloat16,
                 start_delay_ms: float = 1.0, rank: int = 0, world_size: int = 1, seed: int = 0):
        if not torch.cuda.is_available():
            raise VbError("FeatureStoreLoader stages batches in HBM and needs a CUDA device; there is no CPU fall-back")
        if depth < 2:
            raise VbError("depth must be at least 2")
        if feature_dtype not in (torch.bfloat16, torch.float32):
            raise VbError("feature_dtype must be bfloat16 or float32")
        self.dataset = TextTable(data, tokenizer, max_seq_length)
        self.records, self.batch_size, self.shuffle, self.drop_last = records, batch_size, shuffle, drop_last
        if not 0 <= rank < world_size:
            raise VbError(f"rank {rank} outside world of {world_size}")
        self.rank, self.world_size, self.seed, self.epoch = rank, world_size, seed, 0
        self.device = torch.device(device if str(device) != "cuda" else f"cuda:{torch.cuda.current_device()}")
        self.depth, self.feature_dtype, self.start_delay = depth, feature_dtype, start_delay_ms * 1e-3
        self._layouts: Dict[int, BatchLayout] = {}
        with torch.cuda.device(self.device):
            self._stream = torch.cuda.Stream()
            full = self._layout(batch_size)
            self._slots = [_Slot(full, self.device, feature_dtype) for _ in range(depth)]
        self.h2d_bytes = 0            # bytes shipped host -> device so far (for the bench's e2e accounting)

    def set_epoch(self, epoch: int) -> None:
        """Data-parallel shuffling is seeded by ``seed + epoch`` (as ``DistributedSampler.set_epoch``)."""
        self.epoch = epoch

    def _order(self) -> List[int]:
        if self.world_size == 1:
            return epoch_order(len(self.dataset), self.shuffle)        # the reference loader's order (global torch RNG)
        return shard_order(len(self.dataset), self.shuffle, self.rank, self.world_size, self.seed, self.epoch)

    def __len__(self) -> int:
        n = (len(self.dataset) + self.world_size - 1) // self.world_size
        return n // self.batch_size if self.drop_last else (n + self.batch_size - 1) // self.batch_size

    def _layout(self, b: int) -> BatchLayout:
        lay = self._layouts.get(b)
        if lay is None:
            lay = self._layouts[b] = BatchLayout(b, self.dataset.max_seq_length, self.records.num_regions,
                                                 self.records.feature_dim, self.records.box_width)
        return lay

    # producer side -------------------------------------------------------------------------------------------------
    def _produce(self, plan: List[List[int]], out: "queue.Queue", free: "queue.Queue", stop: threading.Event) -> None:
        try:
            torch.cuda.set_device(self.device)
            for indices in plan:
                try:
                    slot_id = free.get_nowait()
                except queue.Empty:
                    slot_id = free.get()                          # woken by the consumer asking for its next batch:
                    if self.start_delay > 0:                      # stay off the GIL while it launches that step
                        time.sleep(self.start_delay)
                if stop.is_set() or slot_id is None:
                    return
                slot = self._slots[slot_id]
                slot.released.synchronize()                       # the consumer's work on this slot's last batch is done
                slot.ready.synchronize()                          # ... and so is its copy, had the epoch been abandoned
                lay = self._layout(len(indices))
                host, dev, _ = slot.views(lay, self.records.boxes_are_raw)
                pack_batch(self.dataset, self.records, indices, host)
                with torch.cuda.stream(self._stream):
                    slot.dev[:lay.nbytes].copy_(slot.host[:lay.nbytes], non_blocking=True)
                    raw, b = self.records.boxes_are_raw, lay.batch
                    feat16 = slot.feat16[:b] if slot.feat16 is not None else None
                    if feat16 is not None or raw:                 # one launch: bf16 features and/or normalised boxes
                        ops.lmdb_regions(dev["features"] if feat16 is not None else None, feat16,
                                         dev["boxes"] if raw else None, slot.spatial[:b] if raw else None,
                                         BOX_DIV, AREA_DIV, stream=self._stream)
                    slot.ready.record(self._stream)
                self.h2d_bytes += lay.nbytes
                out.put((slot_id, lay))
            out.put(None)
        except BaseException as e:                                 # surface producer failures in the consumer
            out.put(e)

    # consumer side -------------------------------------------------------------------------------------------------
    def __iter__(self) -> Iterator[Dict[str, torch.Tensor]]:
        plan = batch_plan(self._order(), self.batch_size, self.drop_last)
        out: "queue.Queue" = queue.Queue()
        free: "queue.Queue" = queue.Queue()
        for i in range(self.depth):
            free.put(i)
        stop = threading.Event()
        worker = threading.Thread(target=self._produce, args=(plan, out, free, stop), daemon=True, name="vb-ingest")
        worker.start()
        held: List[int] = []
        try:
            while True:
                item = out.get()
                if item is None:
                    break
                if isinstance(item, BaseException):
                    raise item
                slot_id, lay = item
                slot = self._slots[slot_id]
                cur = torch.cuda.current_stream(self.device)
                cur.wait_event(slot.ready)
                batch = dict(slot.views(lay, self.records.boxes_are_raw)[2])
                held.append(slot_id)
                if len(held) > max(1, self.depth - 2):             # hand the oldest slot back once its work is enqueued
                    old = held.pop(0)
                    self._slots[old].released.record(cur)
                    free.put(old)
                yield batch
        finally:
            stop.set()
            free.put(None)
            cur = torch.cuda.current_stream(self.device)
            for s in held:
                self._slots[s].released.record(cur)
            worker.join(timeout=30)


# ------------------------------------------------------------------------------------------------ reference-named factories
def _bert_tokenizer():
    from transformers import BertTokenizer
    return BertTokenizer.from_pretrained("bert-base-uncased")          # lmdb_dataset.py:273


def _three(train_data, val_data, test_data, records, tokenizer, batch_size, max_seq_length, device, depth):
    """Train: shuffled, drop_last, sharded over the ranks of an initialised process group (one process per GPU);
    validation / test: sequential and whole on every rank, as the reference evaluates them."""
    import torch.distributed as dist
    rank, world = (dist.get_rank(), dist.get_world_size()) if dist.is_available() and dist.is_initialized() else (0, 1)

    def make(data, train):
        return FeatureStoreLoader(data, records, tokenizer, max_seq_length, batch_size, shuffle=train, drop_last=train,
                                  device=device, depth=depth, rank=rank if train else 0, world_size=world if train else 1)
    return make(train_data, True), make(val_data, False), make(test_data, False)


def create_lmdb_dataloaders(train_data, val_data, test_data, lmdb_path: str = "data/03_features/mmf/detectron.lmdb",
                            batch_size: int = 32, max_seq_length: int = 128, num_regions: int = 100,
                            visual_feature_dim: int = 2048, num_workers: int = 0, auto_download: bool = True, *,
                            device="cuda", tokenizer=None, records=None, depth: int = 3
                            ) -> Tuple[FeatureStoreLoader, FeatureStoreLoader, FeatureStoreLoader]:
    """Signature and loader settings of lmdb_dataset.py:249-319 (train: shuffle + drop_last; val/test: sequential).
    ``num_regions`` / ``visual_feature_dim`` are ignored as there (the store is fixed at 100 x 2048); ``num_workers`` is
    meaningless here (one producer thread); ``auto_download`` is not honoured — a missing store is an error."""
    records = records if records is not None else LMDBRecords.open(lmdb_path, 100, 2048)
    return _three(train_data, val_data, test_data, records, tokenizer or _bert_tokenizer(), batch_size, max_seq_length,
                  device, depth)


def create_precomputed_dataloaders(train_data, val_data, test_data, features_path: str, id_map_path: str,
                                   batch_size: int = 32, max_seq_length: int = 128, num_regions: int = 100,
                                   visual_feature_dim: int = 2048, num_workers: int = 0, *, device="cuda", tokenizer=None,
                                   records=None, depth: int = 3
                                   ) -> Tuple[FeatureStoreLoader, FeatureStoreLoader, FeatureStoreLoader]:
    """Signature and loader settings of precomputed_dataset.py:134-228."""
    records = records if records is not None else ArrayRecords.open(features_path, id_map_path, num_regions,
                                                                    visual_feature_dim)
    return _three(train_data, val_data, test_data, records, tokenizer or _bert_tokenizer(), batch_size, max_seq_length,
                  device, depth)
