"""Inert stand-ins for the CUDA runtime objects the ingest loader touches (streams, events, pinned memory) and an oracle-backed
replacement of its one kernel, so that the loader's host logic runs in the GPU-less container.  Test infrastructure only."""
import contextlib

import torch

from oracle import ingest_oracle as io


class _Event:
    def record(self, stream=None):
        pass

    def synchronize(self):
        pass


class _Stream:
    cuda_stream = 0

    def wait_event(self, event):
        pass


def _regions(features=None, features_bf16=None, boxes=None, spatial=None, box_div=1000.0, area_div=1e6, stream=None):
    if features is not None:
        features_bf16.copy_(features.to(torch.bfloat16))
    if boxes is not None:
        b = boxes.reshape(-1, boxes.shape[-1]).numpy()
        spatial.copy_(torch.from_numpy(io.process_boxes(b, b.shape[0])).view(spatial.shape))


def apply(setattr_fn):
    """Install the stand-ins through ``setattr_fn(obj, name, value)`` (pytest's monkeypatch.setattr, or plain setattr in a
    spawned worker process)."""
    from multimodal_classification_b200 import ops
    for name, value in [("is_available", lambda: True), ("current_device", lambda: 0), ("set_device", lambda d: None),
                        ("device", lambda d: contextlib.nullcontext()), ("stream", lambda s: contextlib.nullcontext()),
                        ("current_stream", lambda d=None: _Stream()), ("Stream", _Stream), ("Event", _Event)]:
        setattr_fn(torch.cuda, name, value)
    setattr_fn(torch.Tensor, "pin_memory", lambda self: self)
    setattr_fn(ops, "lmdb_regions", _regions)
