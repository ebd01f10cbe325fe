"""Data-parallel training of the ViLBERT replica: one process per GPU, gradients averaged with NCCL all-reduce over
contiguous ranges ("buckets") of the flat fp32 gradient buffer.  The reference has no multi-GPU code at all (SURVEY.md
§2.1); the batch shards naturally (§8e), so the only exchange step is this gradient all-reduce.

Buckets follow the order in which the backward pass finishes blocks (classifier/poolers first, text layer 11, co-layer
5, visual layer 5, ... embeddings last).  Each all-reduce is issued on a dedicated communication stream as soon as the
producing stream(s) have written the bucket, so NCCL traffic over NVLink overlaps the remaining backward kernels; when
the backward pass is captured in a CUDA graph the collectives are captured with it (same graph, parallel branch).
"""
from __future__ import annotations

from typing import Dict, List, Tuple

import torch
import torch.distributed as dist


def block_ranges(offsets: Dict[str, int], sizes: Dict[str, int], order: List[Tuple[str, List[str]]], s_end: int):
    """[(bucket name, lo, hi)] for contiguous runs of parameters; `order` lists, per bucket, the parameter keys it holds.
    The last bucket is extended to `s_end` (small parameters and embedding tables, finished last)."""
    out = []
    for name, keys in order:
        lo = min(offsets[k] for k in keys)
        hi = max(offsets[k] + sizes[k] for k in keys)
        out.append([name, lo, hi])
    out.sort(key=lambda b: b[1])
    for a, b in zip(out, out[1:]):
        assert a[2] <= b[1], ("overlapping buckets", a, b)
        a[2] = b[1]          # absorb alignment padding
    out[-1][2] = s_end
    assert out[0][1] == 0
    return {n: (lo, hi) for n, lo, hi in out}


def all_reduce_mean(t: torch.Tensor, group) -> None:
    """In-place mean over the ranks of `group` (NCCL: ReduceOp.AVG; gloo, used by the CPU tests: SUM then scale)."""
    if dist.get_backend(group) == "nccl":
        dist.all_reduce(t, op=dist.ReduceOp.AVG, group=group)
    else:
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
        t.mul_(1.0 / dist.get_world_size(group))


def attach(model, group=None) -> None:
    """Make `model` (multimodal_classification_b200.vilbert.ViLBERTForClassification) average its gradients over
    `group` inside every backward pass.  Parameters must already be identical on all ranks (same seed / same
    state_dict); call broadcast_parameters() otherwise."""
    model._ddp_group = group if group is not None else dist.group.WORLD
    model._engine = None


def broadcast_parameters(model, group=None, src: int = 0) -> None:
    group = group if group is not None else dist.group.WORLD
    with torch.no_grad():
        for p in model.parameters():
            dist.broadcast(p.data, src=src, group=group)


def shutdown(*models, timeout_s: float = 10.0) -> bool:
    """Tear the process group down at the end of a run.  CUDA graphs that captured NCCL collectives keep the communicator
    busy, so the models' engines (and with them the graphs) are released first; ``destroy_process_group`` is then given
    `timeout_s` seconds on a helper thread.  Returns False when it did not finish (the caller should then leave with
    ``os._exit``: the work is done, only the teardown is stuck)."""
    import gc
    import threading
    for m in models:
        m._engine = None
    gc.collect()
    if torch.cuda.is_available():
        torch.cuda.synchronize()
    t = threading.Thread(target=dist.destroy_process_group, daemon=True)
    t.start()
    t.join(timeout_s)
    return not t.is_alive()
