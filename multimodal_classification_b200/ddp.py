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


class SwitchExchange:
    """bf16 gradient exchange through NVSwitch by our own kernels (csrc/comm.cu) instead of NCCL: one symmetric bf16 buffer laid
    out like the flat gradient buffer -- the weight-gradient GEMMs write into it directly -- whose ranges are mean-all-reduced in
    place (multimem.ld_reduce / multimem.st when the fabric offers a multicast mapping, peer loads / stores otherwise).
    torch.distributed._symmetric_memory allocates and maps the memory (plumbing only)."""
    FLAG_SLOTS = 4

    def __init__(self, numel: int, device, group, fp32_out: bool = False):
        import torch.distributed._symmetric_memory as symm
        from . import _lib
        self.group = group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        n = (numel + 127) // 128 * 128
        try:                                             # older torch: groups are enabled explicitly (a no-op / deprecated later)
            symm.enable_symm_mem_for_group(group.group_name)
        except Exception:
            pass
        self.buf = symm.empty(n, dtype=torch.bfloat16, device=device)
        self.buf.zero_()
        self.handle = symm.rendezvous(self.buf, group)
        # optional second symmetric buffer: the fp32 gradient buffer itself, so that the reduced values are broadcast widened
        self.grad = self.grad_handle = None
        if fp32_out:
            self.grad = symm.empty(n, dtype=torch.float32, device=device)
            self.grad.zero_()
            self.grad_handle = symm.rendezvous(self.grad, group)
        self.flags = symm.empty(self.FLAG_SLOTS * self.world, dtype=torch.int32, device=device)
        self.flags.zero_()
        self.flag_handle = symm.rendezvous(self.flags, group)
        torch.cuda.synchronize(device)
        dist.barrier(group)                              # every rank's flags are zero before the first hand-shake
        a = self.args = _lib.ExchangeArgs()
        mc = int(self.handle.multicast_ptr or 0)         # 0 when the fabric / driver offers no multicast mapping
        self.multicast = mc != 0
        a.mc_base = mc if mc else None
        a.peer_bases = int(self.handle.buffer_ptrs_dev)
        a.flag_ptrs = int(self.flag_handle.buffer_ptrs_dev)
        a.rank, a.world, a.flag_slots, a.ctas = self.rank, self.world, self.FLAG_SLOTS, int(__import__("os").environ.get("VB_DDP_CTAS", "0"))
        if self.grad_handle is not None:
            gmc = int(self.grad_handle.multicast_ptr or 0)
            a.out_mc_base = gmc if (gmc and mc) else None
            a.out_peer_bases = int(self.grad_handle.buffer_ptrs_dev)

    def all_reduce_mean(self, lo: int, hi: int) -> None:
        """Mean over ranks of buf[lo:hi] (on the current stream; lo / hi multiples of 8): written to grad[lo:hi] in fp32 on every
        rank when the exchange owns the fp32 gradient buffer, else in place in bf16."""
        import ctypes as C
        from . import _lib
        stream = torch.cuda.current_stream().cuda_stream
        _lib.check(_lib.lib().vb_allreduce_mean_bf16(C.byref(self.args), lo, hi, lo, stream), "vb_allreduce_mean_bf16")


def switch_available(group=None, device=None) -> bool:
    """Collective probe: can every rank of `group` build (and use) a SwitchExchange?  False on boxes without symmetric-memory /
    peer-access support, so that callers can fall back to the NCCL transport."""
    group = group if group is not None else dist.group.WORLD
    device = device if device is not None else torch.device("cuda", torch.cuda.current_device())
    ok = 1
    try:
        ex = SwitchExchange(1 << 16, device, group, fp32_out=True)
        ex.buf.fill_(float(dist.get_rank(group) + 1))
        torch.cuda.synchronize(device)
        dist.barrier(group)
        ex.all_reduce_mean(0, 1 << 16)
        torch.cuda.synchronize(device)
        n = dist.get_world_size(group)
        ok = int(bool((ex.grad[:1 << 16] == (n + 1) / 2.0).all()))
        ex2 = SwitchExchange(1 << 16, device, group)                # and the in-place bf16 form
        ex2.buf.fill_(float(dist.get_rank(group) + 1))
        torch.cuda.synchronize(device)
        dist.barrier(group)
        ex2.all_reduce_mean(0, 1 << 16)
        torch.cuda.synchronize(device)
        ok &= int(bool((ex2.buf[:1 << 16].float() == (n + 1) / 2.0).all()))
    except Exception as e:      # noqa: BLE001 -- any failure means "not available here"
        print(f"[ddp] switch transport unavailable: {e!r}"[:300], flush=True)
        ok = 0
    t = torch.tensor([ok], device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MIN, group=group)
    return bool(t.item())


def attach(model, group=None, compress=None, transport: str = "nccl") -> None:
    """Make `model` (multimodal_classification_b200.vilbert.ViLBERTForClassification) average its gradients over
    `group` inside every backward pass.  Parameters must already be identical on all ranks (same seed / same
    state_dict); call broadcast_parameters() otherwise.

    compress=None (default): the exchange is in fp32, numerically the single-GPU step on the global batch.
    compress="bf16" (opt-in): every bucket is narrowed to bf16 by a cast kernel, all-reduced (498 MB instead of 995 MB per
    step over NVLink) and widened back into the fp32 gradient buffer -- the usual bf16 gradient-compression trade (one extra
    rounding of the averaged gradient, relative error <= 2^-8).
    transport="switch" (with compress="bf16"): no NCCL and no narrowing pass -- the weight-gradient GEMMs write bf16 straight
    into a symmetric buffer that our own NVSwitch kernels reduce in place (SwitchExchange); the result is widened into the fp32
    gradient views once."""
    if compress not in (None, "bf16"):
        raise ValueError("compress must be None or 'bf16'")
    if transport not in ("nccl", "switch"):
        raise ValueError("transport must be 'nccl' or 'switch'")
    if transport == "switch" and compress != "bf16":
        raise ValueError("transport='switch' exchanges bf16 gradients: pass compress='bf16'")
    model._ddp_group = group if group is not None else dist.group.WORLD
    model._ddp_compress = compress
    model._ddp_transport = transport
    model._engine = None


FLUSH_BYTES = 160 << 20     # exchange finished buckets once this many gradient bytes are waiting (NCCL transport)
SWITCH_FLUSH_BYTES = 80 << 20   # the same for the switch transport (fp32-equivalent bytes): its three launches cost ~30 us per message


def merge_ranges(ranges):
    """Sorted union of [lo, hi) ranges (adjacent buckets of the flat buffer become one message)."""
    out = []
    for lo, hi in sorted(ranges):
        if out and lo <= out[-1][1]:
            out[-1][1] = max(out[-1][1], hi)
        else:
            out.append([lo, hi])
    return [(lo, hi) for lo, hi in out]


def all_reduce_mean_ranges(grad: torch.Tensor, ranges, group, staging=None) -> None:
    """grad[lo:hi] <- mean over ranks for every range, as ONE coalesced collective launch.  With `staging` (a bf16 buffer
    shaped like grad) the exchange is done in bf16: cast kernel -> all-reduce -> widening kernel."""
    if staging is not None:
        from . import ops
        for lo, hi in ranges:
            ops.cast_bf16(grad[lo:hi], staging[lo:hi])
        msgs = [staging[lo:hi] for lo, hi in ranges]
    else:
        msgs = [grad[lo:hi] for lo, hi in ranges]
    if len(msgs) == 1 or dist.get_backend(group) != "nccl":
        for m in msgs:
            all_reduce_mean(m, group)
    else:
        try:
            with dist._coalescing_manager(group=group, device=grad.device, async_ops=False):
                for m in msgs:
                    dist.all_reduce(m, op=dist.ReduceOp.AVG, group=group)
        except (AttributeError, TypeError):      # torch without the (private) coalescing manager: one launch per message
            for m in msgs:
                all_reduce_mean(m, group)
    if staging is not None:
        for lo, hi in ranges:
            ops.cast_f32(staging[lo:hi], grad[lo:hi])


def all_reduce_mean_bf16(grad: torch.Tensor, staging: torch.Tensor, group) -> None:
    """grad (fp32, contiguous slice of the flat gradient buffer) <- mean over ranks, exchanged as bf16 through `staging`."""
    from . import ops
    ops.cast_bf16(grad, staging)
    all_reduce_mean(staging, group)
    ops.cast_f32(staging, grad)


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
