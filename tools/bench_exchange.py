"""Gradient exchange alone: mean all-reduce of the whole bf16 gradient buffer (249 M elements = 498 MB) over N GPUs, our
NVSwitch kernels (csrc/comm.cu) against NCCL on the same buffer.  CUDA events, max over ranks; bus bandwidth as nccl-tests
define it (2 (N-1)/N x bytes / time).

    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 --master-port 29620 tools/bench_exchange.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
from multimodal_classification_b200 import ddp

rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
n = 249_000_000 // 1024 * 1024

def timed(fn, iters=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize(); dist.barrier()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(iters):
        fn()
    e.record(); torch.cuda.synchronize()
    t = torch.tensor([s.elapsed_time(e) / iters], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t.item()

def report(name, ms, nbytes):
    if rank == 0:
        print(f"{name:46s} {ms * 1e3:8.1f} us   algbw {nbytes / ms / 1e6:7.1f} GB/s   busbw {2 * (world - 1) / world * nbytes / ms / 1e6:7.1f} GB/s", flush=True)

x = torch.ones(n, dtype=torch.bfloat16, device=dev)
report(f"NCCL all-reduce bf16 {n * 2 / 1e6:.0f} MB", timed(lambda: dist.all_reduce(x, op=dist.ReduceOp.AVG)), n * 2)
if ddp.switch_available(dist.group.WORLD, dev):
    ex = ddp.SwitchExchange(n, dev, dist.group.WORLD)
    if rank == 0:
        print("switch exchange: multicast mapping", ex.multicast, flush=True)
    for ctas in (16, 32, 48, 96, 148):
        ex.args.ctas = ctas
        report(f"switch all-reduce bf16 in place, {ctas} CTAs", timed(lambda: ex.all_reduce_mean(0, n)), n * 2)
    for mb in (20, 80):
        m = mb * (1 << 20) // 2
        ex.args.ctas = 48
        report(f"switch all-reduce bf16 in place, {mb} MB message", timed(lambda: ex.all_reduce_mean(0, m)), m * 2)
    exf = ddp.SwitchExchange(n, dev, dist.group.WORLD, fp32_out=True)
    exf.args.ctas = 48
    report("switch all-reduce bf16 -> fp32 broadcast, 48 CTAs", timed(lambda: exf.all_reduce_mean(0, n)), n * 2)
else:
    print("switch transport unavailable")
dist.barrier()
dist.destroy_process_group()
