"""How long do the gradient-bucket all-reduces take on their own (no compute)?  torchrun --nproc-per-node N tools/nccl_probe.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
rank, lr = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr); dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
n = 248_826_882
for dtype in (torch.float32, torch.bfloat16):
    buf = torch.ones(n, dtype=dtype, device=dev)
    for nb in (1, 27):
        chunks = buf.chunk(nb)
        for _ in range(3):
            for c in chunks: dist.all_reduce(c, op=dist.ReduceOp.AVG)
        torch.cuda.synchronize(); dist.barrier()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(5):
            for c in chunks: dist.all_reduce(c, op=dist.ReduceOp.AVG)
        e.record(); torch.cuda.synchronize()
        ms = s.elapsed_time(e) / 5
        if rank == 0:
            gb = n * buf.element_size() / 1e9
            print(f"{dtype} {nb:2d} buckets: {ms:6.2f} ms  algbw {gb / ms * 1e3:6.0f} GB/s", flush=True)
dist.barrier()
os._exit(0)
