"""Two-rank data-parallel smoke: gradients after the bucket all-reduce must equal the mean of the per-rank gradients.
Launch: python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 tools/ddp_smoke.py"""
import os, sys, faulthandler
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
faulthandler.dump_traceback_later(60, exit=True)
import torch, torch.distributed as dist
from multimodal_classification_b200.vilbert import ViLBERTForClassification
from multimodal_classification_b200 import ddp
from oracle import vilbert_oracle as vo

rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
def say(*a):
    print(f"[rank {rank}]", *a, flush=True)
cfg = vo.tiny_config()
torch.manual_seed(0)
model = ViLBERTForClassification(cfg, num_labels=2).to(dev).eval()
solo = ViLBERTForClassification(cfg, num_labels=2).to(dev).eval()
solo.load_state_dict(model.state_dict())
compress = None if os.environ.get("VB_DDP_FP32", "0") == "1" else "bf16"
transport = os.environ.get("VB_DDP_TRANSPORT", "nccl")
ddp.attach(model, dist.group.WORLD, compress=compress, transport=transport)
batch = {k: v.to(dev) for k, v in vo.synthetic_batch(cfg, batch=4, seq=32, regions=20, seed=50 + rank).items()}
say("built", "transport", transport)
for step in range(4):
    model.zero_grad(set_to_none=True)
    out = model(**batch)
    say("fwd", step, float(out["loss"]))
    out["loss"].backward()
    torch.cuda.synchronize()
    say("bwd", step)
# reference: un-attached replica's own gradient, averaged over ranks by hand
solo.zero_grad(set_to_none=True)
solo(**batch)["loss"].backward()
worst = 0.0
for (k, p), (_, q) in zip(model.named_parameters(), solo.named_parameters()):
    if q.grad is None:
        continue
    g = q.grad.clone()
    dist.all_reduce(g, op=dist.ReduceOp.AVG)
    worst = max(worst, ((p.grad - g).abs().max() / (g.abs().max() + 1e-12)).item())
if model._engine.comm_switch is not None:
    say("switch exchange: multicast", model._engine.comm_switch.multicast)
say(f"worst relative gradient mismatch vs hand-averaged: {worst:.6e} ")
# bf16 exchange: one extra rounding (2^-8); through the switch the local gradient is rounded too (written in bf16 by the GEMM)
assert worst < (1e-3 if compress is None else (1.5e-2 if transport == "switch" else 1e-2)), worst
dist.barrier()
say("OK")
clean = ddp.shutdown(model, solo)
say("process group destroyed" if clean else "destroy_process_group stuck: leaving with os._exit")
sys.stdout.flush()
os._exit(0)
