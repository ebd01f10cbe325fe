"""Where does the end-to-end step (host buffers in, loss out) spend its time?  Host wall clock per phase, with a device
synchronize after each phase so that CPU-side and GPU-side costs separate."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from multimodal_classification_b200.vilbert import ViLBERTForClassification, get_facebook_vilbert_config
from oracle import vilbert_oracle as vo

dev = torch.device("cuda")
cfg = get_facebook_vilbert_config()
torch.manual_seed(0)
model = ViLBERTForClassification(cfg, num_labels=2).to(dev).train()
host = vo.synthetic_batch(cfg, batch=16, seq=128, regions=100, seed=1234)
pinned = {k: v.pin_memory() for k, v in host.items()}
def sync(): torch.cuda.synchronize()
acc = {}
def tick(name, t0):
    sync(); acc[name] = acc.get(name, 0.0) + (time.perf_counter() - t0)
for it in range(25):
    if it == 5: acc.clear()
    t = time.perf_counter(); b = {k: v.to(dev, non_blocking=True) for k, v in pinned.items()}; tick("h2d", t)
    t = time.perf_counter(); model.zero_grad(set_to_none=True); tick("zero_grad", t)
    t = time.perf_counter(); model._engine and setattr(model._engine.flat, "_version", -1); out = model(**b); t1 = time.perf_counter() - t; tick("forward", t)
    acc["forward_cpu"] = acc.get("forward_cpu", 0.0) + t1
    t = time.perf_counter(); out["loss"].backward(); t1 = time.perf_counter() - t; tick("backward", t)
    acc["backward_cpu"] = acc.get("backward_cpu", 0.0) + t1
    t = time.perf_counter(); out["loss"].item(); tick("item", t)
for k, v in acc.items():
    print(f"{k:14s} {v / 20 * 1e3:7.3f} ms/step")
