"""Host-side timeline of the end-to-end step (pinned batch in, loss.item() out): when, after the previous step's loss has
been read, do the forward / backward graph launches leave the host?  CPU clock only, no extra synchronisation."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from multimodal_classification_b200.vilbert import ViLBERTForClassification, get_facebook_vilbert_config
from oracle import vilbert_oracle as vo

dev = torch.device("cuda")
cfg = get_facebook_vilbert_config()
torch.manual_seed(0)
model = ViLBERTForClassification(cfg, num_labels=2).to(dev).train()
params = list(model.parameters())
host = vo.synthetic_batch(cfg, batch=16, seq=128, regions=100, seed=1234)
pinned = {k: v.pin_memory() for k, v in host.items()}
resident = {k: v.to(dev) for k, v in host.items()}
marks = []
orig = torch.cuda.CUDAGraph.replay
def replay(self):
    marks.append(("replay>", time.perf_counter()))
    orig(self)
    marks.append(("replay<", time.perf_counter()))
torch.cuda.CUDAGraph.replay = replay

def run(mode, n=40):
    rows = []
    for it in range(n + 5):
        marks.clear()
        t0 = time.perf_counter()
        b = {k: v.to(dev, non_blocking=True) for k, v in pinned.items()} if mode == "pinned" else resident
        t_h2d = time.perf_counter()
        for p in params:
            p.grad = None
        model._engine and setattr(model._engine.flat, "_version", -1)
        t_zero = time.perf_counter()
        out = model(**b)
        t_fwd = time.perf_counter()
        out["loss"].backward()
        t_bwd = time.perf_counter()
        out["loss"].item()
        t_end = time.perf_counter()
        if it >= 5 and len(marks) == 4:
            rows.append([t_h2d - t0, t_zero - t0, marks[0][1] - t0, marks[1][1] - t0, t_fwd - t0, marks[2][1] - t0, marks[3][1] - t0,
                         t_bwd - t0, t_end - t0])
    names = ["h2d issued", "grads cleared", "fwd replay call", "fwd replay returned", "forward() returned", "bwd replay call",
             "bwd replay returned", "backward() returned", "loss.item() returned"]
    print(f"--- {mode}: median host clock since step start, {len(rows)} steps")
    for i, nm in enumerate(names):
        col = sorted(r[i] for r in rows)
        print(f"  {nm:22s} {col[len(col) // 2] * 1e3:7.3f} ms")

run("pinned")
run("resident")
