"""Inference sweep (BASELINE.json configs[4]): ViLBERT-base eval forward, bs 16..512, 128 tokens x 100 regions, one B200."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from multimodal_classification_b200.vilbert import ViLBERTForClassification, get_facebook_vilbert_config
from oracle import vilbert_oracle as vo

cfg = get_facebook_vilbert_config()
torch.manual_seed(0)
model = ViLBERTForClassification(cfg, num_labels=2).cuda().eval()
FWD_GF = 50.83
for bs in (16, 32, 64, 128, 256, 512):
    batch = {k: v.cuda() for k, v in vo.synthetic_batch(cfg, batch=bs, seq=128, regions=100, seed=1234).items() if k != "labels"}
    with torch.no_grad():
        for _ in range(3):
            out = model(**batch)
        torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        iters = 10
        s.record()
        for _ in range(iters):
            out = model(**batch)
        e.record(); torch.cuda.synchronize()
    ms = s.elapsed_time(e) / iters
    assert torch.isfinite(out["logits"]).all()
    print(json.dumps({"batch": bs, "ms": round(ms, 3), "samples_per_s": round(bs / ms * 1e3, 1), "tflops": round(FWD_GF * bs / ms, 1)}), flush=True)
