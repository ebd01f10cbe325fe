"""A/B of the feature-store loader settings on one GPU: the bench's train step fed by
ingest.FeatureStoreLoader for several producer start delays / ring depths, beside the plain pinned-copy-per-step loop.
    python tools/bench_ingest.py [steps]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from multimodal_classification_b200.vilbert import ViLBERTForClassification, get_facebook_vilbert_config  # noqa: E402
from oracle import vilbert_oracle as vo  # noqa: E402  (synthetic batch generator only)


def main():
    steps = int(sys.argv[1]) if len(sys.argv) > 1 else 60
    dev = torch.device("cuda", 0)
    cfg = get_facebook_vilbert_config()
    torch.manual_seed(0)
    model = ViLBERTForClassification(cfg, num_labels=2).to(dev).train()
    params = list(model.parameters())
    host = vo.synthetic_batch(cfg, batch=bench.B, seq=bench.T, regions=bench.R, seed=1234)
    pinned = {k: v.pin_memory() for k, v in host.items()}

    def step(batch):
        for p in params:
            p.grad = None
        if model._engine is not None:
            model._engine.flat._version = -1
        out = model(**batch)
        out["loss"].backward()
        return out["loss"]

    def timed(fn, n):
        torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(n):
            fn()
        e.record()
        torch.cuda.synchronize()
        return s.elapsed_time(e)

    def pinned_step():
        return step({k: v.to(dev, non_blocking=True) for k, v in pinned.items()}).item()
    for _ in range(5):
        pinned_step()
    print("pinned copy per step      : %.3f ms/step" % (timed(pinned_step, steps) / steps), flush=True)
    for kw in ({"start_delay_ms": 0.0}, {"start_delay_ms": 1.0}, {"start_delay_ms": 2.5}, {"start_delay_ms": 1.0, "depth": 2},
               {"start_delay_ms": 1.0, "depth": 5}):
        r = bench.time_ingest(torch, dev, step, timed, steps, **kw)
        print("loader %-28s: %.3f ms/step  (loader alone %.0f samples/s)" % (kw, r["ms_per_step"], r["loader_only_samples_per_sec"]),
              flush=True)
    print("pinned copy per step again: %.3f ms/step" % (timed(pinned_step, steps) / steps), flush=True)


if __name__ == "__main__":
    main()
