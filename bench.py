#!/usr/bin/env python
"""Headline benchmark: ViLBERT-base training step (fwd+bwd, bf16 compute / fp32 master, dropout on as the reference
trains) on Hateful-Memes-shaped synthetic batches: bs=16 per GPU, 128 tokens, 100 regions x 2048-d + 5-d boxes
(BASELINE.json configs[1]; data-parallel over N GPUs = configs[4]).

    python bench.py --gpus N --steps K --warmup W          # this framework (CUDA path)
    python bench.py --impl reference ...                   # the CPU arm: the reference's algorithm (oracle port) on host cores

Prints ONE JSON line (rank 0).  See DESIGN.md "Measurement" for what each key means.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "ViLBERT train samples/sec"
UNIT = "samples/s"
FLOP_PER_SAMPLE_FWD_BWD = 152.07e9   # SURVEY.md §8d (hand-derived, equals torch's FlopCounterMode on the reference)
B, T, R = 16, 128, 100


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"bf16_burst": d.get("bf16_tflops", 1590.0), "bf16_sustained": d.get("bf16_tflops_sustained", 1400.0),
                "hbm": d.get("hbm_gbs", 6650.0), "source": "measured (MEASURED_PEAKS.json)"}
    return {"bf16_burst": 1590.0, "bf16_sustained": 1400.0, "hbm": 6650.0, "source": "fallback (B200_PROFILING.md)"}


class ClockSampler:
    """nvidia-smi sampling during the timed region (B200_PROFILING.md 'clocks line')."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.proc = None
        self.path = None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.index)], stdout=open(self.path, "w"),
                                         stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.proc is None:
            return out
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        rows = []
        for line in open(self.path):
            f = [x.strip() for x in line.split(",")]
            if len(f) >= 9:
                rows.append(f)
        os.unlink(self.path)
        if not rows:
            return out
        sm = sorted(float(r[1]) for r in rows if r[1].replace(".", "").isdigit())
        reasons = set()
        for r in rows:
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        out.update({"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": float(rows[0][2]), "reasons": sorted(reasons),
                    "samples": len(rows), "power_w_max": max(float(r[3]) for r in rows)})
        return out


def cpu_port_step(cfg, sd, batch):
    """One fwd+bwd of the oracle (CPU restatement of the reference's algorithm, fp32, autograd)."""
    from oracle import vilbert_oracle as vo
    vo.loss_and_grads(sd, cfg, batch)


def time_cpu_port(max_seconds: float, batch_size: int):
    import torch
    from oracle import vilbert_oracle as vo
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    cfg = vo.facebook_config()
    sd = vo.seeded_state_dict(cfg)
    batch = vo.synthetic_batch(cfg, batch=batch_size, seq=T, regions=R, seed=1234)
    cpu_port_step(cfg, sd, batch)   # warm-up
    times = []
    t_end = time.time() + max_seconds
    while len(times) < 3 and (time.time() < t_end or not times):
        t0 = time.time()
        cpu_port_step(cfg, sd, batch)
        times.append(time.time() - t0)
    times.sort()
    return batch_size / times[len(times) // 2], cores, len(times)


def run_reference_arm(args):
    """--impl reference: the reference's own algorithm for this path on the host cores.  The reference is a Python
    package that cannot travel to the GPU box, so this arm runs oracle/vilbert_oracle.py (the pinned restatement, kind
    "port") with every host thread; each step is fwd+bwd on a bounded sample (sub-batch) of the bs=16 workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    from oracle import vilbert_oracle as vo
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    cfg = vo.facebook_config()
    sd = vo.seeded_state_dict(cfg)
    budget = 150.0
    bs = B
    while True:
        batch = vo.synthetic_batch(cfg, batch=bs, seq=T, regions=R, seed=1234)
        t0 = time.time()
        cpu_port_step(cfg, sd, batch)
        probe = time.time() - t0
        if probe * (args.steps + args.warmup) <= budget or bs <= 1:
            break
        bs //= 2
    for _ in range(args.warmup):
        cpu_port_step(cfg, sd, batch)
    t0 = time.time()
    for _ in range(args.steps):
        cpu_port_step(cfg, sd, batch)
    dt = (time.time() - t0) / max(1, args.steps)
    value = bs / dt
    sample = f"fwd+bwd of the oracle port on a {bs}-sample slice of the bs={B} batch, eval-mode (no dropout RNG), fp32, {cores} threads"
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "vilbert_base_train_step_bs16_t128_r100", "batch_per_step": bs, "tokens": T, "regions": R},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def gemm_kernel_roofline(torch, peaks):
    """Live, in-process timing of the dominant kernel (gemm_bf16_kernel) on the most FLOP-heavy shape of the step
    (text FFN: [2048,768]x[768,3072], 18 launches forward and 36 backward-equivalents), CUDA events on the launch stream."""
    from multimodal_classification_b200 import ops
    m, n, k = B * T, 3072, 768
    a = torch.randn(m, k, device="cuda").to(torch.bfloat16)
    w = torch.randn(n, k, device="cuda").to(torch.bfloat16)
    bias = torch.randn(n, device="cuda")
    out = torch.empty(m, n, device="cuda", dtype=torch.bfloat16)
    pre = torch.empty_like(out)
    for _ in range(5):
        ops.gemm(a, w, out, bias=bias, act=ops.ACT_GELU, preact=pre, b_streamed=True)
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    iters = 50
    torch.cuda.synchronize()
    s.record()
    for _ in range(iters):
        ops.gemm(a, w, out, bias=bias, act=ops.ACT_GELU, preact=pre, b_streamed=True)
    e.record()
    torch.cuda.synchronize()
    dt = s.elapsed_time(e) / iters * 1e-3
    return {"kernel": "gemm_bf16_kernel (cta_group::2 pairs, multicast clusters): text FFN-1 2048x3072x768 +bias+GELU+preact",
            "tflops": 2.0 * m * n * k / dt / 1e12, "us": dt * 1e6,
            "frac_of_burst_peak": 2.0 * m * n * k / dt / 1e12 / peaks["bf16_burst"]}


def time_ingest(torch, dev, step, timed, steps, **loader_kw):
    """Train samples/s with every batch coming through multimodal_classification_b200.ingest.FeatureStoreLoader from host
    records (lmdb_dataset.py-shaped pickles held in a dict standing in for detectron.lmdb), loss read back every step."""
    import pickle
    import numpy as np
    import pandas as pd
    from multimodal_classification_b200 import ingest
    rng = np.random.default_rng(1234)
    n_rec = 64
    store = {}
    for i in range(n_rec):
        x1, y1 = rng.uniform(0, 700, R), rng.uniform(0, 700, R)
        boxes = np.stack([x1, y1, x1 + rng.uniform(50, 300, R), y1 + rng.uniform(50, 300, R)], 1).astype(np.float32)
        store[str(i).encode()] = pickle.dumps({"features": np.abs(rng.standard_normal((R, 2048))).astype(np.float32),
                                               "boxes": boxes}, protocol=4)

    class SyntheticTokenizer:      # no vocabulary file offline: seeded ids with ragged lengths, padded like the real one
        def __call__(self, texts, max_length, **kw):
            n = len(texts)
            lens = rng.integers(8, max_length + 1, n)
            mask = (np.arange(max_length)[None, :] < lens[:, None]).astype(np.int64)
            return {"input_ids": rng.integers(1, 30522, (n, max_length)) * mask, "attention_mask": mask,
                    "token_type_ids": np.zeros((n, max_length), np.int64)}

    rows = (steps + 6) * B
    df = pd.DataFrame({"id": [i % n_rec for i in range(rows)], "text": ["synthetic"] * rows,
                       "label": rng.integers(0, 2, rows).tolist()})
    loader = ingest.FeatureStoreLoader(df, ingest.LMDBRecords(store.get, R, 2048), SyntheticTokenizer(), T, B, drop_last=True,
                                       device=dev, **loader_kw)
    # producer alone: decode + pack + H2D + unpack, no model
    t0 = time.perf_counter()
    n = 0
    for batch in loader:
        n += batch["labels"].shape[0]
    torch.cuda.synchronize()
    loader_only = n / (time.perf_counter() - t0)
    it = iter(loader)

    def ingest_step():
        return step(next(it)).item()
    for _ in range(3):
        ingest_step()
    ms = timed(ingest_step, steps)
    blob = loader._layout(B).nbytes
    del it
    out = {"value": B / (ms / steps * 1e-3), "unit": UNIT, "ms_per_step": ms / steps, "steps": steps,
           "h2d_bytes_per_step": int(blob), "d2h_bytes_per_step": 4, "loader_only_samples_per_sec": loader_only,
           "source": f"{n_rec} pickled records ({R}x2048 fp32 features + {R}x4 boxes) in host memory, 1 producer thread, ring of 3"}
    try:      # beside it: the reference's loader shape (per-sample Dataset + default collate + pin + copy), oracle port, one thread
        from oracle import ingest_oracle as io
        tok, ids = SyntheticTokenizer(), [str(i % n_rec) for i in range(rows)]

        class PerSample:
            def __call__(self, text, max_length, **kw):
                e = tok([text], max_length)
                return {k: v[0] for k, v in e.items()}
        per_sample, done, t0 = PerSample(), 0, time.perf_counter()
        while time.perf_counter() - t0 < 2.0:
            samples = [io.lmdb_sample(ids[(done + j) % rows], "synthetic", 0, store.get, per_sample, T, R, 2048) for j in range(B)]
            batch = {k: torch.from_numpy(v).pin_memory().to(dev, non_blocking=True) for k, v in io.collate(samples).items()}
            done += B
        torch.cuda.synchronize()
        del batch
        out["reference_style_loader_samples_per_sec"] = done / (time.perf_counter() - t0)
        out["reference_style_loader"] = "oracle port of LMDBFeaturesDataset.__getitem__ + default collate + pin + H2D, 1 thread, same records"
    except Exception as e:
        out["reference_style_loader_error"] = repr(e)[:160]
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--eval-dropout-off", action="store_true", help="time with dropout off (parity configuration)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    if args.impl == "reference":
        run_reference_arm(args)
        return

    import torch
    import torch.distributed as dist
    from multimodal_classification_b200 import _lib
    from multimodal_classification_b200.vilbert import ViLBERTForClassification, get_facebook_vilbert_config
    from multimodal_classification_b200 import ddp as vb_ddp
    from oracle import vilbert_oracle as vo   # only for the synthetic batch generator and the cpu_baseline leg

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback for the product path)")
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    peaks = measured_peaks()

    cfg = get_facebook_vilbert_config()
    torch.manual_seed(0)
    model = ViLBERTForClassification(cfg, num_labels=2).to(dev)
    model.train(not args.eval_dropout_off)
    if world > 1 and os.environ.get("VB_DDP_SKIP", "0") != "1":      # VB_DDP_SKIP: diagnostic only (replicas without exchange)
        if os.environ.get("VB_DDP_FLUSH_MB"):
            vb_ddp.FLUSH_BYTES = int(os.environ["VB_DDP_FLUSH_MB"]) << 20
        vb_ddp.attach(model, dist.group.WORLD, compress=None if os.environ.get("VB_DDP_FP32", "0") == "1" else "bf16")
    host = vo.synthetic_batch(cfg, batch=B, seq=T, regions=R, seed=1234 + rank)
    pinned = {k: v.pin_memory() for k, v in host.items()}
    resident = {k: v.to(dev) for k, v in host.items()}
    h2d = sum(v.numel() * v.element_size() for v in host.values())

    params = list(model.parameters())

    def step(batch):
        for p in params:            # what optimizer.zero_grad() does in the reference loop (nodes.py:784); Module.zero_grad
            p.grad = None           # re-walks the module tree every call and costs 0.7 ms of host time here
        if model._engine is not None:
            model._engine.flat._version = -1      # weights change every real training step: refresh the bf16 shadows
        out = model(**batch)
        out["loss"].backward()
        return out["loss"]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(steps):
            fn()
        e.record()
        barrier()
        ms = torch.tensor([s.elapsed_time(e)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return ms.item()

    # ---- resident-input throughput ("value")
    for _ in range(args.warmup):
        step(resident)
    torch.cuda.synchronize()
    lc0 = _lib.launch_count()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ms = timed(lambda: step(resident), args.steps)
    launches_direct = _lib.launch_count() - lc0
    eng = model._engine
    plan = next(iter(eng.plans.values()))
    launches = launches_direct // max(1, args.steps) + plan.fwd_launches + plan.bwd_launches   # kernels per step
    ms_per_step = ms / args.steps
    value = world * B / (ms_per_step * 1e-3)

    # ---- end to end through the public API with host buffers: H2D of the batch and D2H of the loss inside the region
    def e2e_step():
        dev_batch = {k: v.to(dev, non_blocking=True) for k, v in pinned.items()}
        loss = step(dev_batch)
        return loss.item()
    for _ in range(3):
        e2e_step()
    ms_e2e = timed(e2e_step, args.steps)
    clocks = sampler.stop() if rank == 0 else {}
    e2e_value = world * B / (ms_e2e / args.steps * 1e-3)

    # ---- the same step followed by the fused clip + AdamW + shadow-refresh pass, resident inputs
    from multimodal_classification_b200.optim import FusedAdamW
    opt = FusedAdamW(model, lr=1e-5, weight_decay=0.01, max_grad_norm=1.0)

    def opt_step():
        for p in params:
            p.grad = None
        out = model(**resident)
        out["loss"].backward()
        opt.step()
    for _ in range(3):
        opt_step()
    ms_opt = timed(opt_step, args.steps)

    # ---- the same step fed by the feature-store loader: records decoded from an in-memory LMDB image
    #      by the producer thread, one pinned blob + one H2D + one unpack launch per batch, overlapped with the previous step
    ingest = None
    if world == 1:
        try:
            ingest = time_ingest(torch, dev, step, timed, max(args.steps, 60))
        except Exception as e:      # the headline line must still print
            ingest = {"error": repr(e)[:200]}

    if rank == 0:
        step_tflops = FLOP_PER_SAMPLE_FWD_BWD * value / world / 1e12      # per GPU
        dom = gemm_kernel_roofline(torch, peaks)
        # dominant kernel (the GEMM family is ~70 % of the step): algorithmic FLOPs of one launch / its CUDA-event time, against
        # the measured BURST bf16 peak (a kernel timed alone); `traffic` = dram__bytes_read.sum + dram__bytes_write.sum of that
        # launch from the ncu --set full capture in profiles/r01c_dominant_gemm_full.md (inputs 7.87 MB; the 25 MB of outputs
        # stay in L2 for the next kernel).  The whole step against the SUSTAINED peak is reported beside it.
        roofline = {"bound": "tensor", "achieved": dom["tflops"], "peak": peaks["bf16_burst"], "unit": "TFLOP/s",
                    "frac": dom["tflops"] / peaks["bf16_burst"], "traffic": 7.94e6, "traffic_unit": "bytes per launch (DRAM)",
                    "kernel": dom["kernel"], "us_per_launch": dom["us"], "flop_per_launch": 2.0 * 2048 * 3072 * 768,
                    "peak_source": peaks["source"],
                    "step": {"achieved": step_tflops, "peak": peaks["bf16_sustained"], "frac": step_tflops / peaks["bf16_sustained"],
                             "unit": "TFLOP/s", "scope": "whole step: %d samples x 152.07 GFLOP per graph replay, sustained bf16 peak" % B}}
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "bf16", "data": "synthetic",
                "config": {"workload": "vilbert_base_train_step_bs16_t128_r100", "per_gpu_batch": B, "global_batch": B * world,
                           "tokens": T, "regions": R, "dropout": bool(model.training), "step": "fwd+bwd (+bf16 weight-shadow refresh)",
                           "parallelism": f"dp{world}", "l2": "working set ~2 GB/step (weights+grads) > 126 MB L2, no flush",
                           "cuda_graphs": eng.use_graphs},
                "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4,
                        "ms_per_step": ms_e2e / args.steps},
                "with_fused_optimizer": {"value": world * B / (ms_opt / args.steps * 1e-3), "unit": UNIT, "ms_per_step": ms_opt / args.steps,
                                         "step": "fwd+bwd + fused clip/AdamW/bf16-shadow pass (2 launches)"},
                "gpu_launches": int(launches) * args.steps, "gpu_launches_per_step": int(launches), "roofline": roofline, "clocks": clocks}
        if ingest is not None:
            line["ingest"] = ingest
        if world == 1 and not args.no_cpu_baseline:
            v, cores, n = time_cpu_port(25.0, B)
            line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                                    "sample": f"median of {n} fwd+bwd passes of the oracle port over the same bs={B} batch, eval mode, fp32"}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        vb_ddp.shutdown(model)
        sys.stdout.flush()
        os._exit(0)     # graphs that captured NCCL collectives can wedge interpreter teardown; the run is complete


if __name__ == "__main__":
    main()
