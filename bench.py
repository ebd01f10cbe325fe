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
# DRAM bytes of one launch of the dominant GEMM shape, from the committed ncu --set full capture
TRAFFIC_BYTES = 7.94e6
TRAFFIC_SOURCE = "profiles/r01c_dominant_gemm_full.md (text FFN-1 2048x3072x768: 7.94 MB read = its 7.87 MB of inputs, ~1 KB written; outputs stay in L2)"
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
    # the very first pass pays thread-pool start-up and the first touch of 1 GB of weights: warm up BEFORE probing, or a cold
    # probe halves the batch for no reason (round 1: the N=1 scaling run fell to an 8-sample slice)
    cpu_port_step(cfg, sd, vo.synthetic_batch(cfg, batch=2, seq=T, regions=R, seed=1))
    while True:
        batch = vo.synthetic_batch(cfg, batch=bs, seq=T, regions=R, seed=1234)
        cpu_port_step(cfg, sd, batch)
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
            "config": {"workload": "vilbert_base_train_step_bs16_t128_r100", "per_gpu_batch": bs, "global_batch": bs, "tokens": T,
                       "regions": R, "dropout": False, "step": "fwd+bwd (fp32 autograd on the host cores)", "parallelism": "cpu",
                       "grad_exchange": "none", "l2": "n/a (CPU)", "cuda_graphs": False},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def graph_time(torch, fn, rep=20, iters=10):
    """us per launch of fn(): `rep` back-to-back launches captured in a CUDA graph (as in the graph-replayed step: no ctypes /
    launch overhead), replayed `iters` times between two CUDA events on the replay stream."""
    fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(rep):
            fn()
    g.replay()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(iters):
        g.replay()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / (iters * rep) * 1e3


# every GEMM site of the bs-16 step: (name, M, N, K, launches per step in each direction, GELU epilogue) -- SURVEY.md Appendix B
GEMM_SITES = [("t.qkv", 2048, 2304, 768, 12, 0), ("t.attn_out", 2048, 768, 768, 12, 0), ("t.ffn1", 2048, 3072, 768, 18, 1),
              ("t.ffn2", 2048, 768, 3072, 18, 0), ("v.qkv", 1600, 3072, 1024, 12, 0), ("v.1024", 1600, 1024, 1024, 18, 0),
              ("v.ffn1", 1600, 1024, 1024, 12, 1), ("c.tqkv", 2048, 3072, 768, 6, 0), ("c.dense2", 2048, 768, 1024, 6, 0),
              ("img_emb", 1600, 1024, 2048, 1, 0)]


def gemm_family_roofline(torch, peaks, wgrad_ctas):
    """The dominant kernel FAMILY, honestly: every (site, direction) GEMM of the step is timed live (CUDA events over graph
    replays, launched with the engine's own epilogue flags), and the family figure is sum(2MNK x launches) / sum(time x
    launches) -- not the best shape.  Weight gradients are timed twice: chip-wide and with the engine's SM cap (they run on
    side streams under that cap, in the shadow of the dgrad chain)."""
    from multimodal_classification_b200 import ops
    bf = torch.bfloat16
    gelu_grad_fwd = os.environ.get("VB_GELU_GRAD_FWD", "1") != "0"      # as vilbert._Engine

    def rnd(*shape):
        return (torch.randn(*shape, device="cuda") * 0.5).to(bf)
    rows, fl_sum, t_sum, t_sum_capped = [], 0.0, 0.0, 0.0
    for name, m, n, k, cnt, gelu in GEMM_SITES:
        x, w, bias = rnd(m, k), rnd(n, k), torch.randn(n, device="cuda")
        y, pre, dy, dx, aux = (torch.empty(m, n, device="cuda", dtype=bf), torch.empty(m, n, device="cuda", dtype=bf), rnd(m, n),
                               torch.empty(m, k, device="cuda", dtype=bf), rnd(m, k))
        dw = torch.zeros(n, k, device="cuda")
        if gelu:
            t_f = graph_time(torch, lambda: ops.gemm(x, w, y, bias=bias, act=ops.ACT_GELU, preact=pre, b_streamed=True,
                                                     preact_grad=gelu_grad_fwd))
        else:
            t_f = graph_time(torch, lambda: ops.gemm(x, w, y, bias=bias, b_streamed=True))
        # the dgrad of FFN-2 carries gelu' of FFN-1: stored by the forward (the engine's default) or recomputed from the pre-activation
        mode = (ops.AUX_MUL if gelu_grad_fwd else ops.AUX_MUL_GELU_GRAD) if name in ("t.ffn2",) else ops.AUX_ADD
        t_d = graph_time(torch, lambda: ops.gemm(dy, w, dx, b_mn_major=True, aux=aux, aux_mode=mode, b_streamed=True))
        t_w = graph_time(torch, lambda: ops.gemm(dy, x, dw, a_mn_major=True, b_mn_major=True, d_streamed=True))
        t_wc = graph_time(torch, lambda: ops.gemm(dy, x, dw, a_mn_major=True, b_mn_major=True, d_streamed=True, max_ctas=wgrad_ctas))
        fl = 2.0 * m * n * k
        for d, t in (("fwd", t_f), ("dgrad", t_d), ("wgrad", t_w)):
            rows.append({"site": name, "dir": d, "m": m, "n": n, "k": k, "launches_per_step": cnt, "us": t, "tflops": fl / t / 1e6})
        fl_sum += 3 * cnt * fl
        t_sum += cnt * (t_f + t_d + t_w)
        t_sum_capped += cnt * (t_f + t_d + t_wc)
    best = max(rows, key=lambda r: r["tflops"])
    worst = min(rows, key=lambda r: r["tflops"])
    dominant = max(rows, key=lambda r: r["us"] * r["launches_per_step"])
    return {"family_tflops": fl_sum / t_sum / 1e6, "family_tflops_wgrad_capped": fl_sum / t_sum_capped / 1e6,
            "serial_us_per_step": t_sum, "serial_us_per_step_wgrad_capped": t_sum_capped, "flop_per_step": fl_sum,
            "best": best, "worst": worst, "dominant": dominant, "rows": rows}


def time_roi_stage(torch, peaks):
    """BASELINE.json configs[2] (driver-visible): the ResNet-152 C5 + RoI feature stage, images/s at the reference's own shape
    (600x600, RoIPool 14x14, resnet152_roi.py:126-133) and at the BASELINE shape (448x448, RoIAlign 7x7), batch 1 and 16, plus
    the RoI pooling kernels alone against the measured HBM peak.  FLOPs per image: 214.9 / 104.1 GF."""
    from multimodal_classification_b200 import ops
    from multimodal_classification_b200.resnet152_roi import ResNet152ROIExtractor
    out = {"legs": []}
    for size, roi, mode, gf in ((600, 14, "roi_pool", 214.9), (448, 7, "roi_align", 104.1)):
        torch.manual_seed(0)
        ext = ResNet152ROIExtractor(device="cuda", weights=None, roi_size=roi, image_size=size, pool_mode=mode)
        for b in (1, 16):
            imgs = torch.randn(b, 3, size, size, device="cuda")
            for _ in range(3):
                ext.extract_batch(imgs)
            torch.cuda.synchronize()
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            iters = 10
            s.record()
            for _ in range(iters):
                ext.extract_batch(imgs)
            e.record()
            torch.cuda.synchronize()
            ms = s.elapsed_time(e) / iters
            out["legs"].append({"image": size, "pool": f"{mode}-{roi}", "batch": b, "ms": ms, "images_per_s": b / ms * 1e3,
                                "tflops": gf * b / ms, "frac_of_burst_bf16": gf * b / ms / peaks["bf16_burst"]})
        del ext
    # the Visual Genome Faster R-CNN region extractors (row f-4) on the same kernels, random-init weights, their scored branches
    # switched on as a checkpoint would: ms per picture of the whole CUDA graph (trunk + proposals + scoring + selection)
    try:
        from multimodal_classification_b200.fasterrcnn_vg import FasterRCNNVGExtractor
        from multimodal_classification_b200.fasterrcnn_vg_rpn import FasterRCNNVGRPNExtractor
        torch.manual_seed(0)
        vg = FasterRCNNVGExtractor(weights_path="/nonexistent/vg.pth", device="cuda", weights=None)
        vg.has_vg_weights = True
        x = torch.randn(1, 3, 600, 1000, device="cuda")
        rpn = FasterRCNNVGRPNExtractor(weights_path="/nonexistent/vg.pth", device="cuda", weights=None)
        y = torch.randn(1, 3, 600, 800, device="cuda")
        for name, fn in (("fasterrcnn_vg 600x1000, 200 scored candidates -> 36 regions", lambda: vg.extract_batch(x)),
                         ("fasterrcnn_vg_rpn 600x800, 22800 anchors -> NMS -> 300 scored -> 36 regions",
                          lambda: rpn.extract_preprocessed(y, 6.25, 128, 96))):
            for _ in range(3):
                fn()
            torch.cuda.synchronize()
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            for _ in range(10):
                fn()
            e.record()
            torch.cuda.synchronize()
            out.setdefault("vg_extractors", []).append({"extractor": name, "ms_per_picture": s.elapsed_time(e) / 10})
        del vg, rpn
    except Exception as err:      # the headline line must still print
        out["vg_extractors"] = {"error": repr(err)[:200]}
    # the pooling kernels alone: 36 boxes on one C4 map (stride 16, 1024 channels), bytes = map read once + output written once
    for size, roi, mode in ((600, 14, "roi_pool"), (448, 7, "roi_align")):
        hw = (size + 15) // 16
        x = (torch.randn(1, hw, hw, 1024, device="cuda")).to(torch.bfloat16)
        g = torch.Generator(device="cpu").manual_seed(0)
        xy = torch.rand(36, 2, generator=g) * size * 0.5
        wh = torch.rand(36, 2, generator=g) * size * 0.45 + 32
        rois = torch.cat([torch.zeros(36, 1), xy, xy + wh], 1).clamp(max=size - 1).cuda().contiguous()
        o = torch.empty(36, roi, roi, 1024, device="cuda", dtype=torch.bfloat16)
        if mode == "roi_pool":
            us = graph_time(torch, lambda: ops.roi_pool_nhwc(x, rois, o, 1.0 / 16))
        else:
            us = graph_time(torch, lambda: ops.roi_align_nhwc(x, rois, o, 1.0 / 16, 2, False))
        nbytes = x.numel() * 2 + o.numel() * 2
        out.setdefault("kernels", []).append({"kernel": f"{mode}_nhwc_kernel", "image": size, "us": us, "bytes": nbytes,
                                              "gbs": nbytes / us / 1e3, "frac_of_hbm_peak": nbytes / us / 1e3 / peaks["hbm"]})
    return out


def time_inference_sweep(torch, model, cfg, vo, dev, sizes=(16, 512)):
    """BASELINE.json configs[4], second half: eval-mode forward throughput (no gradient state kept), resident inputs."""
    rows = []
    was_training = model.training
    model.eval()
    for bs in sizes:
        batch = {k: v.to(dev) for k, v in vo.synthetic_batch(cfg, batch=bs, seq=T, regions=R, seed=1234).items() if k != "labels"}
        with torch.no_grad():
            for _ in range(3):
                out = model(**batch)
            torch.cuda.synchronize()
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            iters = 10
            s.record()
            for _ in range(iters):
                out = model(**batch)
            e.record()
            torch.cuda.synchronize()
        ms = s.elapsed_time(e) / iters
        rows.append({"batch": bs, "ms": ms, "samples_per_s": bs / ms * 1e3, "tflops": 50.83 * bs / ms,
                     "finite": bool(torch.isfinite(out["logits"]).all())})
        model._engine.plans.clear()      # the bs-512 plan holds ~28 GB of activations
        torch.cuda.empty_cache()
    model.train(was_training)
    return rows


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
    ap.add_argument("--core-only", action="store_true", help="headline legs only (profiling runs): no stock-optimizer / ingest / RoI / inference legs")
    ap.add_argument("--steps-only", action="store_true", help="warm-up + K resident steps and nothing else (the command ncu wraps for the launch list)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    if args.core_only:
        os.environ["VB_BENCH_SKIP_EXTRA"] = os.environ["VB_BENCH_SKIP_STOCK"] = os.environ["VB_BENCH_SKIP_INGEST"] = "1"
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
    grad_exchange = "none (1 GPU)"
    model = ViLBERTForClassification(cfg, num_labels=2).to(dev)
    model.train(not args.eval_dropout_off)
    if world > 1 and os.environ.get("VB_DDP_SKIP", "0") != "1":      # VB_DDP_SKIP: diagnostic only (replicas without exchange)
        if os.environ.get("VB_DDP_FLUSH_MB"):
            vb_ddp.FLUSH_BYTES = vb_ddp.SWITCH_FLUSH_BYTES = int(os.environ["VB_DDP_FLUSH_MB"]) << 20
        # gradient exchange of the scaling runs: bf16 through our own NVSwitch kernels (multimem reduce in the switch; the weight-
        # gradient GEMMs write bf16 straight into the symmetric buffer).  VB_DDP_TRANSPORT=nccl: bucketed NCCL all-reduce of bf16
        # casts; VB_DDP_FP32=1: fp32 over NCCL (numerically the single-GPU step).  A box without symmetric-memory support falls
        # back to NCCL and says so in config.grad_exchange.
        transport = os.environ.get("VB_DDP_TRANSPORT", "switch")
        if os.environ.get("VB_DDP_FP32", "0") == "1":
            grad_exchange = "fp32 (NCCL all-reduce)"
            vb_ddp.attach(model, dist.group.WORLD, compress=None)
        elif transport == "switch" and vb_ddp.switch_available(dist.group.WORLD, dev):
            grad_exchange = "bf16 (own NVSwitch kernels: multimem.ld_reduce / multimem.st)"
            vb_ddp.attach(model, dist.group.WORLD, compress="bf16", transport="switch")
        else:
            grad_exchange = "bf16 (NCCL all-reduce)" + (" [switch transport unavailable on this box]" if transport == "switch" else "")
            vb_ddp.attach(model, dist.group.WORLD, compress="bf16")
    host = vo.synthetic_batch(cfg, batch=B, seq=T, regions=R, seed=1234 + rank)
    pinned = {k: v.pin_memory() for k, v in host.items()}
    resident = {k: v.to(dev) for k, v in host.items()}
    h2d = sum(v.numel() * v.element_size() for v in host.values())

    params = list(model.parameters())

    def step(batch, mark=True):
        for p in params:            # what optimizer.zero_grad() does in the reference loop (nodes.py:784); Module.zero_grad
            p.grad = None           # re-walks the module tree every call and costs 0.7 ms of host time here
        if mark:
            model.parameters_updated()            # weights change every real training step: refresh the bf16 shadows
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

    if args.steps_only:
        if rank == 0:
            print(json.dumps({"metric": METRIC, "value": value, "unit": UNIT, "ms_per_step": ms_per_step, "steps": args.steps,
                              "warmup": args.warmup, "gpu_launches_per_step": int(launches), "note": "steps only (profiling run)"}), flush=True)
        return

    # ---- end to end through the public API with host buffers: H2D of the batch and D2H of the loss inside the region
    def e2e_step():
        # "the optimizer has stepped" is marked where optimizer.step() sits in the reference's loop (nodes.py:795-799): before
        # the next batch is copied, so the shadow refresh may run beside that copy (torch optimizers mark it by a hook)
        model.parameters_updated()
        dev_batch = {k: v.to(dev, non_blocking=True) for k, v in pinned.items()}
        loss = step(dev_batch, mark=False)
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

    # ---- the reference's own loop around the same step (nodes.py:757-760, 784-799): stock AdamW over the 523 parameter
    #      views, clip_grad_norm_, LambdaLR warm-up schedule, loss read back -- on a second replica (own optimizer state)
    ms_stock = None
    if world == 1 and os.environ.get("VB_BENCH_SKIP_STOCK", "0") != "1":
        from transformers import get_linear_schedule_with_warmup
        stock = torch.optim.AdamW(model.parameters(), lr=1e-5, weight_decay=0.01, eps=1e-8)
        sched = get_linear_schedule_with_warmup(stock, 10, 10 ** 6)

        def stock_step():
            stock.zero_grad()
            out = model(**resident)
            loss = out["loss"]
            loss.backward()
            torch.nn.utils.clip_grad_norm_(model.parameters(), 1.0)
            stock.step()
            sched.step()
            return loss.item()
        for _ in range(3):
            stock_step()
        n_stock = max(5, args.steps // 2)
        ms_stock = timed(stock_step, n_stock) / n_stock
        del stock, sched

    # ---- the same step fed by the feature-store loader: records decoded from an in-memory LMDB image
    #      by the producer thread, one pinned blob + one H2D + one unpack launch per batch, overlapped with the previous step
    ingest = None
    if world == 1 and os.environ.get("VB_BENCH_SKIP_INGEST", "0") != "1":
        try:
            ingest = time_ingest(torch, dev, step, timed, max(args.steps, 60))
        except Exception as e:      # the headline line must still print
            ingest = {"error": repr(e)[:200]}

    if rank == 0:
        step_tflops = FLOP_PER_SAMPLE_FWD_BWD * value / world / 1e12      # per GPU
        fam = gemm_family_roofline(torch, peaks, eng.wgrad_ctas)
        # Dominant kernel FAMILY (gemm_bf16_kernel: ~70 % of the step's kernel time): `achieved` is the launch-weighted figure over
        # all 30 (site, direction) GEMMs of the step, each timed live with CUDA events; best / worst / dominant single shapes
        # beside it.  Peak = the measured BURST bf16 figure (kernels timed alone; the 0.1 s timed region of the step runs at
        # full clocks too).  `traffic` = dram__bytes_read.sum + dram__bytes_write.sum of one launch of the dominant shape from
        # the ncu --set full capture named in `traffic_source`.
        roofline = {"bound": "tensor", "achieved": fam["family_tflops"], "peak": peaks["bf16_burst"], "unit": "TFLOP/s",
                    "frac": fam["family_tflops"] / peaks["bf16_burst"],
                    "kernel": "gemm_bf16_kernel family (tcgen05 cta_group::2 pairs, TMA multicast clusters): 30 (site, direction) shapes of the step, launch-weighted",
                    "flop_per_step_family": fam["flop_per_step"], "serial_us_per_step": fam["serial_us_per_step"],
                    "as_scheduled": {"achieved": fam["family_tflops_wgrad_capped"], "frac": fam["family_tflops_wgrad_capped"] / peaks["bf16_burst"],
                                     "serial_us_per_step": fam["serial_us_per_step_wgrad_capped"],
                                     "note": f"weight-gradient GEMMs capped to {eng.wgrad_ctas} CTAs as the engine launches them (side streams)"},
                    "best": {k: fam["best"][k] for k in ("site", "dir", "us", "tflops")},
                    "worst": {k: fam["worst"][k] for k in ("site", "dir", "us", "tflops")},
                    "dominant": {k: fam["dominant"][k] for k in ("site", "dir", "us", "tflops", "launches_per_step")},
                    "traffic": TRAFFIC_BYTES, "traffic_unit": "bytes per launch (DRAM) of the dominant shape", "traffic_source": TRAFFIC_SOURCE,
                    "peak_source": peaks["source"],
                    "step": {"achieved": step_tflops, "peak": peaks["bf16_burst"], "frac": step_tflops / peaks["bf16_burst"],
                             "unit": "TFLOP/s", "scope": "whole step: %d samples x 152.07 GFLOP per graph replay, BURST bf16 peak (0.1 s region at full clocks)" % B,
                             "frac_of_sustained": step_tflops / peaks["bf16_sustained"]},
                    "per_shape": [{k: (round(r[k], 2) if isinstance(r[k], float) else r[k]) for k in ("site", "dir", "us", "tflops")} for r in fam["rows"]]}
        roi = inference = None
        if world == 1 and os.environ.get("VB_BENCH_SKIP_EXTRA", "0") != "1":
            try:
                roi = time_roi_stage(torch, peaks)
            except Exception as e:
                roi = {"error": repr(e)[:200]}
            try:
                inference = time_inference_sweep(torch, model, cfg, vo, dev)
            except Exception as e:
                inference = {"error": repr(e)[:200]}
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "bf16", "data": "synthetic",
                "config": {"workload": "vilbert_base_train_step_bs16_t128_r100", "per_gpu_batch": B, "global_batch": B * world,
                           "tokens": T, "regions": R, "dropout": bool(model.training), "step": "fwd+bwd (+bf16 weight-shadow refresh)",
                           "parallelism": f"dp{world}", "grad_exchange": grad_exchange, "l2": "working set ~2 GB/step (weights+grads) > 126 MB L2, no flush",
                           "cuda_graphs": eng.use_graphs},
                "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4,
                        "ms_per_step": ms_e2e / args.steps},
                "with_fused_optimizer": {"value": world * B / (ms_opt / args.steps * 1e-3), "unit": UNIT, "ms_per_step": ms_opt / args.steps,
                                         "step": "fwd+bwd + fused clip/AdamW/bf16-shadow pass (2 launches)"},
                "gpu_launches": int(launches) * args.steps, "gpu_launches_per_step": int(launches), "roofline": roofline, "clocks": clocks}
        if ms_stock is not None:
            line["with_stock_optimizer"] = {"value": B / (ms_stock * 1e-3), "unit": UNIT, "ms_per_step": ms_stock,
                                            "step": "the reference's loop (nodes.py:784-799): zero_grad, fwd, bwd, clip_grad_norm_, torch.optim.AdamW.step, LambdaLR.step, loss.item()"}
        if ingest is not None:
            line["ingest"] = ingest
        if roi is not None:
            line["roi_stage"] = roi
        if inference is not None:
            line["inference"] = inference
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
