"""vb_lmdb_regions alone: CUDA-event time and achieved HBM GB/s against MEASURED_PEAKS.json, at the bench
batch (16 x 100 regions: 19.7 MB, L2-sized) and at batches whose traffic exceeds the 126 MB L2.  Algorithmic bytes per launch:
rows * 2048 * (4 read + 2 written) + rows * (16 read + 20 written).  Also the launch `ncu --set full -k regex:lmdb_regions`
captures (last size).    python tools/bench_ingest_kernel.py"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from multimodal_classification_b200 import ops  # noqa: E402

peak = 6556.8
try:
    peak = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass
F = 2048
for batch in (16, 128, 512):
    rows = batch * 100
    feat = torch.rand(rows, F, device="cuda")
    boxes = torch.rand(rows, 4, device="cuda") * 900
    out = torch.empty(rows, F, dtype=torch.bfloat16, device="cuda")
    spatial = torch.empty(rows, 5, device="cuda")
    nbytes = rows * F * 6 + rows * 36
    for _ in range(5):
        ops.lmdb_regions(feat, out, boxes, spatial)
    reps = 50
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    s.record()
    for _ in range(reps):
        ops.lmdb_regions(feat, out, boxes, spatial)
    e.record()
    torch.cuda.synchronize()
    us = s.elapsed_time(e) / reps * 1e3
    print("batch %4d  rows %6d  %7.1f MB  %8.2f us  %7.0f GB/s  %.1f %% of measured HBM peak %.0f GB/s"
          % (batch, rows, nbytes / 1e6, us, nbytes / us / 1e3, 100 * nbytes / us / 1e3 / peak, peak), flush=True)
