import sys, torch
sys.path.insert(0, ".")
from multimodal_classification_b200 import ops
for b in (1, 16):
    img = torch.randn(b, 3, 600, 600, device="cuda")
    col = torch.empty(b * 300 * 300, 152, dtype=torch.bfloat16, device="cuda")
    w = torch.randn(64, 152, device="cuda").to(torch.bfloat16)
    out = torch.empty(b * 300 * 300, 64, dtype=torch.bfloat16, device="cuda")
    pooled = torch.empty(b, 150, 150, 64, dtype=torch.bfloat16, device="cuda")
    for name, fn in (("stem_im2col", lambda: ops.stem_im2col(img, col)), ("stem_gemm", lambda: ops.gemm(col, w, out, act=ops.ACT_RELU)),
                     ("maxpool", lambda: ops.maxpool_nhwc(out.view(b, 300, 300, 64), pooled, 3, 2, 1))):
        for _ in range(3): fn()
        torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(20): fn()
        e.record(); torch.cuda.synchronize()
        us = s.elapsed_time(e) / 20 * 1e3
        print(f"batch {b} {name}: {us:.1f} us  ({col.numel() * 2 / us / 1e6:.2f} TB/s of col bytes)")
