"""The RoI feature stage alone, for ncu (launch list / full captures): reference shape 600 x 600 / RoIPool-14 (default) or the
BASELINE shape 448 x 448 / RoIAlign-7 (--align), one eager pass + graph capture + `--replays` replays.
    python tools/prof_roi.py [--batch 16] [--align] [--replays 1]"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=16)
    ap.add_argument("--align", action="store_true")
    ap.add_argument("--replays", type=int, default=1)
    args = ap.parse_args()
    from multimodal_classification_b200.resnet152_roi import ResNet152ROIExtractor
    from oracle import roi_oracle as ro
    size, roi, mode = (448, 7, "roi_align") if args.align else (600, 14, "roi_pool")
    ext = ResNet152ROIExtractor(device="cuda", weights=None, roi_size=roi, image_size=size, pool_mode=mode)
    ext.backbone.load_state_dict(ro.seeded_backbone_state(0))
    imgs = torch.randn(args.batch, 3, size, size, device="cuda")
    ext.extract_batch(imgs)                     # eager + capture
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.profiler.start()                 # ncu --profile-from-start off: only the replays are profiled
    s.record()
    for _ in range(args.replays):
        feats, _ = ext.extract_batch(imgs)
    e.record()
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
    print(f"{size}x{size} {mode}-{roi} batch {args.batch}: {s.elapsed_time(e) / args.replays:.3f} ms per replay, finite {bool(torch.isfinite(feats).all())}")


if __name__ == "__main__":
    main()
