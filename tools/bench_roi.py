"""RoI feature stage throughput (BASELINE.json configs[2]): ResNet-152 C5 trunk + RoI pooling over 36 boxes per image.

    python tools/bench_roi.py [--cpu]      # --cpu also times the oracle (fp32, host cores) on one image

Reference shape: 600x600, RoIPool 14x14 (resnet152_roi.py:126, 133); BASELINE shape: 448x448, RoIAlign 7x7.
FLOPs per image (measured with FlopCounterMode on the reference): 214.9 GF @600/RoIPool-14, 104.1 GF @448/RoI-7.
"""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

def main():
    ap = argparse.ArgumentParser(); ap.add_argument("--cpu", action="store_true"); ap.add_argument("--grid", action="store_true"); args = ap.parse_args()
    from multimodal_classification_b200.resnet152_roi import ResNet152ROIExtractor
    from oracle import roi_oracle as ro
    sd = ro.seeded_backbone_state(0)
    rows = []
    for size, roi, mode, gf in () if args.grid else ((600, 14, "roi_pool", 214.9), (448, 7, "roi_align", 104.1)):
        ext = ResNet152ROIExtractor(device="cuda", weights=None, roi_size=roi, image_size=size, pool_mode=mode)
        ext.backbone.load_state_dict(sd)
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
            e.record(); torch.cuda.synchronize()
            ms = s.elapsed_time(e) / iters
            rows.append({"image": size, "roi": f"{mode}-{roi}", "batch": b, "ms": ms, "images_per_s": b / ms * 1e3,
                         "tflops": gf * b / ms})
            print(json.dumps(rows[-1]), flush=True)
    if args.grid:       # the grid extractors: whole trunk at 224x224, 7x7 -> 6x6 average grid; 23.0 / 15.6 GF per image (FlopCounterMode)
        from multimodal_classification_b200.resnet_grid import ResNetFeatureExtractor, ResNetVGExtractor
        for name, make, state, gf in (("resnet152_grid", ResNetFeatureExtractor, ro.grid_backbone_state(sd), 23.0),
                                      ("resnet101_vg_grid", ResNetVGExtractor,
                                       ro.vg_backbone_state(ro.seeded_backbone_state(1, (3, 4, 23, 3))), 15.6)):
            ext = make(device="cuda", weights=None)
            ext.backbone.load_state_dict(state)
            for b in (1, 16, 64):
                imgs = torch.randn(b, 3, 224, 224, device="cuda")
                for _ in range(3):
                    ext.extract_batch(imgs)
                torch.cuda.synchronize()
                s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                s.record()
                for _ in range(10):
                    ext.extract_batch(imgs)
                e.record(); torch.cuda.synchronize()
                ms = s.elapsed_time(e) / 10
                print(json.dumps({"extractor": name, "batch": b, "ms": ms, "images_per_s": b / ms * 1e3, "tflops": gf * b / ms}), flush=True)
        return
    if args.cpu:
        torch.set_num_threads(os.cpu_count() or 1)
        img = torch.randn(1, 3, 600, 600)
        ro.extract_features(sd, img)
        t0 = time.time(); ro.extract_features(sd, img); dt = time.time() - t0
        print(json.dumps({"cpu_oracle_600_roipool14_s_per_image": dt, "cores": os.cpu_count()}))

if __name__ == "__main__":
    main()
