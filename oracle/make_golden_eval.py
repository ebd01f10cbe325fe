"""Generate tests/golden/vilbert_eval512.npz: logits of the UNMODIFIED reference ViLBERTForClassification (eval mode, fp32, CPU)
on a fixed 512-sample synthetic eval set (8 seeded batches of 64, full configuration, seeded weights) for the AUROC-ordering
check the north-star asks for.  Authoring container only (/root/reference does not travel)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference/src")
from oracle import vilbert_oracle as vo  # noqa: E402

SEEDS = list(range(9000, 9008))


def main():
    from multimodalclassification.models.vilbert_facebook_arch import ViLBERTForClassification
    torch.set_num_threads(os.cpu_count() or 1)
    cfg = vo.facebook_config()
    sd = vo.seeded_state_dict(cfg)
    torch.manual_seed(0)
    m = ViLBERTForClassification(cfg, num_labels=2)
    m.load_state_dict(sd, strict=True)
    m.eval()
    logits, labels, autocast = [], [], []
    with torch.no_grad():
        for s in SEEDS:
            b = vo.synthetic_batch(cfg, batch=64, seq=128, regions=100, seed=s)
            logits.append(m(**b)["logits"].numpy())
            labels.append(b["labels"].numpy())
            # yardstick: the reference itself under PyTorch's CPU bf16 autocast on the same samples (what bf16 arithmetic costs
            # on this model; the GPU test's bar is relative to it, as for the gradients in make_golden.py)
            with torch.autocast("cpu", dtype=torch.bfloat16):
                autocast.append(m(**b)["logits"].float().numpy())
            print("seed", s, "done", flush=True)
    path = os.path.join(ROOT, "tests", "golden", "vilbert_eval512.npz")
    ref, ac = np.concatenate(logits).astype(np.float32), np.concatenate(autocast).astype(np.float32)
    yard = float(np.abs(ac - ref).max() / np.abs(ref).max())
    print("autocast yardstick max|dlogit|/max|logit| over 512 samples:", yard)
    np.savez_compressed(path, logits=ref, labels=np.concatenate(labels), seeds=np.asarray(SEEDS), autocast_logits=ac,
                        yard_logits=np.asarray(yard))
    print("wrote", path)


if __name__ == "__main__":
    main()
