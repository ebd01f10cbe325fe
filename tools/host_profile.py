"""cProfile of the host side of the end-to-end training step (pinned batch -> H2D -> forward -> backward -> loss.item()):
where the Python time between two graph launches goes.  python tools/host_profile.py [steps]"""
import cProfile
import os
import pstats
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402


def main():
    from multimodal_classification_b200.vilbert import ViLBERTForClassification, get_facebook_vilbert_config
    from oracle import vilbert_oracle as vo
    steps = int(sys.argv[1]) if len(sys.argv) > 1 else 100
    cfg = get_facebook_vilbert_config()
    torch.manual_seed(0)
    dev = torch.device("cuda")
    model = ViLBERTForClassification(cfg, num_labels=2).to(dev).train()
    host = vo.synthetic_batch(cfg, batch=16, seq=128, regions=100, seed=1234)
    pinned = {k: v.pin_memory() for k, v in host.items()}
    params = list(model.parameters())

    def step():
        model.parameters_updated()
        batch = {k: v.to(dev, non_blocking=True) for k, v in pinned.items()}
        for p in params:
            p.grad = None
        out = model(**batch)
        out["loss"].backward()
        return out["loss"].item()
    for _ in range(5):
        step()
    torch.cuda.synchronize()
    prof = cProfile.Profile()
    prof.enable()
    for _ in range(steps):
        step()
    prof.disable()
    st = pstats.Stats(prof)
    st.sort_stats("tottime").print_stats(28)
    st.sort_stats("cumulative").print_stats(28)


if __name__ == "__main__":
    main()
