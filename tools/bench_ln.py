"""LayerNorm kernel dissection: forward, backward, backward without the column sums (dgamma/dbeta/dbias atomics)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from multimodal_classification_b200 import ops
from tools.bench_kernels import graph_time, rnd

seed = torch.tensor([1234], dtype=torch.int64, device="cuda")
for name, m, h in (("t.ln", 2048, 768), ("v.ln", 1600, 1024)):
    x, res, dy = rnd(m, h), rnd(m, h), rnd(m, h)
    g, b = torch.ones(h, device="cuda"), torch.zeros(h, device="cuda")
    y, dx, dres = torch.empty_like(x), torch.empty_like(x), torch.empty_like(x)
    mean, rstd = torch.empty(m, device="cuda"), torch.empty(m, device="cuda")
    dg, db, dbias = torch.zeros(h, device="cuda"), torch.zeros(h, device="cuda"), torch.zeros(h, device="cuda")
    for drop in (0.0, 0.1):
        sd = seed if drop else None
        t_f = graph_time(lambda: ops.layernorm_fwd(x, res, g, b, y, mean, rstd, p_in=drop, site_in=2, seed=sd))
        t_b = graph_time(lambda: ops.layernorm_bwd(dy, x, res, g, mean, rstd, dx=dx, dres=dres if drop else None, dgamma=dg, dbeta=db, dbias=dbias, p_in=drop, site_in=2, seed=sd))
        t_n = graph_time(lambda: ops.layernorm_bwd(dy, x, res, g, mean, rstd, dx=dx, dres=dres if drop else None, p_in=drop, site_in=2, seed=sd))
        print(f"{name} m{m} h{h} drop{drop}: fwd {t_f:6.2f}  bwd {t_b:6.2f}  bwd-no-colsums {t_n:6.2f} us", flush=True)
