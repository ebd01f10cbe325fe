"""Data-parallel gradient exchange on real GPUs (SURVEY.md §8 row e): two ranks, one process per GPU over NCCL, launched
with torchrun exactly as the driver launches bench.py.  After backward, every rank's gradients must equal the hand-computed
mean of the per-rank single-GPU gradients -- with the fp32 exchange (default) and with the opt-in bf16 exchange, the setting
the scaling bench runs on (`config.grad_exchange`).  Skipped on a box with fewer than two GPUs; the same exchange is covered
on the CPU over gloo in tests/test_engine_schedule_cpu.py and tests/test_ddp_cpu.py."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("exchange,bar", [("fp32", 1e-3), ("bf16", 1e-2), ("bf16-switch", 1.5e-2)])
def test_two_rank_gradient_exchange_equals_hand_average(exchange, bar):
    """fp32 over NCCL (default), bf16 over NCCL, and bf16 through our own NVSwitch kernels (weight gradients written in bf16:
    one rounding before and one after the fp32 in-switch sum)."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    env = dict(os.environ, VB_DDP_FP32="1" if exchange == "fp32" else "0",
               VB_DDP_TRANSPORT="switch" if exchange == "bf16-switch" else "nccl")
    port = 29500 + os.getpid() % 400 + {"fp32": 0, "bf16": 1, "bf16-switch": 2}[exchange]
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "tools", "ddp_smoke.py")]
    r = subprocess.run(cmd, env=env, cwd=ROOT, capture_output=True, text=True, timeout=240)
    out = r.stdout + r.stderr
    assert r.returncode == 0, out[-3000:]
    assert out.count("OK") >= 2, out[-3000:]
    import re
    worst = [float(x) for x in re.findall(r"hand-averaged: ([0-9.eE+-]+?)(?=\[|\s|$)", out)]      # the two ranks' lines may interleave
    assert len(worst) >= 2 and max(worst) < bar, (worst, out[-1500:])
