"""The data-parallel train step on real NCCL: N ranks against N oracle replicas with averaged gradients
and one Adam step.  Needs >= 2 GPUs (``gpurun --gpus 2 -- python -m pytest tests -m gpu``); skipped on one."""
import json
import os
import socket
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_dp_step_nccl_vs_oracle_replicas(precision):
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    world = 2
    out = os.path.join(ROOT, "gpurun_out", f"dp{world}_parity_{precision}.json")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(_free_port()),
           os.path.join(ROOT, "tests", "dp_worker.py"), "--out", out, "--precision", precision]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    with open(out) as f:
        rep = json.load(f)
    assert rep["ok"] and rep["world"] == world and len(rep["steps"]) == 2
