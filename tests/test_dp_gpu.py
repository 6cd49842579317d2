"""-m gpu, needs >= 2 GPUs: the data-parallel step on live NCCL ranks (tools/dp_check.py -> DataParallel.verify_step):
all-reduced gradients == sum of per-rank gradients (graph replay and eager), parameters bit-identical across ranks."""
import json
import os
import socket
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.gpu
@pytest.mark.timeout(600)
def test_data_parallel_numerics_two_ranks():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), os.path.join(ROOT, "tools", "dp_check.py"), "--size", "128", "--batch", "2"]
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=540)
    lines = [l for l in p.stdout.splitlines() if l.startswith("{")]
    assert lines, p.stderr[-2000:]
    r = json.loads(lines[-1])
    assert r["ok"], r
