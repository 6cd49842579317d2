#!/usr/bin/env python
"""Data-parallel numerics on live ranks (UNet/model.py:223, :230-235): all-reduced gradients == sum of the per-rank gradients,
parameters bit-identical on every rank after Adam, CUDA-graph replay == eager launch sequence.
  torchrun --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 tools/dp_check.py [--size 256] [--batch 4]
Rank 0 prints one JSON line."""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--size", type=int, default=256)
    ap.add_argument("--batch", type=int, default=4)
    ap.add_argument("--classes", type=int, default=2)
    ap.add_argument("--steps", type=int, default=3, help="optimisation steps after the check: replicas must stay identical")
    args = ap.parse_args()
    from unetb200.dist import DataParallel
    from unetb200.model import UNet
    dp = DataParallel(backend="nccl")
    dev = torch.device("cuda", dp.local_rank)
    m = UNet(args.classes, args.batch * dp.world_size, 1, learning_rate=1e-3, precision="bf16", seed=100 + dp.rank, dist=dp)   # different seeds:
    dp.broadcast_params(m)                                                                                                   # broadcast must fix it
    rng = np.random.default_rng(7 + dp.rank)
    x = torch.tensor(rng.normal(size=(args.batch, 1, args.size, args.size)).astype(np.float32), device=dev)
    lab = torch.tensor(rng.integers(0, args.classes, size=(args.batch, args.size, args.size)).astype(np.uint8), device=dev)
    out = dp.verify_step(m, x, lab)
    for _ in range(args.steps):
        m.train_step(x, lab)
    torch.cuda.synchronize()
    chk = m.P.view(torch.int32).to(torch.int64).sum().reshape(1)
    allc = [torch.empty_like(chk) for _ in range(dp.world_size)]
    torch.distributed.all_gather(allc, chk)
    out["params_identical_after_steps"] = bool(all(int(c) == int(allc[0]) for c in allc))
    out["ok"] = bool(out["ok"] and out["params_identical_after_steps"])
    clean = dp.shutdown(m)
    out["teardown_clean"] = bool(clean)
    if dp.rank == 0:
        print(json.dumps(out), flush=True)
    if not clean:
        os._exit(0 if out["ok"] else 1)
    sys.exit(0 if out["ok"] else 1)


if __name__ == "__main__":
    main()
