#!/usr/bin/env python
"""CPU experiment (oracle only, no GPU): what folding the forward BatchNorm into the consumer convolutions does to the bf16
noise floor.  Runs the fp64 oracle train step three ways on the same inputs -- exact, with the current CUDA path's bf16 storage
points (storage="bf16"), and with the folded path's (storage="bf16_fold": y of the 13 folded producers is never rounded, the
consumers round W * gamma * rstd instead of W) -- conditioned on the exact run's activation pattern, and prints the errors of
both emulations against fp64.  usage: python tools/fold_noise_floor.py [seeds...]"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import unet_oracle as O  # noqa: E402


def run(seed, N=2, C=1, H=64, W=48, K=2):
    p = O.init_params(C, K, seed=seed, base=64, randomize_affine=True)
    x, lab = O.synthetic_batch(N, C, H, W, K, seed=seed)
    xt = torch.tensor(x, dtype=torch.float64)
    oh = torch.tensor(np.eye(K, dtype=np.int32)[lab])
    rng = np.random.default_rng(seed)
    dm = {"drop4": torch.tensor(rng.integers(0, 2, size=(N, 512, H // 8, W // 8))), "dropb": torch.tensor(rng.integers(0, 2, size=(N, 1024, H // 16, W // 16)))}
    taps = {}
    ref = O.train_step_grads(p, xt, oh, N, dm, taps=taps)
    relu = {k[:-4]: (v > 0) for k, v in taps.items() if k.endswith("/act") and not k.startswith("up")}
    out = {}
    for mode in ("bf16", "bf16_fold"):
        r = O.train_step_grads(p, xt, oh, N, dm, relu_masks=relu, storage=mode)
        num = sum(float(((r["grads"][k] - g) ** 2).sum()) for k, g in ref["grads"].items() if k.endswith("/kernel"))
        den = sum(float((g ** 2).sum()) for k, g in ref["grads"].items() if k.endswith("/kernel"))
        worst = max(float(torch.linalg.norm(r["grads"][k] - g) / torch.linalg.norm(g)) for k, g in ref["grads"].items() if k.endswith("/kernel"))
        out[mode] = dict(softmax_max=float((r["softmax"] - ref["softmax"]).abs().max()), softmax_rms=float(((r["softmax"] - ref["softmax"]) ** 2).mean().sqrt()),
                         loss_rel=abs(float(r["loss"]) - float(ref["loss"])) / float(ref["loss"]), grad_l2=float(np.sqrt(num / den)), grad_worst_layer=worst)
    return out


if __name__ == "__main__":
    seeds = [int(s) for s in sys.argv[1:]] or [0, 1, 2]
    torch.set_num_threads(os.cpu_count() or 1)
    for s in seeds:
        print(json.dumps({"seed": s, **run(s)}), flush=True)
