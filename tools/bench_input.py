#!/usr/bin/env python
"""Input-side throughput on the device: raw uint16 batch (+ uint8 labels) already in HBM -> augmentation with the reference
reader's constants (UNet/imagereader.py:78-85) -> per-tile z-score.  CUDA events, 4 rotating batches (> L2 not needed: the
kernels are gather / streaming passes over 8-50 MB).  Prints one JSON line.  usage: python tools/bench_input.py [--batch 16] [--size 512]"""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import unetb200.augment as UA  # noqa: E402
from unetb200.model import UNet  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=16)
    ap.add_argument("--size", type=int, default=512)
    ap.add_argument("--channels", type=int, default=1)
    ap.add_argument("--iters", type=int, default=20)
    args = ap.parse_args()
    N, C, S = args.batch, args.channels, args.size
    dev = torch.device("cuda", 0)
    m = UNet(2, N, C, seed=0)
    rng = np.random.RandomState(0)
    raws = [torch.tensor(rng.randint(0, 65535, size=(N, C, S, S)).astype(np.uint16).view(np.int16), device=dev) for _ in range(4)]
    labs = [torch.tensor(rng.randint(0, 2, size=(N, S, S)).astype(np.uint8), device=dev) for _ in range(4)]
    aug = UA.DeviceAugmenter(dev, seed=1)
    params = [UA.draw_params(rng, N, S, S, True, True, 0.1, 0.02, 0.1, 2, None) for _ in range(4)]

    def step(i, augment):
        raw, lab = raws[i % 4], labs[i % 4]
        if augment:
            raw, lab = aug(raw, lab, params[i % 4])
        return m.normalize_batch(raw), lab

    out = {}
    for name, augment in (("zscore_only", False), ("augment_zscore", True)):
        for i in range(3):
            step(i, augment)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for i in range(args.iters):
            step(i, augment)
        b.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b) / args.iters
        out[name] = {"ms_per_batch": ms, "images_per_sec": N / ms * 1e3}
    px = N * C * S * S
    # algorithmic bytes per batch: z-score reads the raw pixels twice (statistics, apply) and writes fp32; the augmentation adds
    # two image warps (u16 -> f32 -> f32), two mask warps (u8 -> f32 -> u8), min/max, noise (read + write), two blur passes
    out["algorithmic_bytes"] = {"zscore_only": px * (2 + 2 + 4), "augment_zscore": px * ((2 + 4) + (4 + 4) + 4 + 8 + 16 + (4 + 4 + 4)) + N * S * S * ((1 + 4) + (4 + 1))}
    for k in ("zscore_only", "augment_zscore"):
        out[k]["achieved_GBps"] = out["algorithmic_bytes"][k] / (out[k]["ms_per_batch"] * 1e-3) / 1e9
    out["config"] = {"batch": N, "channels": C, "size": S, "dtype": "u16", "note": "host parameter draws and the small H2D parameter uploads are inside the timed region"}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
