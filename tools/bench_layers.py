#!/usr/bin/env python
"""Micro-benchmark of single tensor-core entry points through the C ABI (CUDA events, L2 flushed between iterations by
rotating over buffers larger than L2).  Usage: python tools/bench_layers.py [case-substring ...]"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import unetb200._C as C  # noqa: E402

dev = torch.device("cuda")


def st():
    return torch.cuda.current_stream().cuda_stream


def rnd(shape, sparse=False):
    t = torch.randn(shape, device=dev, dtype=torch.float32)
    if sparse:
        t = t * (torch.rand(shape, device=dev) > 0.5)
    return t.to(torch.bfloat16)


def timeit(fn, iters=20, warm=3):
    for _ in range(warm):
        fn(0)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for i in range(iters):
        fn(i)
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters


def conv_case(name, N, H, W, C0, C1, Cout, mode, sparse=False, stats=True, nbuf=3):
    Cin = C0 + C1
    flops = 2.0 * 9 * Cin * Cout * N * H * W
    if mode == "fwd":
        xs = [rnd((N, H, W, C0), sparse) for _ in range(nbuf)]
        x1 = [rnd((N, H, W, C1), sparse) for _ in range(nbuf)] if C1 else None
        w = rnd((Cout, 9 * Cin)) * 0.05
        bias = torch.zeros(Cout, device=dev)
        outs = [torch.empty((N, H, W, Cout), device=dev, dtype=torch.bfloat16) for _ in range(nbuf)]
        part = torch.empty(C.UB_STATS_ROWS * 2 * Cout, device=dev)

        def fn(i):
            j = i % nbuf
            C.call("ub_conv3x3_fwd", xs[j], C0, x1[j] if C1 else None, C1, w, bias, outs[j], part if stats else None, N, H, W, Cout, 1, st())
    elif mode == "dgrad":
        dz = [rnd((N, H, W, Cout), sparse) for _ in range(nbuf)]
        wt = rnd((Cin, 9 * Cout)) * 0.05
        dx0 = [torch.empty((N, H, W, C0), device=dev, dtype=torch.bfloat16) for _ in range(nbuf)]
        dx1 = [torch.empty((N, H, W, C1), device=dev, dtype=torch.bfloat16) for _ in range(nbuf)] if C1 else None

        def fn(i):
            j = i % nbuf
            C.call("ub_conv3x3_dgrad", dz[j], Cout, wt, dx0[j], C0, dx1[j] if C1 else None, C1, N, H, W, st())
    else:
        xs = [rnd((N, H, W, C0)) for _ in range(nbuf)]
        x1 = [rnd((N, H, W, C1)) for _ in range(nbuf)] if C1 else None
        dz = [rnd((N, H, W, Cout), sparse) for _ in range(nbuf)]
        dw = torch.empty(Cout * 9 * Cin, device=dev)
        nb = C.lib.ub_conv3x3_wgrad_workspace_bytes(C0, C1, Cout, N, H, W)
        ws = torch.empty(nb, device=dev, dtype=torch.uint8)

        def fn(i):
            j = i % nbuf
            C.call("ub_conv3x3_wgrad", xs[j], C0, x1[j] if C1 else None, C1, dz[j], Cout, dw, ws, nb, N, H, W, st())
    ms = timeit(fn)
    print(json.dumps(dict(case=name, mode=mode, sparse=sparse, stats=stats, ms=round(ms, 4), tflops=round(flops / ms / 1e9, 1))), flush=True)


LAYERS = {  # name: (H, C0, C1, Cout) at batch 16
    "enc1b": (512, 64, 0, 64), "dec1a": (512, 64, 64, 64), "enc2b": (256, 128, 0, 128), "dec2a": (256, 128, 128, 128),
    "enc3b": (128, 256, 0, 256), "dec3a": (128, 256, 256, 256), "enc4b": (64, 512, 0, 512), "botb": (32, 1024, 0, 1024),
}

if __name__ == "__main__":
    pats = sys.argv[1:]
    variants = [("fwd", False, True), ("fwd", False, False), ("fwd", True, True), ("dgrad", False, False), ("dgrad", True, False),
                ("wgrad", False, False), ("wgrad", True, False)]
    modes = os.environ.get("BL_MODES")
    if modes:
        variants = [v for v in variants if v[0] in modes.split(",")]
    for name, (H, C0, C1, Cout) in LAYERS.items():
        if pats and not any(p in name for p in pats):
            continue
        for mode, sparse, stats in variants:
            conv_case(name, 16, H, H, C0, C1, Cout, mode, sparse, stats)
