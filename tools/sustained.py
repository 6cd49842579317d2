#!/usr/bin/env python
"""Sustained (power-capped) rate of single kernels: each case runs back to back for `secs` seconds while NVML samples SM clock and
board power.  Tells a kernel's energy efficiency apart from its burst speed: the training step runs at the 1000 W cap, so what
bounds it is joules per FLOP / per byte, not the isolated kernel time.  Cases: cuBLAS bf16 8192^3 (the measured-peak reference),
conv3x3 fwd / dgrad / wgrad of several layers, the BatchNorm-backward apply pass (HBM-bound).
  python tools/sustained.py [secs] [case-substring ...]"""
import json
import os
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import unetb200._C as C  # noqa: E402

dev = torch.device("cuda")


class Sampler:
    def __init__(self):
        import pynvml
        pynvml.nvmlInit()
        self.n = pynvml
        self.h = pynvml.nvmlDeviceGetHandleByIndex(0)
        self.rows = []
        self._stop = threading.Event()
        self.t = threading.Thread(target=self.run, daemon=True)

    def run(self):
        while not self._stop.is_set():
            self.rows.append((self.n.nvmlDeviceGetClockInfo(self.h, self.n.NVML_CLOCK_SM), self.n.nvmlDeviceGetPowerUsage(self.h) / 1000.0))
            self._stop.wait(0.01)

    def __enter__(self):
        self.t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        self.t.join()

    def summary(self):
        r = self.rows[len(self.rows) // 3:]          # steady part
        clk = sorted(c for c, _ in r)
        pw = sorted(p for _, p in r)
        return dict(sm_mhz=clk[len(clk) // 2], power_w=round(pw[len(pw) // 2], 1), samples=len(r))


def st():
    return torch.cuda.current_stream().cuda_stream


def rnd(shape):
    return torch.randn(shape, device=dev, dtype=torch.float32).to(torch.bfloat16)


def sustained(name, fn, work, unit, secs):
    for i in range(3):
        fn(i)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 0
    with Sampler() as s:
        t0 = time.time()
        a.record()
        while time.time() - t0 < secs:
            for _ in range(20):
                fn(n)
                n += 1
            torch.cuda.synchronize() if n % 200 == 0 else None
        b.record()
        torch.cuda.synchronize()
    ms = a.elapsed_time(b) / n
    out = dict(case=name, ms=round(ms, 4), rate=round(work / ms / 1e9, 1), unit=unit, iters=n, **s.summary())
    print(json.dumps(out), flush=True)


def conv(name, N, H, C0, C1, Cout, mode, secs, nbuf=3):
    W = H
    Cin = C0 + C1
    flops = 2.0 * 9 * Cin * Cout * N * H * W
    if mode == "fwd":
        xs = [rnd((N, H, W, C0)) for _ in range(nbuf)]
        x1 = [rnd((N, H, W, C1)) for _ in range(nbuf)] if C1 else None
        w = rnd((Cout, 9 * Cin)) * 0.05
        bias = torch.zeros(Cout, device=dev)
        outs = [torch.empty((N, H, W, Cout), device=dev, dtype=torch.bfloat16) for _ in range(nbuf)]
        part = torch.empty(C.UB_STATS_ROWS * 2 * Cout, device=dev)
        fn = lambda i: C.call("ub_conv3x3_fwd", xs[i % nbuf], C0, x1[i % nbuf] if C1 else None, C1, w, bias, outs[i % nbuf], part, N, H, W, Cout, 1, st())
    elif mode == "dgrad":
        dz = [rnd((N, H, W, Cout)) for _ in range(nbuf)]
        wt = rnd((Cin, 9 * Cout)) * 0.05
        dx0 = [torch.empty((N, H, W, C0), device=dev, dtype=torch.bfloat16) for _ in range(nbuf)]
        dx1 = [torch.empty((N, H, W, C1), device=dev, dtype=torch.bfloat16) for _ in range(nbuf)] if C1 else None
        fn = lambda i: C.call("ub_conv3x3_dgrad", dz[i % nbuf], Cout, wt, dx0[i % nbuf], C0, dx1[i % nbuf] if C1 else None, C1, N, H, W, st())
    else:
        xs = [rnd((N, H, W, C0)) for _ in range(nbuf)]
        x1 = [rnd((N, H, W, C1)) for _ in range(nbuf)] if C1 else None
        dz = [rnd((N, H, W, Cout)) for _ in range(nbuf)]
        dw = torch.empty(Cout * 9 * Cin, device=dev)
        nb = C.lib.ub_conv3x3_wgrad_workspace_bytes(C0, C1, Cout, N, H, W)
        ws = torch.empty(nb, device=dev, dtype=torch.uint8)
        fn = lambda i: C.call("ub_conv3x3_wgrad", xs[i % nbuf], C0, x1[i % nbuf] if C1 else None, C1, dz[i % nbuf], Cout, dw, ws, nb, N, H, W, st())
    sustained(f"{name}_{mode}", fn, flops, "TFLOP/s", secs)


def main():
    secs = float(sys.argv[1]) if len(sys.argv) > 1 else 3.0
    pats = sys.argv[2:]
    want = lambda n: not pats or any(p in n for p in pats)
    if want("cublas"):
        a, b = rnd((8192, 8192)), rnd((8192, 8192))
        c = torch.empty((8192, 8192), device=dev, dtype=torch.bfloat16)
        sustained("cublas_bf16_8192", lambda i: torch.matmul(a, b, out=c), 2.0 * 8192 ** 3, "TFLOP/s", secs)
    if want("copy"):
        x = torch.empty(1 << 30, device=dev, dtype=torch.bfloat16)
        y = torch.empty_like(x)
        sustained("copy_2GiB", lambda i: y.copy_(x), 2.0 * x.numel() * 2, "GB/s", secs)
    layers = {"enc1b": (512, 64, 0, 64), "dec1a": (512, 64, 64, 64), "enc2a": (256, 64, 0, 128), "dec2a": (256, 128, 128, 128), "enc3b": (128, 256, 0, 256), "enc4b": (64, 512, 0, 512), "botb": (32, 1024, 0, 1024)}
    for name, (H, C0, C1, Cout) in layers.items():
        for mode in ("fwd", "dgrad", "wgrad"):
            if want(f"{name}_{mode}"):
                conv(name, 16, H, C0, C1, Cout, mode, secs)
    if want("bn_bwd_apply"):
        M, Cc = 16 * 512 * 512, 64
        dy, a = rnd((M, Cc)), rnd((M, Cc))
        v = torch.ones(Cc, device=dev)
        part = torch.empty(C.UB_STATS_ROWS * Cc, device=dev)
        dz = torch.empty_like(dy)
        sustained("bn_bwd_apply_l1", lambda i: C.call("ub_bn_bwd_apply", dy, a, v, v, v, v, v, dz, part, M, Cc, 1, C.UB_BF16, st()), 3.0 * M * Cc * 2, "GB/s", secs)


if __name__ == "__main__":
    main()
