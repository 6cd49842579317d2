#!/usr/bin/env python
"""Runs ub_debug_desc_probe over (mode, shift, sbo, base_offset) and reports which addressing semantic the tensor core
implements for SWIZZLE_128B operands that start off the 1024-byte atom (see csrc/debug.cu)."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import _probe as C  # noqa: E402  (libunetb200_probe.so)

dev = torch.device("cuda")
R = 256
X = (torch.arange(R * 64, device=dev, dtype=torch.float32).reshape(R, 64) % 251 - 125).to(torch.bfloat16)   # exact in bf16
I = torch.eye(64, device=dev, dtype=torch.bfloat16)
Xf = X.float().cpu()
res = []
for mode in (0, 1):
    for sbo in (1024, 1280, 2304):
        for shift in (0, 1, 3, 8, 11):
            for bo in sorted({0, shift & 7}):
                out = torch.full((128, 64), float("nan"), device=dev)
                C.call("ub_debug_desc_probe", X, R, I, out, shift, sbo, bo, mode, torch.cuda.current_stream().cuda_stream)
                torch.cuda.synchronize()
                o = out.cpu()
                g = sbo // 128
                if mode == 0:
                    rows = [shift + (m // 8) * g + m % 8 for m in range(128)]
                    ok_rows = [m for m in range(128) if rows[m] < R]
                    exp = Xf[[rows[m] for m in ok_rows]]
                    good = bool(torch.equal(o[ok_rows], exp))
                else:
                    rows = [shift + (n // 8) * g + n % 8 for n in range(64)]
                    exp = Xf[rows].t()                      # [c][n]
                    good = bool(torch.equal(o[:64], exp))
                res.append(dict(mode=mode, sbo=sbo, shift=shift, base_offset=bo, ok=good))
                print(json.dumps(res[-1]), flush=True)
summ = {}
for m in (0, 1):
    summ[f"mode{m}_absolute_address_swizzle(base_offset=0)"] = all(r["ok"] for r in res if r["mode"] == m and r["base_offset"] == 0)
    summ[f"mode{m}_base_offset=shift&7"] = all(r["ok"] for r in res if r["mode"] == m and r["base_offset"] == (r["shift"] & 7))
print("SUMMARY", json.dumps(summ))
