#!/usr/bin/env python
"""tcgen05.mma issue-rate ceiling per tile shape (operands resident in smem), incl. A operands that start off the
1024-byte swizzle atom / have a non-1024-multiple group stride (the halo-patch views), see csrc/debug.cu."""
import json, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import _probe as C  # noqa: E402  (libunetb200_probe.so)
dev = torch.device("cuda")
nblocks = 148
for mn in (0, 1):
    for (shift, sbo) in ((0, 1024), (1, 1024), (0, 1280), (11, 1280), (1, 2048), (17, 2048), (3, 2304)):
        for N in (64, 128, 256):
            clk = torch.zeros(nblocks, dtype=torch.int64, device=dev)
            iters = 2000
            C.call("ub_debug_mma_rate", N, iters, mn, shift, sbo, nblocks, clk, torch.cuda.current_stream().cuda_stream)
            torch.cuda.synchronize()
            c = clk.float().mean().item() / (iters * 4)
            print(json.dumps(dict(mn_major=mn, a_shift=shift, a_sbo=sbo, N=N, clk_per_mma=round(c, 2), mac_per_clk=round(128 * N * 16 / c, 1))), flush=True)
