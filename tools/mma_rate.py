#!/usr/bin/env python
"""tcgen05.mma issue-rate ceiling per tile shape (operands resident in smem), see csrc/debug.cu."""
import json, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import unetb200._C as C  # noqa: E402
dev = torch.device("cuda")
for nblocks in (1, 148):
    for mn in (0, 1):
        for N in (32, 64, 128, 256):
            clk = torch.zeros(nblocks, dtype=torch.int64, device=dev)
            iters = 2000
            C.call("ub_debug_mma_rate", N, iters, mn, nblocks, clk, torch.cuda.current_stream().cuda_stream)
            torch.cuda.synchronize()
            c = clk.float().mean().item() / (iters * 4)
            print(json.dumps(dict(blocks=nblocks, mn_major=mn, N=N, clk_per_mma=round(c, 2), mac_per_clk=round(128 * N * 16 / c, 1),
                                  smem_read_B_per_clk=round((128 * 32 + N * 32) / c, 1))), flush=True)
