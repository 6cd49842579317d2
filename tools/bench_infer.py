#!/usr/bin/env python
"""Tiled-inference throughput (BASELINE.json config 5: synthetic uint16 image, tile 1024 + halo 96, tiles sharded over
the ranks of a torchrun job).  MPix/s = image pixels / seconds from the raw image in PINNED HOST memory to the final
uint8 mask in host memory (upload, GPU z-score, all tile forwards, mask reduce across ranks, download), max over ranks.

  python tools/bench_infer.py [--size 20000] [--reps 3] [--tile-batch 4]
  torchrun --nproc-per-node 8 tools/bench_infer.py --size 20000
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--size", type=int, default=20000)
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--tile-batch", type=int, default=4)
    ap.add_argument("--classes", type=int, default=2)
    args = ap.parse_args()
    sys.stdout.flush()
    json_fd = os.dup(1)          # stdout carries exactly one JSON line (NCCL prints its banner to fd 1)
    os.dup2(2, 1)
    import torch
    import unetb200.inference as I
    from unetb200.dist import DataParallel
    from unetb200.model import UNet

    world = int(os.environ.get("WORLD_SIZE", "1"))
    dp = DataParallel(backend="nccl") if world > 1 else None
    rank = dp.rank if dp else 0
    torch.cuda.set_device(dp.local_rank if dp else 0)
    dev = torch.device("cuda", torch.cuda.current_device())
    m = UNet(args.classes, 1, 1, 1e-4, precision="bf16", seed=0, dist=dp)
    if dp:
        dp.broadcast_params(m)
    S = args.size
    rng = np.random.default_rng(0)
    # uint16 image with the reference data's statistics; block-wise to keep host memory modest
    host = torch.empty((1, S, S), dtype=torch.int16).pin_memory()
    hv = host.numpy().view(np.uint16)
    for r0 in range(0, S, 2048):
        blk = rng.normal(3045.0, 376.0, size=(min(2048, S - r0), S)).astype(np.float32)
        hv[0, r0:r0 + blk.shape[0]] = np.clip(np.round(blk), 0, 65535).astype(np.uint16)
    pad_y, pad_x = I._pad_amounts(S, S)
    out_host = torch.empty((S, S), dtype=torch.uint8).pin_memory()

    banded = dp is not None and os.environ.get("UB_INFER_BANDED", "0") == "1"      # row-band sharding (first GPU run pending)

    def one():
        if banded:
            I.segment_banded(host, m, dp, I.TILE_SIZE, radius=96, tile_batch=args.tile_batch, out_host=out_host if rank == 0 else None)
            torch.cuda.synchronize(dev)
            return
        raw = host.to(dev, non_blocking=True)
        x = I.zscore_device(raw, m)
        if pad_y or pad_x:
            x = torch.nn.functional.pad(x[None], (0, pad_x, 0, pad_y), mode="reflect")[0].contiguous()
        mask = I.segment_device(x, m, I.TILE_SIZE if S > I.TILE_SIZE else None, tile_batch=args.tile_batch, dist=dp)
        out_host.copy_(mask[:S, :S], non_blocking=True)
        torch.cuda.synchronize(dev)

    one()          # warm-up (allocations, TMEM/smem attribute calls)
    times = []
    for _ in range(args.reps):
        if dp:
            dp.barrier()
        torch.cuda.synchronize(dev)
        t0 = time.perf_counter()
        one()
        if dp:
            dp.barrier()
        times.append(time.perf_counter() - t0)
    t = torch.tensor([min(times)], device=dev)
    if dp:
        t = dp.reduce_max(t)
    secs = float(t.item())
    ntiles = len(I.tile_plan(S + pad_y, S + pad_x, I.TILE_SIZE, 96)) if S > I.TILE_SIZE else 1
    if rank == 0:
        # forward FLOPs actually executed: tiles incl. halo (1.467776 MFLOP/pixel at Cin = 1, K = 2)
        px_exec = sum((tl["y1"] - tl["y0"]) * (tl["x1"] - tl["x0"]) for tl in I.tile_plan(S + pad_y, S + pad_x, I.TILE_SIZE, 96)) if S > I.TILE_SIZE else S * S
        os.write(json_fd, (json.dumps({"metric": "unet_tiled_inference_mpix_per_sec", "value": S * S / secs / 1e6, "unit": "MPix/s", "n_gpus": world,
                          "image": [S, S], "tiles": ntiles, "tile_batch": args.tile_batch, "seconds": secs,
                          "exec_tflops": px_exec * 1.467776e6 / secs / 1e12, "foreground_fraction": float((out_host.numpy() == 1).mean()),
                          "h2d_bytes": int(host.numel() * 2), "d2h_bytes": int(out_host.numel())}) + "\n").encode())
    if dp:
        dp.shutdown()


if __name__ == "__main__":
    main()
