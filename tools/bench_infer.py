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


def measure(size=20000, reps=3, tile_batch=4, classes=2, banded=None, clock_sampler=None):
    """returns (on rank 0; None elsewhere) a dict with the host-to-host and the device-resident MPix/s.  banded None = segment_sharded
    (tile runs, each rank uploads only its rows) whenever there is more than one rank; UB_INFER_BANDED=0 forces the round-robin tile
    sharding in which every rank uploads and normalises the whole image."""
    import torch
    import unetb200.inference as I
    from unetb200.dist import DataParallel
    from unetb200.model import UNet

    world = int(os.environ.get("WORLD_SIZE", "1"))
    dp = DataParallel(backend="nccl") if world > 1 else None
    rank = dp.rank if dp else 0
    torch.cuda.set_device(dp.local_rank if dp else 0)
    dev = torch.device("cuda", torch.cuda.current_device())
    m = UNet(classes, 1, 1, 1e-4, precision="bf16", seed=0, dist=dp)
    if dp:
        dp.broadcast_params(m)
    S = size
    rng = np.random.default_rng(0)
    # uint16 image with the reference data's statistics; block-wise to keep host memory modest
    host = torch.empty((1, S, S), dtype=torch.int16).pin_memory()
    hv = host.numpy().view(np.uint16)
    for r0 in range(0, S, 2048):
        blk = rng.normal(3045.0, 376.0, size=(min(2048, S - r0), S)).astype(np.float32)
        hv[0, r0:r0 + blk.shape[0]] = np.clip(np.round(blk), 0, 65535).astype(np.uint16)
    pad_y, pad_x = I._pad_amounts(S, S)
    out_host = torch.empty((S, S), dtype=torch.uint8).pin_memory()
    if banded is None:
        banded = dp is not None and os.environ.get("UB_INFER_BANDED", "1") == "1"
    tile = I.TILE_SIZE if S > I.TILE_SIZE else None

    def one():
        if banded:
            I.segment_sharded(host, m, dp, I.TILE_SIZE, radius=96, tile_batch=tile_batch, out_host=out_host if rank == 0 else None)
            torch.cuda.synchronize(dev)
            return
        raw = host.to(dev, non_blocking=True)
        x = I.zscore_device(raw, m)
        mask = I.segment_device(x, m, tile, radius=96, tile_batch=tile_batch, dist=dp)
        out_host.copy_(mask[:S, :S], non_blocking=True)
        torch.cuda.synchronize(dev)

    def timed(fn, n):
        ts = []
        for _ in range(n):
            if dp:
                dp.barrier()
            torch.cuda.synchronize(dev)
            t0 = time.perf_counter()
            fn()
            if dp:
                dp.barrier()
            ts.append(time.perf_counter() - t0)
        t = torch.tensor([min(ts)], device=dev)
        if dp:
            t = dp.reduce_max(t)
        return float(t.item())

    one()          # warm-up (allocations, TMEM/smem attribute calls)
    if clock_sampler is not None and rank == 0:
        clock_sampler.start()
    l0 = m.launches
    secs = timed(one, reps)
    launches = (m.launches - l0) // reps
    clocks = clock_sampler.stop() if (clock_sampler is not None and rank == 0) else None
    # device-resident arm: the normalised image already in HBM, the mask left in HBM (tile forwards + argmax only); one rank's share
    secs_dev = None
    if not banded:
        x = I.zscore_device(host.to(dev), m)

        def resident():
            I.segment_device(x, m, tile, radius=96, tile_batch=tile_batch, dist=dp)
            torch.cuda.synchronize(dev)
        secs_dev = timed(resident, max(1, reps - 1))
        del x
    plan = I.tile_plan(S + pad_y, S + pad_x, I.TILE_SIZE, 96) if tile else [dict(y0=0, y1=S + pad_y, x0=0, x1=S + pad_x)]
    res = None
    if rank == 0:
        # forward FLOPs actually executed: tiles incl. halo (1.467776 MFLOP/pixel at Cin = 1, K = 2)
        px_exec = sum((tl["y1"] - tl["y0"]) * (tl["x1"] - tl["x0"]) for tl in plan)
        res = {"metric": "unet_tiled_inference_mpix_per_sec", "value": S * S / secs / 1e6, "unit": "MPix/s", "n_gpus": world,
               "image": [S, S], "tiles": len(plan), "tile_batch": tile_batch, "seconds": secs, "sharding": "tile runs, per-rank row upload" if banded else "round-robin tiles, whole image per rank",
               "device_resident_mpix_per_sec": (S * S / secs_dev / 1e6) if secs_dev else None, "seconds_device_resident": secs_dev,
               "exec_tflops": px_exec * 1.467776e6 / secs / 1e12, "exec_tflop": px_exec * 1.467776e6 / 1e12,
               "foreground_fraction": float((out_host.numpy() == 1).mean()), "mask_crc": int(np.bitwise_xor.reduce(out_host.numpy().view(np.uint32).ravel())) if (S * S) % 4 == 0 else None,
               "h2d_bytes": int(host.numel() * 2), "d2h_bytes": int(out_host.numel()), "gpu_launches": int(launches), "clocks": clocks}
    if dp:
        dp.shutdown(m)
    return res


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
    res = measure(args.size, args.reps, args.tile_batch, args.classes)
    if res is not None:
        os.write(json_fd, (json.dumps(res) + "\n").encode())


if __name__ == "__main__":
    main()
