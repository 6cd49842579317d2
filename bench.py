#!/usr/bin/env python
"""bench.py -- U-Net 512x512 training throughput (BASELINE.json metric, config 2 / config 3).

  python bench.py [--gpus N] [--steps K] [--warmup W]            this repo's CUDA path (one process per GPU under torchrun)
  python bench.py --impl reference [--steps K] [--warmup W]      the reference graph on the host cores (oracle, torch-CPU fp32)

A step = fwd + loss + bwd (+ bucketed NCCL all-reduce at N > 1) + Adam + weight repack on one batch of 16 synthetic
1x512x512 images per GPU, 2 classes, bf16 storage / fp32 accumulate, dropout on.  Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

BATCH, CH, HW, CLASSES = 16, 1, 512, 2
METRIC = "unet512_train_images_per_sec"
WORKLOAD = "config2: UNet synthetic 1ch 512x512, 2 classes, batch 16/GPU, fwd+bwd+Adam"
CLASS_WEIGHTS = None


def select_workload(name):
    """config2 (default, the metric's configuration) or config4 (BASELINE.json configs[3]: 3-channel uint8-like 1024x1024 tiles,
    8 classes, class-weighted softmax-CE, 4 images per GPU = the same pixels per step)"""
    global BATCH, CH, HW, CLASSES, METRIC, WORKLOAD, CLASS_WEIGHTS
    if name == "config4":
        BATCH, CH, HW, CLASSES = 4, 3, 1024, 8
        METRIC = "unet1024_train_images_per_sec"
        WORKLOAD = "config4: UNet synthetic 3ch 1024x1024, 8 classes, class-weighted CE, batch 4/GPU, fwd+bwd+Adam"
        CLASS_WEIGHTS = [0.5, 1.0, 1.5, 2.0, 0.75, 1.25, 1.75, 0.9]


# ---------------------------------------------------------------------------------------------- algorithmic work
def layer_flops(nc=None, K=None, H=None, W=None, b=64):
    """per-image dense-MAC FLOPs per layer: name -> (kind, fwd_flops)"""
    nc = CH if nc is None else nc
    K = CLASSES if K is None else K
    H = HW if H is None else H
    W = HW if W is None else W
    out = {}
    lv = lambda l: (H >> (l - 1)) * (W >> (l - 1))
    enc = [("enc1a", nc, b, 1), ("enc1b", b, b, 1), ("enc2a", b, 2 * b, 2), ("enc2b", 2 * b, 2 * b, 2), ("enc3a", 2 * b, 4 * b, 3),
           ("enc3b", 4 * b, 4 * b, 3), ("enc4a", 4 * b, 8 * b, 4), ("enc4b", 8 * b, 8 * b, 4), ("bota", 8 * b, 16 * b, 5), ("botb", 16 * b, 16 * b, 5)]
    for n, ci, co, l in enc:
        out[n] = ("first" if n == "enc1a" else "conv", 2 * 9 * ci * co * lv(l))
    for l in (4, 3, 2, 1):
        c = b << (l - 1)
        out[f"up{l}"] = ("deconv", 2 * 4 * 2 * c * c * lv(l + 1))
        out[f"dec{l}a"] = ("conv", 2 * 9 * 2 * c * c * lv(l))
        out[f"dec{l}b"] = ("conv", 2 * 9 * c * c * lv(l))
    out["head"] = ("head", 2 * b * K * lv(1))
    return out


def step_flops(batch=None, **kw):
    batch = BATCH if batch is None else batch
    fl = layer_flops(**kw)
    fwd = sum(v for _, v in fl.values())
    train = 3 * fwd - fl["enc1a"][1]            # no dgrad for the first layer
    fam = {  # per kernel family, per step
        "conv3_fwd": batch * sum(v for k, v in fl.values() if k == "conv"),                     # the 17 conv3x3 forward launches
        "igemm_fwd": batch * (sum(v for k, v in fl.values() if k in ("conv", "deconv")) * 2),   # fwd + dgrad of every tcgen05 layer
        "igemm_wgrad": batch * sum(v for k, v in fl.values() if k in ("conv", "deconv")),
    }
    return batch * fwd, batch * train, fam


FAMILY = {"ub_conv3x3_fwd": "igemm_fwd", "ub_conv3x3_fwd_cases": "igemm_fwd", "ub_conv3x3_fwd_bn": "igemm_fwd", "ub_deconv2x2_fwd_bn": "igemm_fwd", "ub_conv3x3_dgrad": "igemm_fwd", "ub_conv3x3_dgrad_bnred": "igemm_fwd", "ub_deconv2x2_fwd": "igemm_fwd", "ub_deconv2x2_dgrad": "igemm_fwd", "ub_deconv2x2_dgrad_bnred": "igemm_fwd",
          "ub_conv3x3_wgrad": "igemm_wgrad", "ub_deconv2x2_wgrad": "igemm_wgrad"}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(bf16=float(d.get("bf16_tflops_sustained", d.get("bf16_tflops", 1400.0))), hbm=float(d.get("hbm_gbs", 6650.0)), src="measured")
    return dict(bf16=1400.0, hbm=6650.0, src="fallback")


# ---------------------------------------------------------------------------------------------- clocks
class ClockSampler:
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index=0):
        self.index = index
        self.rows = []
        self._stop = threading.Event()
        self._t = threading.Thread(target=self._run, daemon=True)

    def _run(self):
        # NVML in-process (a sample every ~5 ms, so that even a 0.5 s timed region carries dozens); nvidia-smi as the fallback
        try:
            import pynvml
            pynvml.nvmlInit()
            # NVML enumerates in PCI order and ignores CUDA_VISIBLE_DEVICES: map the CUDA ordinal through the visible list
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            idx = self.index
            if vis:
                ids = [v.strip() for v in vis.split(",") if v.strip()]
                if self.index < len(ids) and ids[self.index].isdigit():
                    idx = int(ids[self.index])
            h = pynvml.nvmlDeviceGetHandleByIndex(idx)
            mx = pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM)
            bits = [pynvml.nvmlClocksEventReasonHwSlowdown, pynvml.nvmlClocksEventReasonHwThermalSlowdown,
                    pynvml.nvmlClocksEventReasonSwThermalSlowdown, pynvml.nvmlClocksEventReasonSwPowerCap]
            while not self._stop.is_set():
                sm = pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)
                rs = pynvml.nvmlDeviceGetCurrentClocksEventReasons(h)
                self.rows.append([str(sm), str(mx)] + ["Active" if rs & b else "Not Active" for b in bits])
                self._stop.wait(0.005)
            return
        except Exception:
            pass
        while not self._stop.is_set():
            try:
                o = subprocess.run(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits"],
                                   capture_output=True, text=True, timeout=5).stdout.strip()
                if o:
                    self.rows.append([c.strip() for c in o.split(",")])
            except Exception:
                pass
            self._stop.wait(0.05)

    def start(self):
        self._t.start()

    def stop(self):
        self._stop.set()
        self._t.join(timeout=6)
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows for i in range(4) if len(r) > 2 + i and r[2 + i].lower().startswith("active")})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons, "samples": len(sm)}


# ---------------------------------------------------------------------------------------------- data
def synthetic_host_batches(nbatches, seed):
    """uint16-like smooth noise, z-scored per tile; labels: thresholded smooth field, ~29% foreground"""
    from scipy.ndimage import gaussian_filter
    rng = np.random.default_rng(seed)
    xs, ls = [], []
    for _ in range(nbatches):
        img = gaussian_filter(rng.normal(3045.0, 376.0, size=(BATCH, CH, HW, HW)).astype(np.float32), sigma=(0, 0, 2, 2))
        img = np.clip(np.round(img), 0, 65535)
        mu = img.mean(axis=(2, 3), keepdims=True)
        sd = img.std(axis=(2, 3), keepdims=True)
        xs.append(((img - mu) / np.where(sd <= 1.0, 1.0, sd)).astype(np.float32))
        if CLASSES == 2:
            f = gaussian_filter(rng.normal(size=(BATCH, HW, HW)).astype(np.float32), sigma=(0, 4, 4))
            ls.append((f > np.quantile(f, 0.71)).astype(np.uint8))
        else:          # argmax of K smoothed noise fields
            f = gaussian_filter(rng.normal(size=(BATCH, CLASSES, HW, HW)).astype(np.float32), sigma=(0, 0, 4, 4))
            ls.append(f.argmax(1).astype(np.uint8))
    return xs, ls


# ---------------------------------------------------------------------------------------------- reference arm
def cpu_train_step_rate(n_img, hw, steps, warmup, threads):
    """oracle (torch-CPU fp32 restatement of UNet/model.py) train steps; returns (images/s normalised to 512^2, secs/step)"""
    import torch
    from oracle import unet_oracle as O
    torch.set_num_threads(threads)
    p = O.init_params(CH, CLASSES, seed=0, base=64, dtype=torch.float32)
    opt = O.KerasAdam(p, 3e-4)
    x, lab = O.synthetic_batch(n_img, CH, hw, hw, CLASSES, seed=0)
    oh = torch.tensor(np.eye(CLASSES, dtype=np.int32)[lab])
    xt = torch.tensor(x)
    g = torch.Generator().manual_seed(0)
    times = []
    for s in range(warmup + steps):
        dm = {"drop4": torch.randint(0, 2, (n_img, 512, hw // 8, hw // 8), generator=g),
              "dropb": torch.randint(0, 2, (n_img, 1024, hw // 16, hw // 16), generator=g)}
        t0 = time.perf_counter()
        O.train_step(p, opt, xt, oh, n_img, dm)
        dt = time.perf_counter() - t0
        if s >= warmup:
            times.append(dt)
    spt = float(np.mean(times))
    return n_img * (hw * hw) / float(HW * HW) / spt, spt


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    # calibrate on one 128^2 image, then size the per-step sample so that (steps+warmup) steps fit in ~200 s
    _, t_cal = cpu_train_step_rate(1, 128, 1, 1, threads)
    per_img_512 = t_cal * 16.0
    budget = 200.0
    total = args.steps + args.warmup
    n_img, hw = 1, 512
    if per_img_512 * total <= budget:
        n_img = int(max(1, min(BATCH, budget // (per_img_512 * total))))
    else:
        hw = 256 if (per_img_512 / 4) * total <= budget else 128
    rate, spt = cpu_train_step_rate(n_img, hw, args.steps, args.warmup, threads)
    sample = f"{n_img} image(s) of {hw}x{hw} per step (same graph, fp32, dropout on), rate normalised to 512x512 images"
    line = {"impl": "reference", "metric": METRIC, "value": rate, "unit": "img/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": spt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "note": "TensorFlow is not installable here; this arm times the oracle's torch-CPU fp32 restatement of the reference graph"},
            "cpu_baseline": {"value": rate, "unit": "img/s", "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": rate, "unit": "img/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def run_inference(args):
    """--workload config5 (BASELINE.json configs[4]): tiled inference of a synthetic size x size uint16 image (inference.py tile 1024 +
    halo 96), rows sharded over the ranks.  `value` = device-resident MPix/s (N = 1) / host-to-host (N > 1: the sharded path starts
    from the host image), `e2e` = pinned host image -> uint8 mask in host memory."""
    rank = int(os.environ.get("RANK", "0"))
    S = args.size
    if args.impl == "reference":
        if rank != 0:
            return
        import torch
        from oracle import unet_oracle as O
        threads = os.cpu_count() or 1
        torch.set_num_threads(threads)
        p = O.init_params(1, 2, seed=0, base=64, dtype=torch.float32)
        fn = O.make_model_fn(p)
        rng = np.random.default_rng(0)
        img = O.zscore_normalize(np.clip(np.round(rng.normal(3045.0, 376.0, size=(512, 512))), 0, 65535).astype(np.float32))
        times = []
        for s in range(args.warmup + args.steps):
            t0 = time.perf_counter()
            np.argmax(fn(np.ascontiguousarray(img[None, None])), axis=-1)          # one 512 x 512 tile of the reference's tile loop per step
            if s >= args.warmup:
                times.append(time.perf_counter() - t0)
        spt = float(np.mean(times))
        # the reference recomputes the halo: a 1024 tile yields an 832 zone, so image pixels per tile pixel = (832 / 1024)^2
        rate = 512 * 512 * (832.0 / 1024.0) ** 2 / spt / 1e6
        line = {"impl": "reference", "metric": "unet_tiled_inference_mpix_per_sec", "value": rate, "unit": "MPix/s", "n_gpus": args.gpus,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": spt * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
                "dtype": "f32", "data": "synthetic",
                "config": {"workload": f"config5: tiled inference, {S}x{S} uint16, tile 1024 halo 96", "note": "oracle torch-CPU fp32 forward on one 512x512 tile per step, scaled by the zone/tile area ratio of 1024 tiles"},
                "cpu_baseline": {"value": rate, "unit": "MPix/s", "cores": threads, "kind": "port", "sample": "one 512x512 tile forward per step"},
                "e2e": {"value": rate, "unit": "MPix/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line), flush=True)
        return
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import bench_infer as BI
    local = int(os.environ.get("LOCAL_RANK", "0"))
    res = BI.measure(S, max(1, min(args.steps, 5)), 4, 2, clock_sampler=ClockSampler(local) if rank == 0 else None)
    if res is None:
        return
    peaks = measured_peaks()
    world = res["n_gpus"]
    dev_rate = res["device_resident_mpix_per_sec"]
    achieved = res["exec_tflop"] / (res["seconds_device_resident"] or res["seconds"]) / world
    line = {"metric": res["metric"], "value": dev_rate if dev_rate else res["value"], "unit": "MPix/s", "n_gpus": world, "steps": max(1, min(args.steps, 5)),
            "warmup": 1, "ms_per_step": res["seconds"] * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "bf16",
            "data": "synthetic",
            "config": {"workload": f"config5: tiled inference, {S}x{S} uint16 1 channel, 2 classes, tile 1024 halo 96, {res['tiles']} tiles, {res['sharding']}",
                       "l2": "400 MPix image (0.8 GB raw, 1.6 GB normalised) and per-tile activations far exceed the 126 MB L2",
                       "step": "one whole image"},
            "e2e": {"value": res["value"], "unit": "MPix/s", "h2d_bytes_per_step": res["h2d_bytes"], "d2h_bytes_per_step": res["d2h_bytes"],
                    "ms_per_step": res["seconds"] * 1e3},
            "gpu_launches": res["gpu_launches"], "clocks": res["clocks"],
            "roofline": {"bound": "tensor", "kernel": "conv3_kernel (inference forward, all tiles incl. halo)", "achieved": achieved, "peak": peaks["bf16"],
                         "unit": "TFLOP/s", "frac": achieved / peaks["bf16"], "traffic": None, "peak_source": peaks["src"],
                         "note": "executed forward FLOPs of all tiles (halo recompute included) / seconds, per GPU; whole forward, not one kernel"},
            "cpu_baseline": None, "foreground_fraction": res["foreground_fraction"], "mask_crc": res["mask_crc"]}
    os.write(json_fd, (json.dumps(line) + "\n").encode())


def write_layer_table(path, layer_ms):
    """per (layer, entry point): ms per step, and TFLOP/s for the tensor-core entry points (events around each launch)"""
    fl = layer_flops()
    rows = []
    for (layer, name), ms in sorted(layer_ms.items(), key=lambda kv: -kv[1]):
        tf = None
        if name in FAMILY and layer in fl:
            tf = BATCH * fl[layer][1] / (ms * 1e-3) / 1e12
        rows.append({"layer": layer, "entry": name, "ms": round(ms, 4), "tflops": round(tf, 1) if tf else None})
    os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
    with open(path, "w") as f:
        json.dump(rows, f, indent=0)


# ---------------------------------------------------------------------------------------------- CUDA arm
def run_cuda(args):
    # stdout carries exactly one JSON line: libraries that print to fd 1 (NCCL's version banner) are sent to stderr
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)
    import torch
    from unetb200.dist import DataParallel
    from unetb200.model import UNet

    world = int(os.environ.get("WORLD_SIZE", "1"))
    dp = DataParallel(backend="nccl") if world > 1 else None
    rank = dp.rank if dp else 0
    local = dp.local_rank if dp else 0
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    model = UNet(CLASSES, BATCH * world, CH, learning_rate=3e-4 / 10, precision="bf16", seed=0, dist=dp,   # warm-up LR (train.py:129)
                 class_weights=CLASS_WEIGHTS)
    if dp:
        dp.broadcast_params(model)

    nb = 4
    xs, ls = synthetic_host_batches(nb, seed=rank)
    hx = [torch.from_numpy(x).pin_memory() for x in xs]
    hl = [torch.from_numpy(l).pin_memory() for l in ls]
    dx = [t.to(dev) for t in hx]
    dl = [t.to(dev) for t in hl]

    def sync_all():
        torch.cuda.synchronize(dev)
        if dp:
            dp.barrier()
            torch.cuda.synchronize(dev)

    # ---- device-resident arm ("value")
    for i in range(args.warmup):
        model.train_step(dx[i % nb], dl[i % nb])
    sync_all()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    l0 = model.launches
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.steps):
        model.train_step(dx[i % nb], dl[i % nb])
    e1.record()
    sync_all()
    launches = model.launches - l0
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if dp:
        ms = dp.reduce_max(ms)
    ms_total = float(ms.item())
    clocks = sampler.stop() if rank == 0 else None
    final_loss = float(model.metrics[0].item())

    # ---- end-to-end arm: pinned host -> device copies and the loss read-back inside the timed region, through the package's public
    # host pipeline (unetb200.train.StepPipeline: the upload of batch i + 1 overlaps step i, the loss of step i is read while step
    # i + 1 runs -- every batch is copied and every loss reaches the host inside the timed region)
    from unetb200.train import StepPipeline
    pipe = StepPipeline(model)
    for i in range(3):
        pipe.feed(hx[i % nb], hl[i % nb])
    pipe.flush()
    sync_all()
    t0 = torch.cuda.Event(enable_timing=True)
    t1 = torch.cuda.Event(enable_timing=True)
    t0.record()
    host_losses = []
    for i in range(args.steps):
        v = pipe.feed(hx[i % nb], hl[i % nb])
        if v is not None:
            host_losses.append(v)
    host_losses += pipe.flush()                  # runs the last uploaded batch and reads the outstanding losses
    t1.record()
    sync_all()
    assert len(host_losses) == args.steps, (len(host_losses), args.steps)
    ms_e = torch.tensor([t0.elapsed_time(t1)], device=dev)
    if dp:
        ms_e = dp.reduce_max(ms_e)
    ms_e2e = float(ms_e.item())

    # the same without the pipeline (copy, step, blocking loss read in sequence): what the copies and the read-back cost when exposed
    dxe = torch.empty_like(dx[0])
    dle = torch.empty_like(dl[0])
    ns = max(1, min(args.steps, 10))
    sync_all()
    t0.record()
    for i in range(ns):
        dxe.copy_(hx[i % nb], non_blocking=True)
        dle.copy_(hl[i % nb], non_blocking=True)
        float(model.train_step(dxe, dle).item())
    t1.record()
    sync_all()
    ms_serial = t0.elapsed_time(t1) / ns

    # ---- per-kernel-family timing (extra steps after the timed region, CUDA events around every launch)
    fam_ms, kern_ms = {}, {}
    nprof = 2
    if rank == 0:
        model.profile = []
    for i in range(nprof):                      # every rank runs these steps (they contain the gradient all-reduce)
        model.train_step(dx[i % nb], dl[i % nb])
    torch.cuda.synchronize(dev)
    if rank == 0:
        layer_ms = {}
        for name, layer, a, b in model.profile:
            t = a.elapsed_time(b)
            kern_ms[name] = kern_ms.get(name, 0.0) + t / nprof
            f = FAMILY.get(name)
            if f:
                fam_ms[f] = fam_ms.get(f, 0.0) + t / nprof
            if name in ("ub_conv3x3_fwd", "ub_conv3x3_fwd_cases", "ub_conv3x3_fwd_bn"):
                fam_ms["conv3_fwd"] = fam_ms.get("conv3_fwd", 0.0) + t / nprof
            layer_ms[(layer, name)] = layer_ms.get((layer, name), 0.0) + t / nprof
        model.profile = None
        if args.layers:
            write_layer_table(args.layers, layer_ms)
    dp_check = None
    if dp and os.environ.get("UB_DP_NOREDUCE", "0") == "1":
        dp_check = {"skipped": "UB_DP_NOREDUCE=1: attribution run without the gradient all-reduce (replicas diverge by design)"}
    elif dp:
        dp.barrier()
        # data-parallel numerics on the live ranks: all-reduced G == sum of per-rank gradients, parameters identical after Adam
        dp_check = dp.verify_step(model, dx[0], dl[0])
        dp_check["buckets"] = len(dp._buckets)
        dp.barrier()

    def finish():
        # graphs that hold captured NCCL kernels must be gone before the communicator is torn down; a stuck teardown
        # would turn a finished measurement into a hung job, so the process leaves without destroying the group
        model._graphs.clear()
        import gc
        gc.collect()
        torch.cuda.synchronize(dev)
        if dp:
            clean = dp.shutdown(model)
            sys.stderr.write(f"rank {rank}: process group teardown {'clean' if clean else 'TIMED OUT'}\n")
            sys.stderr.flush()
            if not clean:
                os._exit(0)

    if rank != 0:
        finish()
        return
    fwd_fl, train_fl, fam_fl = step_flops()
    peaks = measured_peaks()
    # dominant kernel: conv3_kernel, conv3x3 forward (17 launches per step, the largest single entry of the step).
    # achieved = algorithmic FLOPs of those launches / their CUDA-event time (events on the launching stream, eager steps
    # run after the timed region); traffic = their DRAM bytes per step from the committed ncu --set full capture.
    top = "conv3_fwd"
    achieved = fam_fl[top] / (fam_ms[top] * 1e-3) / 1e12 if fam_ms.get(top) else None
    traffic = None
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tp) and args.workload == "config2":
        try:
            traffic = json.load(open(tp)).get(top, {}).get("dram_bytes_per_step")
        except Exception:
            traffic = None
    value = BATCH * world * args.steps / (ms_total * 1e-3)
    e2e = BATCH * world * args.steps / (ms_e2e * 1e-3)
    # CPU baseline: bounded sample on the host cores (rank 0, N=1 only)
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        rate, spt = cpu_train_step_rate(2, 256, 1, 1, threads)
        cpu = {"value": rate, "unit": "img/s", "cores": threads, "kind": "port",
               "sample": f"1 timed + 1 warm-up oracle train step on 2 images of 256x256 ({spt:.1f} s/step), normalised to 512x512 images; TensorFlow unavailable"}
    line = {
        "metric": METRIC, "value": value, "unit": "img/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16",
        "data": "synthetic",
        "config": {"workload": WORKLOAD, "global_batch": BATCH * world, "image": [CH, HW, HW], "classes": CLASSES,
                   "parallelism": f"dp{world}", "l2": "per-step working set (>10 GB of activations) far exceeds the 126 MB L2; 4 rotating input batches",
                   "lr": "3e-5 (warm-up epoch lr/10, train.py:129)", "dropout": "on (Philox)"},
        "e2e": {"value": e2e, "unit": "img/s", "h2d_bytes_per_step": int(hx[0].numel() * 4 + hl[0].numel()), "d2h_bytes_per_step": 8,
                "ms_per_step": ms_e2e / args.steps, "api": "unetb200.train.StepPipeline (one batch of look-ahead, loss read one step behind)",
                "serial_ms_per_step": ms_serial},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": {"bound": "tensor", "kernel": "conv3_pair_kernel / conv3_rows_kernel (the 17 conv3x3 forward launches)", "achieved": achieved, "peak": peaks["bf16"], "unit": "TFLOP/s",
                     "frac": (achieved / peaks["bf16"]) if achieved else None, "traffic": traffic, "peak_source": peaks["src"],
                     "launches_per_step": 17, "algorithmic_tflop_per_step": fam_fl[top] / 1e12,
                     "family_ms_per_step": fam_ms, "family_tflop_per_step": {k: v / 1e12 for k, v in fam_fl.items()},
                     "step_tflop": train_fl / 1e12, "step_frac_of_peak": train_fl / (ms_total / args.steps * 1e-3) / 1e12 / peaks["bf16"]},
        "cpu_baseline": cpu,
        "kernel_ms_per_step": {k: round(v, 3) for k, v in sorted(kern_ms.items(), key=lambda kv: -kv[1])},
        "final_loss": final_loss,
        "dp_check": dp_check,
    }
    os.write(json_fd, (json.dumps(line) + "\n").encode())
    finish()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="cuda", choices=["cuda", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--layers", default=None, help="write a per-layer / per-entry-point timing table (JSON) to this path")
    ap.add_argument("--workload", default="config2", choices=["config2", "config4", "config5"])
    ap.add_argument("--size", type=int, default=20000, help="config5: image edge in pixels")
    args = ap.parse_args()
    if args.workload == "config5":
        return run_inference(args)
    select_workload(args.workload)
    if args.warmup < 3 and args.impl == "cuda":
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_cuda(args)


if __name__ == "__main__":
    main()
