#!/usr/bin/env python
"""BASELINE.json configs[0]: the reference's own fixture -- data/images + data/masks (100 pairs of 256 x 256 uint16 / uint8) through
build_lmdb (UNet/build_lmdb.py:191-230, whole images: 512 tiles do not fit a 256 image) -> training with batch 8, 2 classes.

  python tools/config1.py prepare   # HERE (needs /root/reference/data): unetb200.build_lmdb -> tests/golden/_refdata/{train,test}-HES.lmdb
                                    #   (git-ignored: derived from the reference's data; it travels to the GPU box like the built .so)
  python tools/config1.py oracle    # CPU: the oracle (torch fp32 restatement of UNet/model.py) trains on those records -> loss curve
                                    #   tests/golden/config1_oracle_curve.json (committed: oracle OUTPUT, the checker of the GPU run)
  python tools/config1.py cuda      # GPU: unetb200 trains on the same records, same initial weights, same dropout masks, same order;
                                    #   prints both curves side by side + deviations (JSON)
Both sides: records in database key order (shuffle off, augmentation off), per-tile z-score, lr 3e-5 for the warm-up epoch as
train.py:126-132, Keras-initial weights from the oracle's seeded initialiser, dropout masks from a seeded numpy stream."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
REF = os.path.join(ROOT, "tests", "golden", "_refdata")
CURVE = os.path.join(ROOT, "tests", "golden", "config1_oracle_curve.json")
BATCH, K, C, LR, SEED = 8, 2, 1, 3e-4, 11


def prepare():
    import unetb200.build_lmdb as B
    os.makedirs(REF, exist_ok=True)
    B.main("/root/reference/data/images", "/root/reference/data/masks", REF, "HES", 0.8, "tif", 0, 512, seed=SEED)
    print(sorted(os.listdir(REF)))


def batches(db, steps):
    """the same `steps` batches for both sides: raw pixels [8,1,256,256] in the stored dtype + uint8 labels, database order"""
    import contextlib
    import io
    import unetb200.imagereader as R
    with contextlib.redirect_stdout(io.StringIO()):
        rd = R.ImageReader(os.path.join(REF, db), use_augmentation=False, shuffle=False, number_classes=K)
    for _ in range(steps):
        xi, li = rd.next_raw_batch(BATCH)
        xv = xi.numpy()
        if xv.dtype == np.int16:
            xv = xv.view(np.uint16)
        yield xv.copy(), li.numpy().copy()


def masks(rng, H, W):
    return {"drop4": rng.integers(0, 2, size=(BATCH, 512, H // 8, W // 8)).astype(np.uint8),
            "dropb": rng.integers(0, 2, size=(BATCH, 1024, H // 16, W // 16)).astype(np.uint8)}


def lr_at(step, steps_per_epoch):
    return LR / 10.0 if step <= min(1000, steps_per_epoch) else LR          # train.py:126-132: the first epoch runs at lr / 10


def oracle(steps):
    import torch
    from oracle import unet_oracle as O
    torch.set_num_threads(os.cpu_count() or 1)
    p = O.init_params(C, K, seed=SEED, base=64, dtype=torch.float32)
    opt = O.KerasAdam(p, LR)
    rng = np.random.default_rng(SEED)
    curve, acc = [], []
    t0 = time.time()
    for s, (xv, lv) in enumerate(batches("train-HES.lmdb", steps)):
        x = np.stack([O.zscore_normalize(xv[i].astype(np.float32)) for i in range(BATCH)])
        dm = masks(rng, x.shape[2], x.shape[3])
        opt.lr = lr_at(s, 10)
        r = O.train_step(p, opt, torch.tensor(x), torch.tensor(np.eye(K, dtype=np.int32)[lv]), BATCH, {k: torch.tensor(v) for k, v in dm.items()})
        curve.append(float(r["loss"]))
        acc.append(float(r["acc"]))
        print(f"oracle step {s}: loss {curve[-1]:.5f} acc {acc[-1]:.4f} ({time.time() - t0:.0f} s)", flush=True)
    json.dump({"steps": steps, "batch": BATCH, "lr": LR, "seed": SEED, "loss": curve, "accuracy": acc,
               "what": "oracle (torch-CPU fp32) training on the reference's data/ fixture via build_lmdb, tools/config1.py"}, open(CURVE, "w"), indent=0)


def cuda(steps):
    import torch
    from oracle import unet_oracle as O
    from unetb200.model import UNet
    ref = json.load(open(CURVE))
    steps = min(steps, ref["steps"])
    p = O.init_params(C, K, seed=SEED, base=64, dtype=torch.float32)
    m = UNet(K, BATCH, C, learning_rate=LR, precision="bf16", seed=0)
    m.load_oracle_params({k: v.numpy() for k, v in p.items()})
    rng = np.random.default_rng(SEED)
    curve, acc = [], []
    for s, (xv, lv) in enumerate(batches("train-HES.lmdb", steps)):
        raw = torch.from_numpy(xv.view(np.int16) if xv.dtype == np.uint16 else xv).to(m.device)
        x = m.normalize_batch(raw)                                    # per-tile z-score on the device (imagereader.py:300)
        dm = masks(rng, x.shape[2], x.shape[3])
        m.set_learning_rate(lr_at(s, 10))
        m.train_step(x, torch.from_numpy(lv).to(m.device), dropout_masks=dm)
        met = m.metrics.cpu().numpy()
        curve.append(float(met[0]))
        acc.append(float(met[1]))
    got, want = np.array(curve), np.array(ref["loss"][:steps])
    dev = np.abs(got - want) / want
    out = {"steps": steps, "loss_cuda": [round(v, 5) for v in curve], "loss_oracle": [round(v, 5) for v in want.tolist()],
           "acc_cuda_last": acc[-1], "acc_oracle_last": ref["accuracy"][steps - 1], "dev_first10": float(dev[:10].max()), "dev_mean": float(dev.mean()),
           "dev_max": float(dev.max()), "loss_first": float(want[0]), "loss_last_cuda": float(got[-1]), "loss_last_oracle": float(want[-1])}
    out["ok"] = bool(out["dev_first10"] < 2e-2 and out["dev_mean"] < 5e-2 and out["dev_max"] < 0.25 and got[-1] < 0.8 * got[0])
    return out


if __name__ == "__main__":
    cmd = sys.argv[1] if len(sys.argv) > 1 else "cuda"
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 60
    if cmd == "prepare":
        prepare()
    elif cmd == "oracle":
        oracle(n)
    else:
        print(json.dumps(cuda(n)))
