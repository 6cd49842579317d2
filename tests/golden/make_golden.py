"""Generates tests/golden/*.npz from the oracle (oracle/unet_oracle.py, fp64).  The reference itself cannot run here
(no TensorFlow) and ships no golden vectors, so these fixtures pin the ORACLE, not TensorFlow: parity is "unpinned"
in the sense of the task statement.  Re-run:  python tests/golden/make_golden.py"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import unet_oracle as O  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def graph_case(name, N, C, H, W, K, seed, gb):
    p = O.init_params(C, K, seed=seed, base=64, randomize_affine=True)
    rng = np.random.default_rng(seed + 100)
    x = rng.normal(size=(N, C, H, W)).astype(np.float32)
    lab = rng.integers(0, K, size=(N, H, W)).astype(np.uint8)
    oh = np.eye(K, dtype=np.int32)[lab]
    dm = {"drop4": rng.integers(0, 2, size=(N, 512, H // 8, W // 8)).astype(np.uint8),
          "dropb": rng.integers(0, 2, size=(N, 1024, H // 16, W // 16)).astype(np.uint8)}
    r = O.train_step_grads(p, torch.tensor(x, dtype=torch.float64), torch.tensor(oh), gb, {k: torch.tensor(v) for k, v in dm.items()})
    out = dict(N=N, C=C, H=H, W=W, K=K, seed=seed, gb=gb, x=x, labels=lab, drop4=np.packbits(dm["drop4"]), dropb=np.packbits(dm["dropb"]),
               loss=float(r["loss"]), acc=float(r["acc"]), softmax=r["softmax"].numpy().astype(np.float32))
    names = list(r["grads"].keys())
    out["grad_names"] = np.array(names)
    out["grad_absmax"] = np.array([float(r["grads"][k].abs().max()) for k in names])
    out["grad_l2"] = np.array([float(r["grads"][k].norm()) for k in names])
    # a deterministic sample of 64 entries per tensor
    samp_idx, samp_val = [], []
    for k in names:
        g = r["grads"][k].reshape(-1).numpy()
        idx = np.random.default_rng(7).integers(0, g.size, size=64)
        samp_idx.append(idx)
        samp_val.append(g[idx])
    out["grad_sample_idx"] = np.stack(samp_idx)
    out["grad_sample_val"] = np.stack(samp_val)
    for k, v in r["new_stats"].items():
        if k.startswith(("enc1a", "botb", "up2", "head")):
            out["stat:" + k] = v.numpy().astype(np.float32)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(name, "loss", out["loss"], "acc", out["acc"])


def host_case():
    rng = np.random.default_rng(3)
    img = np.clip(np.round(rng.normal(3045, 376, size=(3, 40, 56))), 0, 65535).astype(np.uint16)
    img[2] = 11
    z = O.zscore_normalize(img)
    lab = rng.integers(0, 4, size=(9, 7))
    plan = O.tile_plan(2000, 2512, 1024, 96)
    np.savez_compressed(os.path.join(HERE, "host.npz"), img=img, zscore=z, lab=lab, onehot=O.one_hot(lab, 4),
                        plan=np.array([[t[k] for k in sorted(t)] for t in plan]), plan_keys=np.array(sorted(plan[0])))


if __name__ == "__main__":
    # sizes chosen so that the bottleneck BatchNorm still sees >= 48 samples per channel (well-conditioned statistics)
    graph_case("graph_c1_k2", N=2, C=1, H=96, W=64, K=2, seed=11, gb=2)
    graph_case("graph_c3_k8", N=1, C=3, H=112, W=112, K=8, seed=12, gb=4)
    host_case()
