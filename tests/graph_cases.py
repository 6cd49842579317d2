"""Whole-graph parity: the CUDA train step (through unetb200.model.UNet -> C ABI) against the oracle and the committed
golden fixtures.  Error metric everywhere: max|got-ref| / max|ref| per tensor (relative to the tensor's largest
magnitude; gradients that are analytically zero -- deconv biases feeding straight into BN -- are judged against the
largest gradient magnitude of the layer's kernel instead)."""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from oracle import unet_oracle as O  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")

# tolerances (north_star): bf16 storage <= 1e-2, fp32 check mode <= 1e-4
TOL = {"bf16": dict(softmax=1e-2, loss=1e-2, grad=3e-2, stat=1e-2),
       "fp32": dict(softmax=1e-4, loss=1e-4, grad=1e-4, stat=1e-4)}


def rel(got, ref, floor=0.0):
    got = np.asarray(got, dtype=np.float64)
    ref = np.asarray(ref, dtype=np.float64)
    return float(np.abs(got - ref).max() / max(np.abs(ref).max(), floor, 1e-30))


def load_gold(name):
    g = dict(np.load(os.path.join(GOLD, name + ".npz")))
    N, H, W = int(g["N"]), int(g["H"]), int(g["W"])
    g["drop4"] = np.unpackbits(g["drop4"])[:N * 512 * (H // 8) * (W // 8)].reshape(N, 512, H // 8, W // 8)
    g["dropb"] = np.unpackbits(g["dropb"])[:N * 1024 * (H // 16) * (W // 16)].reshape(N, 1024, H // 16, W // 16)
    return g


def run_model(g, precision):
    from unetb200.model import UNet
    C, K = int(g["C"]), int(g["K"])
    p = O.init_params(C, K, seed=int(g["seed"]), base=64, randomize_affine=True)
    m = UNet(K, int(g["gb"]), C, learning_rate=3e-4, precision=precision, seed=0)
    m.load_oracle_params({k: v.numpy() for k, v in p.items()})
    m.train_step(torch.tensor(g["x"]), torch.tensor(g["labels"]), dropout_masks={"drop4": g["drop4"], "dropb": g["dropb"]},
                 apply_update=False)
    torch.cuda.synchronize()
    return m, p


def case_golden(name, precision):
    g = load_gold(name)
    m, _ = run_model(g, precision)
    tol = TOL[precision]
    met = m.metrics.cpu().numpy()
    grads = m.export_grads()
    r = dict(e_loss=abs(float(met[0]) - float(g["loss"])) / abs(float(g["loss"])), acc=float(met[1]), acc_ref=float(g["acc"]))
    worst, worst_name = 0.0, ""
    names = [str(n) for n in g["grad_names"]]
    for i, n in enumerate(names):
        flat = grads[n].reshape(-1)
        layer = n.split("/")[0]
        floor = float(g["grad_absmax"][names.index(layer + "/kernel")]) * 1e-3
        e = float(np.abs(flat[g["grad_sample_idx"][i]] - g["grad_sample_val"][i]).max() / max(float(g["grad_absmax"][i]), floor))
        l2 = abs(float(np.linalg.norm(flat)) - float(g["grad_l2"][i])) / max(float(g["grad_l2"][i]), floor)
        e = max(e, l2)
        if e > worst:
            worst, worst_name = e, n
    r["e_grad_worst"] = worst
    r["worst_grad"] = worst_name
    stats = m.export_params()
    es = 0.0
    for k in g:
        if k.startswith("stat:"):
            es = max(es, rel(stats[k[5:]], g[k]))
    r["e_stat"] = es
    # softmax through the public model call in training mode is covered by case_live; here compare loss/acc/grads
    r["ok"] = bool(r["e_loss"] < tol["loss"] and worst < tol["grad"] and es < tol["stat"] and abs(r["acc"] - r["acc_ref"]) < 0.02)
    return r


def case_live(precision, N=2, C=1, H=48, W=32, K=2, seed=21, steps=2):
    """oracle run live: softmax, loss, every gradient, then `steps` full optimisation steps (Adam + moving stats)"""
    from unetb200.model import UNet
    tol = TOL[precision]
    p = O.init_params(C, K, seed=seed, base=64, randomize_affine=True)
    rng = np.random.default_rng(seed)
    m = UNet(K, N, C, learning_rate=1e-3, precision=precision, seed=0)
    m.load_oracle_params({k: v.numpy() for k, v in p.items()})
    opt = O.KerasAdam(p, 1e-3)
    r = {}
    for s in range(steps):
        x = rng.normal(size=(N, C, H, W)).astype(np.float32)
        lab = rng.integers(0, K, size=(N, H, W)).astype(np.uint8)
        oh = np.eye(K, dtype=np.int32)[lab]
        dm = {"drop4": rng.integers(0, 2, size=(N, 512, H // 8, W // 8)).astype(np.uint8),
              "dropb": rng.integers(0, 2, size=(N, 1024, H // 16, W // 16)).astype(np.uint8)}
        if s == 0:
            sm = m.forward_softmax(torch.tensor(x), training=True, dropout_masks=dm).cpu().numpy()
            # undo the moving-stat update of this extra forward so both sides see the same number of updates
            m.load_oracle_params({k: v.numpy() for k, v in p.items()})
        ref = O.train_step(p, opt, torch.tensor(x, dtype=torch.float64), torch.tensor(oh), N, {k: torch.tensor(v) for k, v in dm.items()})
        m.train_step(torch.tensor(x), torch.tensor(oh), dropout_masks=dm)      # one-hot label contract
        torch.cuda.synchronize()
        met = m.metrics.cpu().numpy()
        if s == 0:
            r["e_softmax"] = rel(sm, ref["softmax"].numpy())
            r["argmax_agree"] = float((sm.argmax(-1) == ref["softmax"].numpy().argmax(-1)).mean())
            grads = m.export_grads()
            worst, wn = 0.0, ""
            for n, gref in ref["grads"].items():
                layer = n.split("/")[0]
                floor = float(ref["grads"][layer + "/kernel"].abs().max()) * 1e-3
                e = rel(grads[n], gref.numpy(), floor)
                if e > worst:
                    worst, wn = e, n
            r["e_grad_worst"], r["worst_grad"] = worst, wn
        r[f"e_loss{s}"] = abs(float(met[0]) - float(ref["loss"])) / abs(float(ref["loss"]))
    got = m.export_params()
    ew, es = 0.0, 0.0
    for k, v in p.items():
        e = rel(got[k], v.numpy())
        if "moving" in k:
            es = max(es, e)
        else:
            ew = max(ew, e)
    r["e_params_after"] = ew
    r["e_moving_after"] = es
    r["ok"] = bool(r["e_softmax"] < tol["softmax"] and r["e_grad_worst"] < tol["grad"] and all(r[f"e_loss{s}"] < tol["loss"] for s in range(steps))
                   and es < tol["stat"] and ew < (2e-2 if precision == "bf16" else 1e-3) and r["argmax_agree"] > 0.999 - (0.02 if precision == "bf16" else 0))
    return r


def case_inference(precision):
    """training=False path (moving statistics) + test_step loss against the oracle"""
    from unetb200.model import UNet
    N, C, H, W, K = 1, 1, 64, 48, 2
    p = O.init_params(C, K, seed=5, base=64, randomize_affine=True)
    rng = np.random.default_rng(5)
    for k in p:
        if k.endswith("moving_mean"):
            p[k] = torch.tensor(rng.normal(0.3, 0.1, size=p[k].shape))
        if k.endswith("moving_var"):
            p[k] = torch.tensor(rng.uniform(0.5, 1.5, size=p[k].shape))
    m = UNet(K, N, C, precision=precision, seed=0)
    m.load_oracle_params({k: v.numpy() for k, v in p.items()})
    x = rng.normal(size=(N, C, H, W)).astype(np.float32)
    lab = rng.integers(0, K, size=(N, H, W)).astype(np.uint8)
    sm = m.get_keras_model()(x)
    ref = O.test_step(p, torch.tensor(x, dtype=torch.float64), torch.tensor(np.eye(K, dtype=np.int32)[lab]), N)
    loss = float(m.test_step((torch.tensor(x), torch.tensor(lab))).cpu())
    tol = TOL[precision]
    r = dict(e_softmax=rel(sm, ref["softmax"].numpy()), e_loss=abs(loss - float(ref["loss"])) / float(ref["loss"]),
             argmax_agree=float((sm.argmax(-1) == ref["softmax"].numpy().argmax(-1)).mean()))
    r["ok"] = bool(r["e_softmax"] < tol["softmax"] and r["e_loss"] < tol["loss"] and sm.shape == (N, H, W, K))
    return r


CASES = {
    "golden_c1_k2_fp32": lambda: case_golden("graph_c1_k2", "fp32"),
    "golden_c3_k8_fp32": lambda: case_golden("graph_c3_k8", "fp32"),
    "golden_c1_k2_bf16": lambda: case_golden("graph_c1_k2", "bf16"),
    "golden_c3_k8_bf16": lambda: case_golden("graph_c3_k8", "bf16"),
    "live_fp32": lambda: case_live("fp32"),
    "live_bf16": lambda: case_live("bf16"),
    "inference_fp32": lambda: case_inference("fp32"),
    "inference_bf16": lambda: case_inference("bf16"),
}
