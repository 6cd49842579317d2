"""Whole-graph parity: the CUDA train step / inference forward (unetb200.model.UNet -> C ABI) against the oracle
(oracle/unet_oracle.py, fp64) and the committed golden fixtures (tests/golden/, generated from the oracle).

Error metric: max|got - ref| / max|ref| per tensor (north_star: <= 1e-2 bf16, <= 1e-4 fp32 check mode).

Gradients are compared CONDITIONED on the activation pattern of the implementation under test: the oracle is re-run
with the CUDA path's ReLU masks [a > 0] and max-pool argmax slots injected (oracle/unet_oracle.py `relu_masks`,
`pool_idx`).  The U-Net is piecewise linear in its activations; a pre-activation within rounding distance of zero
flips its mask between ANY two implementations (torch fp32 vs torch fp64 already differ by 5-10 % in max-norm on these
shapes because of 1-3 such flips, see DESIGN.md "Parity methodology"), which says nothing about kernel correctness.
The forward quantities (softmax, loss, accuracy, BN statistics) are compared against the UNCONDITIONED oracle, and
the unconditioned gradient error is reported too (`e_grad_uncond`, bounded loosely).
Analytically-zero gradients (deconv biases feed straight into BatchNorm) are judged against the layer's kernel
gradient magnitude.

bf16 product path, TRAINING mode: 23 BatchNorm-after-ReLU layers with batch statistics amplify a perturbation by
~250x at Keras-initial weights (measured: fp32 arithmetic noise 1e-7 -> 2.5e-5 on the softmax), so ANY implementation
that stores activations in bf16 (2^-9 per element) sits at ~0.1 max-norm on the softmax and ~5 % on gradients against
fp64 -- the oracle itself does when its activations/gradients are rounded to bf16 at the CUDA path's storage points
(`storage="bf16"`).  The criterion for that mode is therefore the NOISE FLOOR: the CUDA path's error against fp64
must not exceed the emulated bf16-storage oracle's error against fp64 by more than a small factor (per tensor).
The absolute <= 1e-2 bound of north_star is enforced where it is attainable: every kernel (tests/kernel_cases.py),
the whole graph in inference mode (case_inference), loss values, and the loss curve (case_curve).
"""
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
TOL = {"bf16": dict(softmax=1e-2, loss=1e-2, grad=1e-2, stat=1e-2, uncond=0.25, agree=0.99),
       "fp32": dict(softmax=1e-4, loss=1e-4, grad=1e-4, stat=1e-4, uncond=0.5, agree=0.9995)}


# bf16 product path, one training step at Keras-initial weights, shapes with >= 512 BatchNorm samples per channel everywhere
# (2 x 1 x 256 x 256 and larger): absolute bounds against the fp64 oracle.  Measured on B200 (profiles/r02_parity.md); the oracle's
# own bf16-storage emulation sits at logits 0.095, softmax rms 0.026, gradients 0.05-0.09 relative L2 per kernel (conditioned).
BF16_BOUNDS_256 = dict(logits=0.15, sm_rms=0.04, stat=0.02, grad_l2=0.09, grad_l2_worst=0.13, agree=0.96)
# the same path at TRAINED weights (case_trained): north_star's bounds hold -- logits / softmax / conditioned gradients <= 1e-2-ish,
# argmax agreement >= 99.9 % (measured: logits 1.9e-3, softmax 2.9e-3, worst kernel gradient 1.1e-2 rel. L2, agreement 99.98 %)
TRAINED_BOUNDS = dict(agree=0.999, infer_softmax=1e-2, logits=1e-2, grad_l2=2e-2)


def rel(got, ref, floor=0.0):
    got = np.asarray(got, dtype=np.float64)
    ref = np.asarray(ref, dtype=np.float64)
    return float(np.abs(got - ref).max() / max(np.abs(ref).max(), floor, 1e-30))


def grad_errors(got, ref):
    """per-tensor max-norm relative error; deconv biases (analytically zero) relative to the kernel gradient"""
    out = {}
    for n, gref in ref.items():
        gref = gref.numpy() if hasattr(gref, "numpy") else np.asarray(gref)
        layer = n.split("/")[0]
        floor = 0.0
        if layer.startswith("up") and n.endswith("/bias"):
            k = ref[layer + "/kernel"]
            floor = float(np.abs(k.numpy() if hasattr(k, "numpy") else k).max())
        out[n] = rel(got[n], gref, floor)
    return out


def make_inputs(N, C, H, W, K, seed):
    rng = np.random.default_rng(seed)
    x = rng.normal(size=(N, C, H, W)).astype(np.float32)
    lab = rng.integers(0, K, size=(N, H, W)).astype(np.uint8)
    dm = {"drop4": rng.integers(0, 2, size=(N, 512, H // 8, W // 8)).astype(np.uint8),
          "dropb": rng.integers(0, 2, size=(N, 1024, H // 16, W // 16)).astype(np.uint8)}
    return x, lab, dm


def oracle_pattern(taps, dmt):
    """activation pattern (ReLU masks, pool slots) of an oracle run, in the format export_activation_pattern() gives"""
    relu = {n[:-4]: (taps[n].detach() > 0) for n in taps if n.endswith("/act") and not n.startswith("up")}
    pool = {}
    for lvl in (1, 2, 3, 4):
        y = taps[f"enc{lvl}b/out"].detach()
        if lvl == 4 and "drop4" in dmt:
            y = y * dmt["drop4"] * 2.0
        n, c, h, w = y.shape
        win = y.reshape(n, c, h // 2, 2, w // 2, 2).permute(0, 1, 2, 4, 3, 5).reshape(n, c, h // 2, w // 2, 4)
        pool[f"pool{lvl}"] = win.argmax(-1)
    return relu, pool


def cuda_logits(m, N, H, W):
    """what the Softmax layer consumes (UNet/model.py:136-142): BatchNorm (batch statistics of the last training forward) of the
    ReLU'd 1x1 conv, rebuilt from the head's saved fp32 activation"""
    L = m.layers["head"]
    K = m.number_classes
    a = m._b("a:head")[:N * H * W * K].view(N, H, W, K).double()
    o = L.off_stat
    mean, rstd = m.mean[o:o + K].double(), m.rstd[o:o + K].double()
    gamma, beta = m.P[L.off_gamma:L.off_gamma + K].double(), m.P[L.off_beta:L.off_beta + K].double()
    return ((a - mean) * rstd * gamma + beta).cpu().numpy()


def case_live(precision, N=2, C=1, H=64, W=48, K=2, seed=21, gb=None, learn=False, absolute=None, floor=False, smooth=0.0):
    """one training step: softmax / loss / accuracy / BN moving statistics vs the oracle; all 92 gradients vs the
    oracle conditioned on the CUDA path's activation pattern.  learn: image and labels are correlated (the oracle's
    synthetic_batch, the shape of bench.py's workload) instead of independent noise"""
    from unetb200.model import UNet
    tol = TOL[precision]
    gb = gb or N
    p = O.init_params(C, K, seed=seed, base=64, randomize_affine=True)
    x, lab, dm = make_inputs(N, C, H, W, K, seed)
    if learn:
        x, lab = O.synthetic_batch(N, C, H, W, K, seed=seed)
    oh = np.eye(K, dtype=np.int32)[lab]
    m = UNet(K, gb, C, learning_rate=1e-3, label_smoothing=smooth, precision=precision, seed=0)
    m.load_oracle_params({k: v.numpy() for k, v in p.items()})
    m.train_step(torch.tensor(x), torch.tensor(oh), dropout_masks=dm, apply_update=False, keep_softmax=True)   # one-hot label contract
    torch.cuda.synchronize()
    met = m.metrics.cpu().numpy()
    sm = m._b("softmax")[:N * H * W * K].view(N, H, W, K).cpu().numpy()
    grads = m.export_grads()
    stats = m.export_params()
    relu, pool = m.export_activation_pattern(N, H, W)

    xt, oht = torch.tensor(x, dtype=torch.float64), torch.tensor(oh)
    dmt = {k: torch.tensor(v) for k, v in dm.items()}
    ref = O.train_step_grads(p, xt, oht, gb, dmt, label_smoothing=smooth)
    refc = O.train_step_grads(p, xt, oht, gb, dmt, relu_masks=relu, pool_idx=pool, label_smoothing=smooth)

    r = {}
    r["e_softmax"] = rel(sm, ref["softmax"].numpy())
    r["e_logits"] = rel(cuda_logits(m, N, H, W), ref["logits"].numpy())
    r["argmax_agree"] = float((sm.argmax(-1) == ref["softmax"].numpy().argmax(-1)).mean())
    r["e_loss"] = abs(float(met[0]) - float(ref["loss"])) / abs(float(ref["loss"]))
    r["e_acc"] = abs(float(met[1]) - float(ref["acc"]))
    r["e_stat"] = max(rel(stats[k], v.numpy()) for k, v in ref["new_stats"].items())
    ec = grad_errors(grads, refc["grads"])
    eu = grad_errors(grads, ref["grads"])
    wc = max(ec, key=ec.get)
    wu = max(eu, key=eu.get)
    r["e_grad"], r["worst_grad"] = ec[wc], wc
    r["e_grad_uncond"], r["worst_uncond"] = eu[wu], wu
    r["e_loss_cond"] = abs(float(refc["loss"]) - float(ref["loss"])) / abs(float(ref["loss"]))
    fl = 0
    t = {}
    O.forward(p, xt, True, dmt, None, t)
    for n, mk in relu.items():
        fl += int(((t[n + "/act"] > 0) != mk).sum())
    r["relu_flips"] = fl
    if os.environ.get("UB_VERBOSE"):
        r["grad_errs"] = {k: float(f"{v:.3g}") for k, v in ec.items()}
    if precision == "fp32":
        r["ok"] = bool(r["e_softmax"] < tol["softmax"] and r["e_loss"] < tol["loss"] and r["e_stat"] < tol["stat"]
                       and r["e_grad"] < tol["grad"] and r["e_grad_uncond"] < tol["uncond"] and r["argmax_agree"] >= tol["agree"]
                       and r["e_acc"] < 0.01 and r["e_loss_cond"] < 1e-3)
        return r
    # ---- bf16: per-tensor relative L2 against the fp64 oracle, conditioned and unconditioned
    l2 = lambda a, b: float(np.linalg.norm(np.asarray(a, dtype=np.float64) - b) / max(np.linalg.norm(b), 1e-300))
    rms = lambda a, b: float(np.sqrt(np.mean((np.asarray(a, dtype=np.float64) - b) ** 2)))
    sm64 = ref["softmax"].numpy()
    r["sm_rms"] = rms(sm, sm64)
    kern = [n for n in refc["grads"] if not (n.startswith("up") and n.endswith("/bias"))]
    l2c = {n: l2(grads[n], refc["grads"][n].numpy()) for n in kern}
    l2u = {n: l2(grads[n], ref["grads"][n].numpy()) for n in kern}
    num_g = sum(float(np.sum((grads[n] - refc["grads"][n].numpy()) ** 2)) for n in kern)
    den_g = sum(float(np.sum(refc["grads"][n].numpy() ** 2)) for n in kern)
    r["grad_l2"] = float(np.sqrt(num_g / den_g))
    wl = max((n for n in kern if n.endswith("/kernel")), key=lambda n: l2c[n])
    r["grad_l2_worst_kernel"], r["grad_l2_worst_kernel_name"] = l2c[wl], wl
    r["grad_l2_uncond_worst_kernel"] = max(l2u[n] for n in kern if n.endswith("/kernel"))
    zb = max(rel(grads[f"up{l}/bias"], 0 * grads[f"up{l}/bias"], float(np.abs(refc["grads"][f"up{l}/kernel"].numpy()).max())) for l in (1, 2, 3, 4))
    r["e_zero_bias"] = zb
    if os.environ.get("UB_VERBOSE"):
        r["grad_l2_cond_per"] = {k: float(f"{v:.3g}") for k, v in l2c.items() if k.endswith("kernel")}
        r["grad_l2_uncond_per"] = {k: float(f"{v:.3g}") for k, v in l2u.items() if k.endswith("kernel")}
    base_ok = bool(r["e_loss"] < tol["loss"] and r["e_acc"] < 0.01 and r["e_loss_cond"] < 1e-3 and zb < 1e-2)
    if absolute is not None:
        # fixed bounds measured on B200 at a shape whose BatchNorm statistics are well conditioned (profiles/r02_parity.md):
        # north_star's 1e-2 holds for the loss; softmax / logits / gradients sit at the bf16-storage floor of this graph
        r["bounds"] = absolute
        r["ok"] = bool(base_ok and r["e_logits"] <= absolute["logits"] and r["sm_rms"] <= absolute["sm_rms"]
                       and r["e_stat"] <= absolute["stat"] and r["grad_l2"] <= absolute["grad_l2"]
                       and r["grad_l2_worst_kernel"] <= absolute["grad_l2_worst"] and r["argmax_agree"] >= absolute["agree"])
        if not floor:
            return r
    # ---- noise-floor criterion against the bf16-storage emulation of the oracle (module docstring)
    taps = {}
    emu = O.train_step_grads(p, xt, oht, gb, dmt, taps=taps, storage="bf16_fold" if m.fold_bn else "bf16", label_smoothing=smooth)      # same storage points as the path under test
    relu_e, pool_e = oracle_pattern(taps, dmt)
    refe = O.train_step_grads(p, xt, oht, gb, dmt, relu_masks=relu_e, pool_idx=pool_e, label_smoothing=smooth)
    sme = emu["softmax"].numpy()
    r["sm_rms_floor"] = rms(sme, sm64)
    r["sm_max_floor"] = rel(sme, sm64)
    r["logits_floor"] = rel(emu["logits"].numpy(), ref["logits"].numpy())
    r["stat_floor"] = max(rel(emu["new_stats"][k].numpy(), v.numpy()) for k, v in ref["new_stats"].items())
    r["agree_floor"] = float((sme.argmax(-1) == sm64.argmax(-1)).mean())
    worst_ratio, worst_name, num_e, den_e = 0.0, "", 0.0, 0.0
    per = {}
    for n in kern:
        be = refe["grads"][n].numpy()
        ee = l2(emu["grads"][n].numpy(), be)
        per[n] = (l2c[n], ee)
        num_e += float(np.sum((emu["grads"][n].numpy() - be) ** 2)); den_e += float(np.sum(be ** 2))
        ratio = l2c[n] / (ee + 2e-3)
        if ratio > worst_ratio:
            worst_ratio, worst_name = ratio, n
    r["grad_l2_floor"] = float(np.sqrt(num_e / den_e))
    r["grad_l2_floor_worst_kernel"] = max(v[1] for k, v in per.items() if k.endswith("/kernel"))
    r["grad_worst_ratio"], r["grad_worst_ratio_name"] = worst_ratio, worst_name
    if os.environ.get("UB_VERBOSE"):
        r["grad_l2_per"] = {k: (float(f"{a:.3g}"), float(f"{b:.3g}")) for k, (a, b) in per.items() if k.endswith("kernel")}
    if absolute is not None:
        return r
    r["ok"] = bool(base_ok
                   and r["sm_rms"] <= 1.5 * r["sm_rms_floor"] + 1e-4 and r["e_softmax"] <= 2.0 * r["sm_max_floor"] + 1e-3
                   and r["e_stat"] <= 2.0 * r["stat_floor"] + 1e-3 and r["argmax_agree"] >= r["agree_floor"] - 0.02
                   and r["grad_l2"] <= 1.25 * r["grad_l2_floor"] + 1e-3 and worst_ratio <= WORST_RATIO_MAX)
    return r


# Per-tensor companion of the aggregate criterion above (grad_l2 <= 1.25 x floor): the worst of ~90 ratios between the CUDA path's error and
# the emulation's error on the same tensor.  Both are single realisations of bf16 rounding noise at an ill-conditioned toy shape (24-63 samples
# per channel at the bottleneck BatchNorm), and a 64-entry beta gradient moves by tens of per cent when only the summation ORDER of the forward
# statistics changes: builds that differ in nothing else measured 1.24, 1.27, 1.43 and 1.61 (profiles/r02_parity.md).  2.0 separates that
# spread from a real defect (a wrong tap or a dropped term shows up as a ratio of 5-50 and fails the aggregate bound as well).
WORST_RATIO_MAX = 2.0


def case_curve(precision="bf16", N=4, C=1, H=64, W=64, K=2, steps=100, lr=1e-3, seed=33):
    """north_star: "the loss curve tracking over the first steps": `steps` optimisation steps (Adam, BN moving stats,
    dropout masks injected) on a fixed stream of batches, CUDA path vs oracle (fp32 torch-CPU); the curves must track
    within a band (chaotic divergence of two non-identical float pipelines is expected to grow slowly)."""
    from unetb200.model import UNet
    p = O.init_params(C, K, seed=seed, base=64, dtype=torch.float32)
    m = UNet(K, N, C, learning_rate=lr, precision=precision, seed=0)
    m.load_oracle_params({k: v.numpy() for k, v in p.items()})
    opt = O.KerasAdam(p, lr)
    rng = np.random.default_rng(seed)
    # a learnable task: label = smoothed-noise field thresholded; images carry the field plus noise
    from scipy.ndimage import gaussian_filter
    got, ref = [], []
    for s in range(steps):
        f = gaussian_filter(rng.normal(size=(N, H, W)), sigma=(0, 3, 3))
        lab = (f > 0).astype(np.uint8)
        x = (f[:, None] * 8 + rng.normal(size=(N, C, H, W)) * 0.5).astype(np.float32)
        dm = {"drop4": rng.integers(0, 2, size=(N, 512, H // 8, W // 8)).astype(np.uint8),
              "dropb": rng.integers(0, 2, size=(N, 1024, H // 16, W // 16)).astype(np.uint8)}
        oh = np.eye(K, dtype=np.int32)[lab]
        r = O.train_step(p, opt, torch.tensor(x), torch.tensor(oh), N, {k: torch.tensor(v) for k, v in dm.items()})
        ref.append(float(r["loss"]))
        got.append(float(m.train_step(torch.tensor(x), torch.tensor(lab), dropout_masks=dm).item()))
    got, ref = np.array(got), np.array(ref)
    dev = np.abs(got - ref) / ref
    out = dict(loss_first=float(ref[0]), loss_last_ref=float(ref[-1]), loss_last=float(got[-1]), dev_max=float(dev.max()),
               dev_mean=float(dev.mean()), dev_first10=float(dev[:10].max()), steps=steps)
    out["ok"] = bool(out["dev_first10"] < 2e-2 and out["dev_mean"] < 5e-2 and out["dev_max"] < 0.25 and ref[-1] < 0.8 * ref[0]
                     and got[-1] < 0.8 * got[0])
    return out


def learnable_batch(N, C, H, W, rng, K=2):
    """an image whose label can be learnt from it: label = argmax of smooth random fields (K = 2: one field thresholded at 0),
    image channels = a mix of the fields plus noise"""
    from scipy.ndimage import gaussian_filter
    f = gaussian_filter(rng.normal(size=(N, K - 1 if K == 2 else K, H, W)), sigma=(0, 0, 3, 3))
    f = f / f.std()
    lab = (f[:, 0] > 0).astype(np.uint8) if K == 2 else f.argmax(1).astype(np.uint8)
    mix = np.cos(1.0 + np.arange(C)[:, None] * 1.7 + np.arange(f.shape[1])[None, :] * 0.9)          # the SAME image/label relation in every batch
    x = (np.einsum("ck,nkhw->nchw", mix, f) * 2.0 + rng.normal(size=(N, C, H, W)) * 0.5).astype(np.float32)
    return x, lab


_TRAINED = {}


def trained_model(C=1, K=2, steps=300, N=8, S=128, seed=51):
    """(model, fp64 parameter dict incl. moving statistics, rng, last-10 training loss): `steps` optimisation steps of the bf16
    product path on the learnable synthetic task -- the regime the reference's inference.py runs in.  Cached per process."""
    from unetb200.model import UNet
    key = (C, K, steps, N, S, seed)
    if key not in _TRAINED:
        rng = np.random.default_rng(seed)
        m = UNet(K, N, C, learning_rate=1e-3, precision="bf16", seed=seed)
        losses = []
        for s in range(steps):
            x, lab = learnable_batch(N, C, S, S, rng, K)
            loss = m.train_step(torch.tensor(x), torch.tensor(lab))      # a view of the metrics buffer: read it now or never
            if s >= steps - 10:
                losses.append(float(loss.item()))
        p = {k: torch.tensor(np.asarray(v), dtype=torch.float64) for k, v in m.export_params().items()}
        _TRAINED[key] = (m, p, rng, float(np.mean(losses)))
    return _TRAINED[key]


def case_trained(C=1, K=2, steps=300, N=8, S=128, seed=51, n_eval=2, s_eval=256, bounds=None):
    """The regime the reference is used in: a TRAINED network.  `steps` optimisation steps of the bf16 product path on a
    learnable synthetic task, then -- at those weights and moving statistics, exported to the fp64 oracle --
      (1) inference (training=False) on fresh images: softmax / logits max error and per-pixel argmax agreement
          (north_star: >= 99.9 %), UNet/inference.py:105-107;
      (2) one more training step (dropout off): loss, logits, and all gradients, conditioned and unconditioned
          (north_star: <= 1e-2)."""
    from unetb200.model import UNet
    m, p, rng, last10 = trained_model(C, K, steps, N, S, seed)
    r = dict(train_loss_last10=last10)
    # (1) inference
    x, lab = learnable_batch(n_eval, C, s_eval, s_eval, rng, K)
    sm = m.get_keras_model()(x)
    ref = O.test_step(p, torch.tensor(x, dtype=torch.float64), torch.tensor(np.eye(K, dtype=np.int32)[lab]), n_eval)
    sm64 = ref["softmax"].numpy()
    r["infer_e_softmax"] = float(np.abs(sm - sm64).max())
    r["infer_sm_rms"] = float(np.sqrt(np.mean((sm - sm64) ** 2)))
    r["infer_argmax_agree"] = float((sm.argmax(-1) == sm64.argmax(-1)).mean())
    r["infer_acc_ref"] = float(ref["acc"])
    loss = float(m.test_step((torch.tensor(x), torch.tensor(lab))).cpu()) * (m.global_batch_size / n_eval)
    r["infer_e_loss"] = abs(loss - float(ref["loss"])) / float(ref["loss"])
    # (2) a training step at the trained weights
    m2 = UNet(K, n_eval, C, learning_rate=1e-3, precision="bf16", seed=0)
    m2.load_oracle_params({k: v.numpy() for k, v in p.items()})
    oh = np.eye(K, dtype=np.int32)[lab]
    m2.train_step(torch.tensor(x), torch.tensor(lab), dropout_masks={}, apply_update=False, keep_softmax=True)
    torch.cuda.synchronize()
    grads = m2.export_grads()
    relu, pool = m2.export_activation_pattern(n_eval, s_eval, s_eval)
    xt, oht = torch.tensor(x, dtype=torch.float64), torch.tensor(oh)
    tr = O.train_step_grads(p, xt, oht, n_eval)
    trc = O.train_step_grads(p, xt, oht, n_eval, relu_masks=relu, pool_idx=pool)
    smt = m2._b("softmax")[:n_eval * s_eval * s_eval * K].view(n_eval, s_eval, s_eval, K).cpu().numpy()
    r["train_e_loss"] = abs(float(m2.metrics[0]) - float(tr["loss"])) / float(tr["loss"])
    r["train_e_logits"] = rel(cuda_logits(m2, n_eval, s_eval, s_eval), tr["logits"].numpy())
    r["train_e_softmax"] = float(np.abs(smt - tr["softmax"].numpy()).max())
    r["train_argmax_agree"] = float((smt.argmax(-1) == tr["softmax"].numpy().argmax(-1)).mean())
    l2 = lambda a, b: float(np.linalg.norm(np.asarray(a, dtype=np.float64) - b) / max(np.linalg.norm(b), 1e-300))
    kern = [n for n in tr["grads"] if n.endswith("/kernel")]
    l2c = {n: l2(grads[n], trc["grads"][n].numpy()) for n in kern}
    l2u = {n: l2(grads[n], tr["grads"][n].numpy()) for n in kern}
    mxc = {n: rel(grads[n], trc["grads"][n].numpy()) for n in kern}
    r["grad_l2_cond_worst"], r["grad_l2_uncond_worst"], r["grad_max_cond_worst"] = max(l2c.values()), max(l2u.values()), max(mxc.values())
    if os.environ.get("UB_VERBOSE"):
        r["grad_l2_cond_per"] = {k: float(f"{v:.3g}") for k, v in l2c.items()}
        r["grad_l2_uncond_per"] = {k: float(f"{v:.3g}") for k, v in l2u.items()}
    b = bounds or TRAINED_BOUNDS
    r["bounds"] = b
    r["ok"] = bool(r["infer_argmax_agree"] >= b["agree"] and r["infer_e_softmax"] <= b["infer_softmax"] and r["infer_e_loss"] < 1e-2
                   and r["train_e_loss"] < 1e-2 and r["train_e_logits"] <= b["logits"] and r["grad_l2_cond_worst"] <= b["grad_l2"]
                   and r["train_argmax_agree"] >= b["agree"] and r["train_loss_last10"] < 0.5)
    return r


def load_gold(name):
    g = dict(np.load(os.path.join(GOLD, name + ".npz")))
    N, H, W = int(g["N"]), int(g["H"]), int(g["W"])
    g["drop4"] = np.unpackbits(g["drop4"])[:N * 512 * (H // 8) * (W // 8)].reshape(N, 512, H // 8, W // 8)
    g["dropb"] = np.unpackbits(g["dropb"])[:N * 1024 * (H // 16) * (W // 16)].reshape(N, 1024, H // 16, W // 16)
    return g


def case_golden(name, precision):
    """committed fixture (oracle fp64 outputs): loss, accuracy, softmax and BN statistics tightly; gradients (64
    samples + L2 norm per tensor, UNconditioned) within the loose bound that activation-pattern flips allow."""
    from unetb200.model import UNet
    g = load_gold(name)
    tol = TOL[precision]
    N, C, H, W, K = int(g["N"]), int(g["C"]), int(g["H"]), int(g["W"]), int(g["K"])
    p = O.init_params(C, K, seed=int(g["seed"]), base=64, randomize_affine=True)
    m = UNet(K, int(g["gb"]), C, learning_rate=3e-4, precision=precision, seed=0)
    m.load_oracle_params({k: v.numpy() for k, v in p.items()})
    m.train_step(torch.tensor(g["x"]), torch.tensor(g["labels"]), dropout_masks={"drop4": g["drop4"], "dropb": g["dropb"]},
                 apply_update=False, keep_softmax=True)
    torch.cuda.synchronize()
    met = m.metrics.cpu().numpy()
    sm = m._b("softmax")[:N * H * W * K].view(N, H, W, K).cpu().numpy()
    grads = m.export_grads()
    r = dict(e_loss=abs(float(met[0]) - float(g["loss"])) / abs(float(g["loss"])), e_acc=abs(float(met[1]) - float(g["acc"])),
             e_softmax=rel(sm, g["softmax"]))
    worst, worst_name = 0.0, ""
    names = [str(n) for n in g["grad_names"]]
    for i, n in enumerate(names):
        flat = grads[n].reshape(-1)
        layer = n.split("/")[0]
        scale = float(g["grad_absmax"][i])
        if layer.startswith("up") and n.endswith("/bias"):
            scale = float(g["grad_absmax"][names.index(layer + "/kernel")])
        e = float(np.abs(flat[g["grad_sample_idx"][i]] - g["grad_sample_val"][i]).max() / scale)
        if e > worst:
            worst, worst_name = e, n
    r["e_grad_uncond"], r["worst_grad"] = worst, worst_name
    stats = m.export_params()
    r["e_stat"] = max(rel(stats[k[5:]], g[k]) for k in g if k.startswith("stat:"))
    r["sm_rms"] = float(np.sqrt(np.mean((sm.astype(np.float64) - g["softmax"]) ** 2)))
    if precision == "fp32":
        r["ok"] = bool(r["e_loss"] < tol["loss"] and r["e_softmax"] < tol["softmax"] and r["e_stat"] < tol["stat"] and r["e_acc"] < 0.01
                       and worst < tol["uncond"])
    else:
        # bf16 training mode: absolute bound on the loss; softmax / statistics / gradients within fixed multiples of the
        # bf16-storage noise floor measured with the emulating oracle on these fixtures (sm rms 0.02, stat 0.025; see the
        # module docstring and case_live for the live, per-tensor version of the criterion)
        r["ok"] = bool(r["e_loss"] < tol["loss"] and r["e_acc"] < 0.01 and r["sm_rms"] < 0.05 and r["e_stat"] < 0.06 and worst < 1.5)
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
    # argmax agreement at these RANDOM weights (softmax within a few 1e-3 of uniform at many pixels) is reported; the north_star
    # bound (>= 99.9 %) is asserted in fp32 here and, for the bf16 product path, at trained weights in case_trained / case_tiled_inference
    r["ok"] = bool(r["e_softmax"] < tol["softmax"] and r["e_loss"] < tol["loss"] and sm.shape == (N, H, W, K)
                   and (precision != "fp32" or r["argmax_agree"] >= 0.999))
    return r


def case_tiled_inference(precision):
    """unetb200.inference (device-resident, batched, disjoint zones, argmax written by the head kernel) against the
    oracle's restatement of the UNet/inference.py tile loop wrapped around (a) the SAME CUDA model call -- checks
    the tiling / composition logic exactly -- and (b) for fp32, the fp64 oracle model -- checks the numerics."""
    import tempfile
    from unetb200.model import UNet
    import unetb200.inference as I
    C, K = 1, 2
    rng = np.random.default_rng(7)
    from scipy.ndimage import gaussian_filter
    if precision == "bf16":
        # trained weights and an image of the training distribution: the regime in which the per-pixel argmax is well defined
        _, p, _, _ = trained_model(C, K)
        img = learnable_batch(1, C, 400, 336, rng, K)[0][0, 0]
    else:
        p = O.init_params(C, K, seed=7, base=64, randomize_affine=True)
        for k in p:
            if k.endswith("moving_mean"):
                p[k] = torch.tensor(rng.normal(0.3, 0.1, size=p[k].shape))
            if k.endswith("moving_var"):
                p[k] = torch.tensor(rng.uniform(0.5, 1.5, size=p[k].shape))
        img = gaussian_filter(rng.normal(size=(400, 336)), 3).astype(np.float32)
        img = (img / img.std()).astype(np.float32)
    m = UNet(K, 1, C, precision=precision, seed=0)
    m.load_oracle_params({k: v.numpy() for k, v in p.items()})
    r = {}
    got = I._inference_tiling(img, m, 288)
    ref = O.inference_tiling(img, m.get_keras_model(), 288, 96)
    r["tiling_agree"] = float((got == ref).mean())
    r["shape_ok"] = bool(got.shape == ref.shape and got.dtype == np.int32)
    ragged = img[:390, :330]                                   # not multiples of 16: reflect padding path
    got2 = I._inference_tiling(ragged, m, 288)
    ref2 = O.inference_tiling(ragged, m.get_keras_model(), 288, 96)
    r["ragged_agree"] = float((got2 == ref2).mean())
    got3 = I._inference(ragged, m)
    ref3 = O.inference_whole(ragged, m.get_keras_model())
    r["whole_agree"] = float((got3 == ref3).mean())
    r["fg_fraction"] = float((ref == 1).mean())
    ok = r["shape_ok"] and min(r["tiling_agree"], r["ragged_agree"], r["whole_agree"]) >= 0.999 and got2.shape == ragged.shape
    # numerics: the CUDA tiler + CUDA model against the reference's tile loop around the fp64 ORACLE model (north_star >= 99.9 %)
    ref64 = O.inference_tiling(img.astype(np.float64), O.make_model_fn(p), 288, 96)
    r["oracle_agree"] = float((got == ref64).mean())
    ok = ok and r["oracle_agree"] >= 0.999
    # file path: uint16 TIFF -> GPU z-score -> mask, against the host z-score + oracle tiler around the CUDA model
    raw = np.clip(np.round(3000 + 400 * img), 0, 65535).astype(np.uint16)
    with tempfile.TemporaryDirectory() as d:
        fp = os.path.join(d, "a.tif")
        from PIL import Image
        Image.fromarray(raw).save(fp)
        old_tile = I.TILE_SIZE
        I.TILE_SIZE = 288
        try:
            got4 = I.segment_file(fp, m)
        finally:
            I.TILE_SIZE = old_tile
    ref4 = O.inference_tiling(O.zscore_normalize(raw.astype(np.float32)), m.get_keras_model(), 288, 96)
    r["file_agree"] = float((got4 == ref4).mean())
    ok = ok and r["file_agree"] >= 0.999 and got4.dtype == np.uint8
    r["ok"] = bool(ok)
    return r


def case_train_driver():
    """unetb200.train.train_model end to end on a small ImageMaskPair LMDB: raw uint16 tiles + uint8 labels -> GPU z-score
    -> train / test steps -> checkpoint, test_loss.csv, tensorboard dirs; then unetb200.inference restores the checkpoint."""
    import tempfile
    from scipy.ndimage import gaussian_filter
    import unetb200.imagereader as R
    import unetb200.train as T
    import unetb200.inference as I
    from unetb200.model import UNet
    rng = np.random.default_rng(11)

    def pairs(n, tag):
        out = []
        for i in range(n):
            f = gaussian_filter(rng.normal(size=(64, 64)), 3)
            img = np.clip(3000 + 4000 * f + rng.normal(0, 30, size=f.shape), 0, 65535).astype(np.uint16)
            out.append((f"{tag}{i:03d}.tif", img[..., None], (f > 0.02).astype(np.uint8)))
        return out

    import contextlib
    import io
    import re
    r = {}
    with tempfile.TemporaryDirectory() as d:
        R.write_database(os.path.join(d, "train.lmdb"), pairs(48, "tr"))
        R.write_database(os.path.join(d, "test.lmdb"), pairs(8, "te"))
        out = os.path.join(d, "out")
        log = io.StringIO()
        with contextlib.redirect_stdout(log):
            test_loss = T.train_model(out, 8, 1, os.path.join(d, "train.lmdb"), os.path.join(d, "test.lmdb"), 0, 2, 0, 1e-3, 40, 10, max_epochs=3)
        lines = re.findall(r"Train Epoch (\d+): Batch (\d+)/40: Loss ([0-9.eE+-]+) Accuracy = ([0-9.eE+-]+)", log.getvalue())
        tr = [float(l[2]) for l in lines]
        r["train_steps"] = len(tr)                       # 3 epochs x steps 0..40 inclusive (UNet/train.py:136-138)
        r["train_loss_first10"] = float(np.mean(tr[:10]))
        r["train_loss_last10"] = float(np.mean(tr[-10:]))
        r["train_acc_last10"] = float(np.mean([float(l[3]) for l in lines[-10:]]))
        r["test_loss"] = [round(float(v), 4) for v in test_loss]
        r["files"] = sorted(os.listdir(out))
        # the two files tf.train.Checkpoint.write leaves behind (UNet/train.py:181-184)
        r["ckpt"] = sorted(os.listdir(os.path.join(out, "checkpoint"))) == ["ckpt.data-00000-of-00001", "ckpt.index"]
        r["csv_rows"] = len(open(os.path.join(out, "test_loss.csv")).read().split())
        tb = [f for f in r["files"] if f.startswith("tensorboard-")]
        r["tb"] = bool(tb) and sorted(os.listdir(os.path.join(out, tb[0]))) == ["test", "train"]
        # the checkpoint holds the best epoch: restoring it (UNet/inference.py:191-192) reproduces that epoch's test loss
        m = UNet(2, 8, 1, 1e-4)
        m.load_checkpoint(os.path.join(out, "checkpoint", "ckpt"))
        rd = R.ImageReader(os.path.join(d, "test.lmdb"), use_augmentation=False, shuffle=False, number_classes=2)
        with contextlib.redirect_stdout(io.StringIO()):
            losses = []
            for _ in range(2):                           # count / batch_size + 1 steps, as the driver's test epoch
                xi, li = rd.next_raw_batch(8)
                losses.append(float(m.test_step((m.normalize_batch(xi.cuda()), li.cuda())).item()))
        r["restored_test_loss"] = float(np.mean(losses))
        name, img, mask = pairs(1, "x")[0]
        pred = I._inference(R.zscore_normalize(img[..., 0].astype(np.float32)), m)
        r["holdout_shape_ok"] = bool(pred.shape == mask.shape)
    r["ok"] = bool(r["ckpt"] and r["tb"] and r["csv_rows"] == len(test_loss) == 3 and "test_loss.csv" in r["files"]
                   and r["train_steps"] == 123 and r["train_loss_last10"] < 0.8 * r["train_loss_first10"] and r["train_acc_last10"] > 0.8
                   and abs(r["restored_test_loss"] - min(test_loss)) < 1e-3 * max(1.0, min(test_loss)) and r["holdout_shape_ok"])
    return r


def case_estimate_radius(C=1, K=2, seed=41):
    """UNet.estimate_radius (UNet/model.py:160-202): the input gradient of the inference-mode graph against the oracle's
    autograd (fp64), and the radius derived from it"""
    from unetb200.model import UNet
    p = O.init_params(C, K, seed=seed, base=64, randomize_affine=True)
    rng = np.random.default_rng(seed)
    for k in p:
        if k.endswith("moving_mean"):
            p[k] = torch.tensor(rng.normal(0.2, 0.1, size=p[k].shape))
        if k.endswith("moving_var"):
            p[k] = torch.tensor(rng.uniform(0.5, 1.5, size=p[k].shape))
    noise = rng.normal(size=(1, C, 192, 192))
    m = UNet(K, 1, C, precision="bf16", seed=0)
    m.load_oracle_params({k: v.numpy() for k, v in p.items()})
    import contextlib
    import io
    with contextlib.redirect_stdout(io.StringIO()):
        radius = m.estimate_radius(noise=noise)
    ref_radius, ref_grad = O.estimate_radius(p, C, noise=noise)
    got = m._last_erf_grad
    r = dict(radius=int(radius), ref_radius=int(ref_radius), e_grad=rel(got, ref_grad),
             support=[int((got.max(0) > 1e-8).sum()), int((ref_grad.max(0) > 1e-8).sum())])
    r["ok"] = bool(radius == ref_radius and r["e_grad"] < 1e-3 and radius % 16 == 0)
    return r


def case_checkpoint_roundtrip():
    """save_checkpoint / load_checkpoint (UNet/train.py:96, :181-184; UNet/model.py:81-83) in the TensorBundle format and the
    native one: weights, moving statistics, Adam moments and the step counter survive bit-for-bit, and training continues
    identically from the restored state."""
    import tempfile
    from unetb200.model import UNet
    rng = np.random.default_rng(5)
    x = rng.normal(size=(2, 1, 32, 48)).astype(np.float32)
    lab = rng.integers(0, 2, size=(2, 32, 48)).astype(np.uint8)
    m = UNet(2, 2, 1, 1e-3, seed=3)
    for _ in range(3):
        m.train_step(x, lab, dropout_masks={})
    r = {}
    with tempfile.TemporaryDirectory() as d:
        for fmt in ("tf", "native"):
            path = os.path.join(d, fmt, "ckpt")
            os.makedirs(os.path.dirname(path))
            m.save_checkpoint(path, format=fmt)
            r[fmt + "_files"] = sorted(os.listdir(os.path.dirname(path)))
            m2 = UNet(2, 2, 1, 1e-3, seed=99)
            m2.load_checkpoint(path)
            same = all(bool(torch.equal(getattr(m, a), getattr(m2, a))) for a in ("P", "M", "V", "MM", "MV", "S")) and m2.step_count == m.step_count
            la = float(m.train_step(x, lab, dropout_masks={}, apply_update=False).item())
            lb = float(m2.train_step(x, lab, dropout_masks={}, apply_update=False).item())
            m.step_count -= 1          # apply_update=False still counted the step
            r[fmt] = bool(same and la == lb and bool(torch.equal(m.G, m2.G)))
        wrong = UNet(3, 2, 1, 1e-3, seed=0)
        try:
            wrong.load_checkpoint(os.path.join(d, "tf", "ckpt"))
            r["mismatch_raises"] = False
        except IOError:
            r["mismatch_raises"] = True
    r["ok"] = bool(r["tf"] and r["native"] and r["mismatch_raises"] and r["tf_files"] == ["ckpt.data-00000-of-00001", "ckpt.index"])
    return r


def case_reader_augmented():
    """ImageReader(use_augmentation=True).device_batch: LMDB records -> raw upload -> device augmentation with the reference's
    reader constants (UNet/imagereader.py:78-85, :283-294) -> per-tile z-score; and one train step on the result"""
    import tempfile
    from scipy.ndimage import gaussian_filter
    import unetb200.imagereader as R
    from unetb200.model import UNet
    rng = np.random.default_rng(13)
    recs = []
    for i in range(6):
        f = gaussian_filter(rng.normal(size=(64, 96)), 3)
        img = np.clip(3000 + 4000 * f + rng.normal(0, 30, size=f.shape), 0, 65535).astype(np.uint16)
        recs.append((f"im{i:03d}.tif", img[..., None], (f > 0.02).astype(np.uint8)))
    r = {}
    with tempfile.TemporaryDirectory() as d:
        R.write_database(os.path.join(d, "t.lmdb"), recs)
        import contextlib
        import io
        with contextlib.redirect_stdout(io.StringIO()):
            plain = R.ImageReader(os.path.join(d, "t.lmdb"), use_augmentation=False, shuffle=False, number_classes=2)
            aug = R.ImageReader(os.path.join(d, "t.lmdb"), use_augmentation=True, shuffle=False, number_classes=2, seed=5)
        m = UNet(2, 4, 1, 1e-3, seed=1)
        x0, l0 = plain.device_batch(4, m)
        x0, l0 = x0.clone(), l0.clone()
        x1, l1 = aug.device_batch(4, m)
        r["shape_ok"] = bool(tuple(x1.shape) == (4, 1, 64, 96) and x1.dtype == torch.float32 and tuple(l1.shape) == (4, 64, 96) and l1.dtype == torch.uint8)
        r["mean"] = float(x1.mean(dim=(1, 2, 3)).abs().max())
        r["std"] = [round(float(v), 4) for v in x1.std(dim=(1, 2, 3), unbiased=False)]
        r["labels_in_range"] = bool(int(l1.max()) <= 1)
        r["differs"] = float((x1 - x0).abs().mean())
        r["fg_plain"], r["fg_aug"] = float(l0.float().mean()), float(l1.float().mean())
        loss = float(m.train_step(x1, l1).item())
        r["loss"] = loss
    r["ok"] = bool(r["shape_ok"] and r["mean"] < 1e-3 and all(abs(v - 1.0) < 1e-2 for v in r["std"]) and r["labels_in_range"]
                   and r["differs"] > 0.1 and abs(r["fg_aug"] - r["fg_plain"]) < 0.25 and np.isfinite(loss))
    return r


def case_config1_refdata(steps=60):
    """BASELINE.json configs[0]: the reference's own data/ fixture (100 image / mask pairs) through unetb200.build_lmdb, then batch-8
    training -- the CUDA path's loss curve against the oracle's curve on the same records, weights, dropout masks and order
    (tools/config1.py; the oracle curve is the committed fixture tests/golden/config1_oracle_curve.json).  The database is derived
    from the reference's data and is not committed: it is built in the development container (`tools/config1.py prepare`) and travels
    with the working tree; where it is absent the case reports `skipped`."""
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import config1
    if not os.path.exists(os.path.join(config1.REF, "train-HES.lmdb", "data.mdb")) or not os.path.exists(config1.CURVE):
        return dict(ok=True, skipped="tests/golden/_refdata not present (built from /root/reference/data by tools/config1.py prepare)")
    return config1.cuda(steps)


def _with_fold(fn, value="0"):
    """run a graph case with UB_FOLD_BN=value: "0" = the y-materialising training forward (every BatchNorm output written), the
    default ("1") folds the producers' BatchNorm into the consumer convolutions (UNet.fold_bn)"""
    def run():
        old = os.environ.get("UB_FOLD_BN")
        os.environ["UB_FOLD_BN"] = value
        try:
            return fn()
        finally:
            if old is None:
                os.environ.pop("UB_FOLD_BN", None)
            else:
                os.environ["UB_FOLD_BN"] = old
    return run


def case_graph_two_shapes():
    """CUDA-graph replay across a buffer reallocation (UNet._ensure): steps at shape A capture a graph, steps at a LARGER shape B grow the
    persistent buffers (the A graph holds the old pointers and must be dropped), then shape A again -- the parameters must equal, bit
    for bit, those of an identical model that launches every step eagerly"""
    from unetb200.model import UNet
    rng = np.random.default_rng(19)
    shapes = [(2, 32, 48)] * 3 + [(2, 64, 64)] * 3 + [(2, 32, 48)] * 3 + [(4, 64, 64)] * 2 + [(2, 32, 48)] * 2
    data = [(rng.normal(size=(n, 1, h, w)).astype(np.float32), rng.integers(0, 2, size=(n, h, w)).astype(np.uint8)) for n, h, w in shapes]
    out = {}
    for mode in ("graph", "eager"):
        m = UNet(2, 2, 1, 1e-3, seed=7)
        m.use_graph = mode == "graph"
        m.dropout_seed = 11
        losses = []
        for x, lab in data:
            losses.append(float(m.train_step(torch.tensor(x), torch.tensor(lab)).item()))
        torch.cuda.synchronize()
        out[mode] = (m.P.clone(), m.MV.clone(), losses, len(m._graphs))
    same_p = bool(torch.equal(out["graph"][0], out["eager"][0]))
    same_mv = bool(torch.equal(out["graph"][1], out["eager"][1]))
    r = dict(params_bit_identical=same_p, moving_var_bit_identical=same_mv, losses_equal=out["graph"][2] == out["eager"][2],
             finite=bool(np.isfinite(out["graph"][2]).all()), graphs_alive=out["graph"][3])
    r["ok"] = bool(same_p and same_mv and r["losses_equal"] and r["finite"])
    return r


def case_step_pipeline(steps=6):
    """unetb200.train.StepPipeline (pinned host batches, upload one batch ahead on a copy stream, loss read one step behind) gives the
    losses of the same steps run one by one with blocking copies -- same kernels, same order, bit for bit -- and hands out exactly one
    loss per batch"""
    import torch
    from unetb200.model import UNet
    from unetb200.train import StepPipeline
    rng = np.random.default_rng(5)
    xs = [torch.from_numpy(rng.normal(size=(2, 1, 64, 64)).astype(np.float32)).pin_memory() for _ in range(3)]
    ls = [torch.from_numpy(rng.integers(0, 2, size=(2, 64, 64)).astype(np.uint8)).pin_memory() for _ in range(3)]
    out = {}
    for mode in ("plain", "pipe"):
        m = UNet(2, 2, 1, learning_rate=1e-3, precision="bf16", seed=7)
        if mode == "plain":
            out[mode] = [float(m.train_step(xs[i % 3].cuda(), ls[i % 3].cuda()).item()) for i in range(steps)]
        else:
            pipe = StepPipeline(m)
            got = []
            for i in range(steps):
                v = pipe.feed(xs[i % 3], ls[i % 3])
                if v is not None:
                    got.append(v)
            got += pipe.flush()
            out[mode] = got
        out[mode + "_p"] = m.P.clone()
    r = dict(n_plain=len(out["plain"]), n_pipe=len(out["pipe"]), losses_equal=out["plain"] == out["pipe"],
             params_bit_identical=bool(torch.equal(out["plain_p"], out["pipe_p"])), finite=bool(np.isfinite(out["pipe"]).all()))
    r["ok"] = bool(r["n_pipe"] == steps and r["losses_equal"] and r["params_bit_identical"] and r["finite"])
    return r


def _with_env(fn, **env):
    """run a graph case with environment switches of unetb200.model.UNet set (e.g. UB_BN_ALGEBRA="1")"""
    def run():
        old = {k: os.environ.get(k) for k in env}
        os.environ.update(env)
        try:
            return fn()
        finally:
            for k, v in old.items():
                if v is None:
                    os.environ.pop(k, None)
                else:
                    os.environ[k] = v
    return run


def case_banded_inference():
    """segment_sharded (tile runs, split z-score) on one rank == zscore_device + segment_device, incl. reflect padding"""
    import unetb200.inference as I
    from unetb200.model import UNet
    rng = np.random.default_rng(8)
    H, W = 1024 + 200 + 6, 1024 + 90
    img = np.clip(rng.normal(3045.0, 376.0, size=(1, H, W)), 0, 65535).astype(np.uint16)
    m = UNet(2, 1, 1, 1e-4, seed=4)
    raw = torch.tensor(img.view(np.int16)).pin_memory()
    x = I.zscore_device(raw.to(m.device), m)
    ref = I.segment_device(x, m, 1024, radius=96)[:H, :W]
    # ... and the in-place tile reader (mirror indexing past the image edge) against explicitly padded, copied tiles
    pad_y, pad_x = I._pad_amounts(H, W)
    xp = torch.nn.functional.pad(x[None], (0, pad_x, 0, pad_y), mode="reflect")[0].contiguous()
    ref2 = torch.zeros((H + pad_y, W + pad_x), dtype=torch.uint8, device=x.device)
    for t in I.tile_plan(H + pad_y, W + pad_x, 1024, 96):
        geo = torch.tensor([[t["cy0"], t["cy1"], t["cx0"], t["cx1"], t["dy"], t["dx"]]], dtype=torch.int32, device=x.device)
        m.predict_tiles_into(xp[None, :, t["y0"]:t["y1"], t["x0"]:t["x1"]].contiguous(), geo, ref2, W + pad_x)
    reader_agree = float((ref2[:H, :W] == ref).float().mean())

    class D:
        rank, world_size = 0, 1
    got = I.segment_sharded(raw, m, D, 1024, radius=96)
    agree = float((got == ref).float().mean())
    return dict(agree=agree, reader_agree=reader_agree, fg=float((ref == 1).float().mean()), ok=bool(agree == 1.0 and reader_agree == 1.0))


PENDING_CASES = {}


# report-only measurements (tests/gpu_probe.py --probe): the bf16 product path against fp64 and against the oracle's
# bf16-storage emulation at shapes whose BatchNorm statistics are well conditioned; results go to profiles/r02_parity.md
PROBE_CASES = {
    "probe_smoke_shape": lambda: case_live("bf16", N=2, C=1, H=64, W=48, K=2, seed=3, floor=True),
    "probe_256_n2": lambda: case_live("bf16", N=2, C=1, H=256, W=256, K=2, seed=3, floor=True),
    "probe_256_n4_learn": lambda: case_live("bf16", N=4, C=1, H=256, W=256, K=2, seed=4, learn=True, floor=True),
    "probe_trained": lambda: case_trained(),
    "probe_nofold_256_n2": _with_fold(lambda: case_live("bf16", N=2, C=1, H=256, W=256, K=2, seed=3, floor=True)),
}


CASES = {
    "reader_augmented": case_reader_augmented,
    "checkpoint_roundtrip": case_checkpoint_roundtrip,
    "estimate_radius_c1": lambda: case_estimate_radius(1, 2, 41),
    "estimate_radius_c3": lambda: case_estimate_radius(3, 4, 42),
    "train_driver": case_train_driver,
    "tiled_inference_bf16": lambda: case_tiled_inference("bf16"),
    "tiled_inference_fp32": lambda: case_tiled_inference("fp32"),
    "live_fp32_c1k2": lambda: case_live("fp32", N=2, C=1, H=64, W=48, K=2, seed=21),
    "live_fp32_c3k8": lambda: case_live("fp32", N=1, C=3, H=80, W=112, K=8, seed=22, gb=4),
    "live_bf16_c1k2": lambda: case_live("bf16", N=2, C=1, H=96, W=64, K=2, seed=23, floor=True),
    "live_bf16_c3k8": lambda: case_live("bf16", N=1, C=3, H=112, W=144, K=8, seed=24, gb=8, floor=True),
    # number_classes > 8: the class-per-lane head kernels inside the whole graph (fp32 check mode <= 1e-4, bf16 at the storage floor)
    # label_smoothing (UNet/model.py:65, :77)
    "live_fp32_c1k2_smooth": lambda: case_live("fp32", N=2, C=1, H=64, W=48, K=2, seed=28, smooth=0.1),
    "live_fp32_c6k3": lambda: case_live("fp32", N=1, C=6, H=48, W=64, K=3, seed=29, gb=2),
    "live_fp32_c3k20": lambda: case_live("fp32", N=1, C=3, H=64, W=80, K=20, seed=26, gb=2),
    "live_bf16_c1k40": lambda: case_live("bf16", N=1, C=1, H=96, W=96, K=40, seed=27, gb=2, floor=True),
    "golden_c1_k2_fp32": lambda: case_golden("graph_c1_k2", "fp32"),
    "golden_c3_k8_fp32": lambda: case_golden("graph_c3_k8", "fp32"),
    "golden_c1_k2_bf16": lambda: case_golden("graph_c1_k2", "bf16"),
    "golden_c3_k8_bf16": lambda: case_golden("graph_c3_k8", "bf16"),
    "curve_bf16": lambda: case_curve("bf16", steps=200),          # north_star: loss curve over the first 200 steps
    "inference_fp32": lambda: case_inference("fp32"),
    "inference_bf16": lambda: case_inference("bf16"),
    # bf16 product path at a shape with well-conditioned BatchNorm statistics: absolute bounds (also what smoke() runs)
    "wellcond_bf16_n2_256": lambda: case_live("bf16", N=2, C=1, H=256, W=256, K=2, seed=3, absolute=BF16_BOUNDS_256),
    # ... and at trained weights: north_star's 1e-2 / 99.9 %
    "trained_bf16": case_trained,
    # graph replay across buffer reallocation (two alternating input shapes) == eager launches, bit for bit
    "graph_two_shapes": case_graph_two_shapes,
    # the reference's data/ fixture -> build_lmdb -> 60 training steps, loss curve vs the oracle's
    "config1_refdata": case_config1_refdata,
    # sharded inference (rows uploaded per rank, split z-score, in-place tile reader) == whole-image path on one rank
    "sharded_inference": case_banded_inference,
    # the y-materialising training forward (UB_FOLD_BN=0; the default folds BatchNorm into the consumer convolutions)
    "nofold_live_bf16_c1k2": _with_fold(lambda: case_live("bf16", N=2, C=1, H=96, W=64, K=2, seed=23, floor=True)),
    "nofold_live_bf16_c3k8": _with_fold(lambda: case_live("bf16", N=1, C=3, H=112, W=144, K=8, seed=24, gb=8, floor=True)),
    "nofold_wellcond_bf16_n2_256": _with_fold(lambda: case_live("bf16", N=2, C=1, H=256, W=256, K=2, seed=3, absolute=BF16_BOUNDS_256)),
    "nofold_golden_c1_k2_bf16": _with_fold(lambda: case_golden("graph_c1_k2", "bf16")),
    # optional schedule: BatchNorm-backward sums from the consumer's weight gradient (UB_BN_ALGEBRA=1)
    "algebra_live_bf16_c1k2": _with_env(lambda: case_live("bf16", N=2, C=1, H=96, W=64, K=2, seed=23, floor=True), UB_BN_ALGEBRA="1"),
    "algebra_wellcond_bf16_n2_256": _with_env(lambda: case_live("bf16", N=2, C=1, H=256, W=256, K=2, seed=3, absolute=BF16_BOUNDS_256), UB_BN_ALGEBRA="1"),
    "algebra_curve_bf16": _with_env(lambda: case_curve("bf16", steps=100), UB_BN_ALGEBRA="1"),
    "nofold_curve_bf16": _with_fold(lambda: case_curve("bf16", steps=100)),
    "step_pipeline": case_step_pipeline,
    "reddeconv_wellcond_bf16_n2_256": _with_env(lambda: case_live("bf16", N=2, C=1, H=256, W=256, K=2, seed=3, absolute=BF16_BOUNDS_256), UB_FUSE_RED_DECONV="1"),
    # BatchNorm-backward sums of enc1a / dec1a inside the 64 -> 64 row-streaming dgrads (csrc/conv3_rows.cuh, RED = 2) are the default;
    # the separate reduction pass stays covered
    "nored64_wellcond_bf16_n2_256": _with_env(lambda: case_live("bf16", N=2, C=1, H=256, W=256, K=2, seed=3, absolute=BF16_BOUNDS_256), UB_FUSE_RED64="0"),
}
