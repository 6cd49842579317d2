"""CPU oracle for the U-Net hot path  --  TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.  The product path (semantic-segmentation-unet_b200/) never does.

PARITY UNPINNED: the reference (usnistgov/semantic-segmentation-unet) ships no tests, golden
vectors or recorded outputs for this path, and its only arithmetic back-end (tensorflow-gpu>=2.0,
UNet/requirements.txt:2, un-vendored, un-pinned) is not installable here.  This file restates
UNet/model.py in torch-CPU (fp64 for checking, fp32 for timing) following the Keras/TF-2.x default
semantics written down in SURVEY.md Appendix A.  It is pinned only by (1) a second, independent
numpy derivation (oracle/unet_numpy.py), (2) finite differences, (3) invariants -- see
tests/test_oracle.py.

Reference anchors (file:line relative to /root/reference):
  conv block   UNet/model.py:28-37     Conv2D(same, relu, channels_first) -> BatchNormalization(axis=1)
  deconv block UNet/model.py:39-48     Conv2DTranspose(k=2,s=2,same, no act) -> BatchNormalization(axis=1)
  pool         UNet/model.py:50-53     MaxPool2D(2)
  concat       UNet/model.py:55-58     [skip, up] along channels
  dropout      UNet/model.py:60-63     rate 0.5
  wiring       UNet/model.py:85-146
  loss/optim   UNet/model.py:65-79, 204-228
  radius       UNet/model.py:160-202
"""
from __future__ import annotations

import math
from collections import OrderedDict

import numpy as np
import torch
import torch.nn.functional as F

SIZE_FACTOR = 16      # UNet/model.py:25
RADIUS = 96           # UNet/model.py:26
BN_EPS = 1e-3         # Keras BatchNormalization default
BN_MOMENTUM = 0.99
ADAM_B1, ADAM_B2, ADAM_EPS = 0.9, 0.999, 1e-7   # Keras Adam defaults (App. A.6)


def layer_specs(number_channels: int, number_classes: int, base: int = 64):
    """(name, kind, cin, cout) in forward order -- UNet/model.py:85-136."""
    b = base
    return [
        ("enc1a", "conv", number_channels, b), ("enc1b", "conv", b, b),
        ("enc2a", "conv", b, 2 * b), ("enc2b", "conv", 2 * b, 2 * b),
        ("enc3a", "conv", 2 * b, 4 * b), ("enc3b", "conv", 4 * b, 4 * b),
        ("enc4a", "conv", 4 * b, 8 * b), ("enc4b", "conv", 8 * b, 8 * b),
        ("bota", "conv", 8 * b, 16 * b), ("botb", "conv", 16 * b, 16 * b),
        ("up4", "deconv", 16 * b, 8 * b), ("dec4a", "conv", 16 * b, 8 * b), ("dec4b", "conv", 8 * b, 8 * b),
        ("up3", "deconv", 8 * b, 4 * b), ("dec3a", "conv", 8 * b, 4 * b), ("dec3b", "conv", 4 * b, 4 * b),
        ("up2", "deconv", 4 * b, 2 * b), ("dec2a", "conv", 4 * b, 2 * b), ("dec2b", "conv", 2 * b, 2 * b),
        ("up1", "deconv", 2 * b, b), ("dec1a", "conv", 2 * b, b), ("dec1b", "conv", b, b),
        ("head", "head", b, number_classes),
    ]


def init_params(number_channels, number_classes, seed=0, base=64, dtype=torch.float64,
                randomize_affine=False):
    """Keras initial state: glorot_uniform kernels, zero bias, gamma 1, beta 0, moving 0/1 (App. A.1-A.3).

    Kernels are kept in the TF layouts: conv [kh,kw,Cin,Cout], deconv [kh,kw,Cout,Cin].
    randomize_affine perturbs bias/gamma/beta so parity tests exercise them.
    """
    rng = np.random.default_rng(seed)
    p = OrderedDict()
    for name, kind, cin, cout in layer_specs(number_channels, number_classes, base):
        k = 3 if kind == "conv" else (2 if kind == "deconv" else 1)
        if kind == "deconv":
            shape = (k, k, cout, cin)
        else:
            shape = (k, k, cin, cout)
        fan_in, fan_out = k * k * cin, k * k * cout
        if kind == "deconv":  # Keras computes fans from the kernel shape [.., out, in]
            fan_in, fan_out = k * k * cout, k * k * cin
        lim = math.sqrt(6.0 / (fan_in + fan_out))
        p[name + "/kernel"] = torch.tensor(rng.uniform(-lim, lim, size=shape), dtype=dtype)
        if randomize_affine:
            p[name + "/bias"] = torch.tensor(rng.normal(0, 0.05, size=(cout,)), dtype=dtype)
            p[name + "/gamma"] = torch.tensor(rng.uniform(0.7, 1.3, size=(cout,)), dtype=dtype)
            p[name + "/beta"] = torch.tensor(rng.normal(0, 0.1, size=(cout,)), dtype=dtype)
        else:
            p[name + "/bias"] = torch.zeros(cout, dtype=dtype)
            p[name + "/gamma"] = torch.ones(cout, dtype=dtype)
            p[name + "/beta"] = torch.zeros(cout, dtype=dtype)
        p[name + "/moving_mean"] = torch.zeros(cout, dtype=dtype)
        p[name + "/moving_var"] = torch.ones(cout, dtype=dtype)
    return p


TRAINABLE_SUFFIXES = ("/kernel", "/bias", "/gamma", "/beta")


def trainable_names(params):
    return [k for k in params if k.endswith(TRAINABLE_SUFFIXES)]


class _StorageRound(torch.autograd.Function):
    """bf16-storage emulation (tests only): round the value where the CUDA path stores an activation in bf16 (forward)
    and/or where it stores that activation's gradient in bf16 (backward).  Arithmetic stays in the oracle's dtype."""

    @staticmethod
    def forward(ctx, x, fwd, bwd):
        ctx.bwd = bwd
        return x.float().to(torch.bfloat16).to(x.dtype) if fwd else x.clone()

    @staticmethod
    def backward(ctx, g):
        return (g.float().to(torch.bfloat16).to(g.dtype) if ctx.bwd else g), None, None


def _q(t, storage, fwd=True, bwd=True):
    return t if storage is None else _StorageRound.apply(t, fwd, bwd)


_FOLDED = {"enc1a", "enc2a", "enc3a", "enc4a", "bota", "dec1a", "dec2a", "dec3a", "dec4a", "up1", "up2", "up3", "up4"}


def _bn(x, name, params, training, new_stats, x_used=None, scale_out=None):
    """BatchNormalization(axis=1), App. A.3: biased batch var to normalise, unbiased into the moving avg.
    x_used (bf16-storage emulation): the stored (rounded) activation that is normalised, while the statistics come
    from the unrounded accumulator values `x` -- the data flow of the CUDA path."""
    g = params[name + "/gamma"].view(1, -1, 1, 1)
    b = params[name + "/beta"].view(1, -1, 1, 1)
    if training:
        mean = x.mean(dim=(0, 2, 3))
        var = x.var(dim=(0, 2, 3), unbiased=False)
        n = x.numel() // x.shape[1]
        if new_stats is not None:
            with torch.no_grad():
                unb = var * (n / max(n - 1, 1))
                new_stats[name + "/moving_mean"] = BN_MOMENTUM * params[name + "/moving_mean"] + (1 - BN_MOMENTUM) * mean
                new_stats[name + "/moving_var"] = BN_MOMENTUM * params[name + "/moving_var"] + (1 - BN_MOMENTUM) * unb
    else:
        mean = params[name + "/moving_mean"]
        var = params[name + "/moving_var"]
    xu = x if x_used is None else x_used
    if scale_out is not None:                      # storage="bf16_fold": the consumer folds gamma * rstd into its weights
        scale_out[name] = (torch.rsqrt(var + BN_EPS) * params[name + "/gamma"]).detach()
    return (xu - mean.view(1, -1, 1, 1)) * torch.rsqrt(var.view(1, -1, 1, 1) + BN_EPS) * g + b


def _conv_block(x, name, params, training, new_stats, taps, relu_masks=None, storage=None, w_scale=None, scales=None):
    """storage="bf16_fold" (the folded-BatchNorm forward prepared in csrc/fold.cu): a producer in _FOLDED does not store its
    BatchNorm output (no rounding of y), and its consumer rounds the weights AFTER scaling them by gamma * rstd per input channel
    (`w_scale`): conv(a, bf16(W s)) == conv(y, bf16(W s) / s) up to the shift term, which the CUDA path adds in fp32."""
    w = params[name + "/kernel"].permute(3, 2, 0, 1)          # HWIO -> OIHW  (App. A.1)
    k = w.shape[-1]
    tensor_core = storage is not None and k == 3 and w.shape[1] >= 64      # bf16 weight shadow; first layer/head stay fp32
    if tensor_core and w_scale is not None:
        sv = w_scale.view(1, -1, 1, 1)
        sv = torch.where(sv == 0, torch.ones_like(sv), sv)
        w = _q(w * sv, storage, True, False) / sv
    elif tensor_core:
        w = _q(w, storage, True, False)
    z = F.conv2d(x, w, params[name + "/bias"], padding=k // 2)
    if relu_masks is not None and name in relu_masks:
        # conditioned comparison (tests only): the activation pattern [z > 0] is taken from the implementation under
        # test, so that sign flips of pre-activations within rounding distance of 0 do not turn a 1e-7 forward
        # difference into an O(1) change of the piecewise-linear backward.  Forward values differ by <= |z| ~ rounding.
        a = z * relu_masks[name].to(z.dtype)
    else:
        a = F.relu(z)
    if taps is not None:
        taps[name + "/act"] = a
    head = (k == 1)
    if storage is None or head:                               # the head keeps fp32 activations on the CUDA path
        y = _bn(a, name, params, training, new_stats)
    elif storage == "bf16_fold" and name in _FOLDED:          # y is never stored: only its gradient is (dgrad writes bf16)
        if tensor_core:
            y = _q(_bn(_q(a, storage), name, params, training, new_stats, scale_out=scales), storage, False, True)
        else:
            y = _q(_bn(a, name, params, training, new_stats, x_used=_q(a, storage), scale_out=scales), storage, False, True)
    elif tensor_core:                                         # tcgen05 epilogue: statistics of the STORED (rounded) activations
        y = _q(_bn(_q(a, storage), name, params, training, new_stats), storage)
    else:                                                     # first layer: statistics from the fp32 values, stored rounded
        y = _q(_bn(a, name, params, training, new_stats, x_used=_q(a, storage)), storage)
    if taps is not None:
        taps[name + "/out"] = y
    return y


def _deconv_block(x, name, params, training, new_stats, taps, storage=None, scales=None):
    w = params[name + "/kernel"].permute(3, 2, 0, 1)          # [kh,kw,Cout,Cin] -> [Cin,Cout,kh,kw] (App. A.2)
    if storage is not None:
        w = _q(w, storage, True, False)
    z = F.conv_transpose2d(x, w, params[name + "/bias"], stride=2)
    if taps is not None:
        taps[name + "/act"] = z
    if storage is None:
        y = _bn(z, name, params, training, new_stats)
    elif storage == "bf16_fold":
        y = _q(_bn(_q(z, storage), name, params, training, new_stats, scale_out=scales), storage, False, True)
    else:
        y = _q(_bn(_q(z, storage), name, params, training, new_stats), storage)
    if taps is not None:
        taps[name + "/out"] = y
    return y


def _dropout(x, mask, training):
    """Dropout(0.5): keep*2 in training, identity otherwise (App. A.5). mask: 1=keep, same shape as x."""
    if not training:
        return x
    if mask is None:
        return x
    return x * mask.to(x.dtype) * 2.0


def _pool(x, idx=None):
    """MaxPool2D(2) (UNet/model.py:50-53).  idx (tests only): [N,C,H/2,W/2] window slot 2*dy+dx chosen by the
    implementation under test; routes value and gradient through that element instead of torch's own argmax."""
    if idx is None:
        return F.max_pool2d(x, 2)
    n, c, h, w = x.shape
    win = x.reshape(n, c, h // 2, 2, w // 2, 2).permute(0, 1, 2, 4, 3, 5).reshape(n, c, h // 2, w // 2, 4)
    return torch.gather(win, -1, idx.long().unsqueeze(-1)).squeeze(-1)


def forward(params, x, training, dropout_masks=None, new_stats=None, taps=None, relu_masks=None, pool_idx=None, storage=None):
    """UNet/model.py:85-146.  x: [N,C,H,W].  Returns (softmax NHWC [N,H,W,K], logits NHWC).

    `logits` is what the Softmax layer consumes: the BN output of the ReLU'd 1x1 conv.
    dropout_masks: {'drop4': [N,8b,H/8,W/8], 'dropb': [N,16b,H/16,W/16]} of {0,1}; None => no dropout.
    Test-only knobs: relu_masks / pool_idx (activation pattern of the implementation under test, see _conv_block /
    _pool) and storage="bf16" (round activations, their gradients and the tensor-core weights where the CUDA path
    stores bf16: gives the noise floor any bf16-storage implementation of this graph has against fp64).
    """
    dm = dropout_masks or {}
    pool_idx = pool_idx or {}
    st = storage

    scales = {} if st == "bf16_fold" else None

    def cb(t, name, src=None):
        """src (bf16_fold only): the producer(s) whose BatchNorm this conv folds -- a name, or (n_skip_channels, up name)"""
        ws = None
        if scales is not None and src is not None:
            if isinstance(src, tuple):
                ws = torch.cat([torch.ones(src[0], dtype=t.dtype), scales[src[1]]])
            else:
                ws = scales[src]
        return _conv_block(t, name, params, training, new_stats, taps, relu_masks, st, ws, scales)

    def fan(t):      # a tensor with two consumers: each branch's gradient is stored (rounded) before the sum
        return _q(t, st, False, True)

    c1 = cb(cb(x, "enc1a"), "enc1b", "enc1a")
    p1 = _pool(fan(c1), pool_idx.get("pool1"))
    c2 = cb(cb(p1, "enc2a"), "enc2b", "enc2a")
    p2 = _pool(fan(c2), pool_idx.get("pool2"))
    c3 = cb(cb(p2, "enc3a"), "enc3b", "enc3a")
    p3 = _pool(fan(c3), pool_idx.get("pool3"))
    c4 = cb(cb(p3, "enc4a"), "enc4b", "enc4a")
    c4 = _dropout(c4, dm.get("drop4"), training)                 # skip-4 carries the dropped tensor (Q2)
    p4 = _pool(fan(c4), pool_idx.get("pool4"))
    bt = cb(cb(p4, "bota"), "botb", "bota")
    bt = _dropout(bt, dm.get("dropb"), training)
    d = bt
    for lvl, skip in ((4, c4), (3, c3), (2, c2), (1, c1)):
        u = _deconv_block(d, f"up{lvl}", params, training, new_stats, taps, st, scales)
        cat = torch.cat([fan(skip), u], dim=1)                    # [skip, up]  UNet/model.py:117
        d = cb(cb(cat, f"dec{lvl}a", (skip.shape[1], f"up{lvl}")), f"dec{lvl}b", f"dec{lvl}a")
    logits = cb(d, "head")                                        # 1x1 + ReLU + BN (Q1)
    logits = logits.permute(0, 2, 3, 1)
    return torch.softmax(logits, dim=-1), logits


def loss_and_accuracy(logits_nhwc, labels_onehot, global_batch_size, label_smoothing=0.0):
    """UNet/model.py:211-215 + App. A.7/A.8.  labels_onehot [N,H,W,K] (any numeric dtype).
    label_smoothing (UNet/model.py:65, :77; Keras CategoricalCrossentropy): y_true * (1 - eps) + eps / K."""
    t = labels_onehot.to(logits_nhwc.dtype)
    if label_smoothing:
        t_s = t * (1.0 - label_smoothing) + label_smoothing / t.shape[-1]
    else:
        t_s = t
    ce = -(t_s * torch.log_softmax(logits_nhwc, dim=-1)).sum(-1)         # [N,H,W]
    loss = (ce.sum(0) / global_batch_size).mean()
    acc = (logits_nhwc.argmax(-1) == t.argmax(-1)).to(torch.float64).mean()
    return loss, acc


def train_step_grads(params, x, labels_onehot, global_batch_size, dropout_masks=None, taps=None, relu_masks=None,
                     pool_idx=None, storage=None, label_smoothing=0.0):
    """fwd(training=True) + loss + grads of every trainable tensor (UNet/model.py:204-221).

    Returns dict(loss, acc, softmax, logits, grads{name: tensor}, new_stats{...}).
    """
    names = trainable_names(params)
    leaves = {}
    for k, v in params.items():
        leaves[k] = v.detach().clone().requires_grad_(k in names)
    new_stats = {}
    sm, logits = forward(leaves, x, True, dropout_masks, new_stats, taps, relu_masks, pool_idx, storage)
    loss, acc = loss_and_accuracy(logits, labels_onehot, global_batch_size, label_smoothing)
    tap_keys = list(taps.keys()) if taps is not None else []
    grads = torch.autograd.grad(loss, [leaves[k] for k in names] + [taps[k] for k in tap_keys])
    out = dict(loss=loss.detach(), acc=acc, softmax=sm.detach(), logits=logits.detach(),
               grads=OrderedDict(zip(names, grads[:len(names)])), new_stats=new_stats)
    if taps is not None:      # dL/d(activation) and dL/d(BN output) of every layer, for per-layer debugging
        out["tap_grads"] = OrderedDict(zip(tap_keys, grads[len(names):]))
    return out


def test_step(params, x, labels_onehot, global_batch_size):
    """UNet/model.py:237-250: training=False (moving stats, no dropout)."""
    with torch.no_grad():
        sm, logits = forward(params, x, False)
        loss, acc = loss_and_accuracy(logits, labels_onehot, global_batch_size)
    return dict(loss=loss, acc=acc, softmax=sm, logits=logits)


class KerasAdam:
    """Keras optimizers.Adam update (App. A.6): eps added to sqrt(v) WITHOUT bias-correcting v."""

    def __init__(self, params, lr):
        self.lr = lr
        self.t = 0
        self.m = {k: torch.zeros_like(params[k]) for k in trainable_names(params)}
        self.v = {k: torch.zeros_like(params[k]) for k in trainable_names(params)}

    def apply(self, params, grads):
        self.t += 1
        lr_t = self.lr * math.sqrt(1 - ADAM_B2 ** self.t) / (1 - ADAM_B1 ** self.t)
        for k, g in grads.items():
            self.m[k] = ADAM_B1 * self.m[k] + (1 - ADAM_B1) * g
            self.v[k] = ADAM_B2 * self.v[k] + (1 - ADAM_B2) * g * g
            params[k] = params[k] - lr_t * self.m[k] / (torch.sqrt(self.v[k]) + ADAM_EPS)


def train_step(params, opt, x, labels_onehot, global_batch_size, dropout_masks=None):
    """One full optimisation step; mutates params (weights + BN moving stats)."""
    r = train_step_grads(params, x, labels_onehot, global_batch_size, dropout_masks)
    opt.apply(params, r["grads"])
    for k, v in r["new_stats"].items():
        params[k] = v
    return r


def estimate_radius(params, number_channels, noise=None, seed=0):
    """UNet/model.py:160-202 (App. A.9).  Returns (radius, grad_img)."""
    N = 2 * RADIUS
    dt = next(iter(params.values())).dtype
    if noise is None:
        noise = np.random.default_rng(seed).normal(size=(1, number_channels, N, N))
    img = torch.tensor(noise, dtype=dt, requires_grad=True)
    mid = N // 2
    sm, _ = forward(params, img, False)
    msk = sm.detach().clone()
    msk[0, mid, mid, :] = 1.0 - msk[0, mid, mid, :]
    loss = (msk - sm).abs().mean(-1).sum()       # MAE(reduction NONE) -> [1,N,N]; tape sums it
    (g,) = torch.autograd.grad(loss, img)
    grad_img = g[0].abs().numpy()
    grad_img = grad_img.mean(0) if number_channels > 1 else grad_img.squeeze()
    vec = np.maximum(grad_img.max(axis=0), grad_img.max(axis=1))
    idx = np.nonzero(vec > 1e-8)[0]
    if len(idx) < 2:
        return RADIUS, grad_img
    erf = int((idx.max() - idx.min()) / 2)
    return int(SIZE_FACTOR * math.ceil(float(erf) / SIZE_FACTOR)), grad_img


# ----------------------------------------------------------------------------------------------
# host-side numpy restatements (UNet/imagereader.py, UNet/inference.py)
# ----------------------------------------------------------------------------------------------

def zscore_normalize(image_data, channels_first=True):
    """UNet/imagereader.py:33-66.  float32, population std, std<=1 => subtract mean only."""
    image_data = np.asarray(image_data).astype(np.float32)
    if image_data.ndim == 3:
        if not channels_first:
            image_data = image_data.transpose((2, 0, 1))
        image_data = image_data.copy()
        for c in range(image_data.shape[0]):
            std = np.std(image_data[c])
            mv = np.mean(image_data[c])
            image_data[c] = (image_data[c] - mv) if std <= 1.0 else (image_data[c] - mv) / std
        if not channels_first:
            image_data = image_data.transpose((1, 2, 0))
    elif image_data.ndim == 2:
        std = np.std(image_data)
        mv = np.mean(image_data)
        image_data = (image_data - mv) if std <= 1.0 else (image_data - mv) / std
    else:
        raise IOError("Input to Z-Score normalization needs to be either a 2D or 3D image [HW, or CHW]")
    return image_data


def one_hot(mask_hw, number_classes):
    """UNet/imagereader.py:302-312: int32 [H,W] -> int32 [H,W,K]; IndexError if a label >= K."""
    m = np.asarray(mask_hw).astype(np.int32)
    out = np.zeros((m.shape[0], m.shape[1], number_classes), dtype=np.int32)
    if m.size and (m.max() >= number_classes or m.min() < 0):
        raise IndexError("label outside [0, number_classes)")
    yy, xx = np.meshgrid(np.arange(m.shape[0]), np.arange(m.shape[1]), indexing="ij")
    out[yy, xx, m] = 1
    return out


def narrow_mask_dtype(mask):
    """UNet/inference.py:215-220 (Q13): three independent ifs; max==65536 stays int32."""
    mx = int(np.max(mask)) if mask.size else 0
    if mx <= 255:
        mask = mask.astype(np.uint8)
    if 255 < mx < 65536:
        mask = mask.astype(np.uint16)
    if mx > 65536:
        mask = mask.astype(np.int32)
    return mask


def tile_plan(height, width, tile_size, radius):
    """Tile geometry of UNet/inference.py:61-95, in the reference's row-major write order.

    Yields dicts with the clamped tile box [y_st,y_end)x[x_st,x_end), the zone box and the halo to strip.
    """
    zone = tile_size - 2 * radius
    assert tile_size % SIZE_FACTOR == 0 and radius % SIZE_FACTOR == 0 and zone >= radius
    plan = []
    for i in range(0, height, zone):
        for j in range(0, width, zone):
            x_st_z, y_st_z = j, i
            x_end_z, y_end_z = j + zone, i + zone
            x_st, y_st, x_end, y_end = x_st_z - radius, y_st_z - radius, x_end_z + radius, y_end_z + radius
            pre_x = pre_y = post_x = post_y = radius
            if x_st < 0:
                x_st, pre_x = 0, 0
            if y_st < 0:
                y_st, pre_y = 0, 0
            if x_end > width:
                post_x, x_end, x_end_z = 0, width, width
            if y_end > height:
                post_y, y_end, y_end_z = 0, height, height
            plan.append(dict(y_st=y_st, y_end=y_end, x_st=x_st, x_end=x_end,
                             y_st_z=y_st_z, y_end_z=y_end_z, x_st_z=x_st_z, x_end_z=x_end_z,
                             pre_x=pre_x, pre_y=pre_y, post_x=post_x, post_y=post_y))
    return plan


def _pad_to_factor(img):
    pad_y = (SIZE_FACTOR - img.shape[0] % SIZE_FACTOR) % SIZE_FACTOR
    pad_x = (SIZE_FACTOR - img.shape[1] % SIZE_FACTOR) % SIZE_FACTOR
    if img.ndim not in (2, 3):
        raise IOError("Invalid number of dimensions for input image. Expecting HW or HWC dimension ordering.")
    if img.ndim == 2:
        img = img.reshape((img.shape[0], img.shape[1], 1))
    return img, pad_y, pad_x


def inference_tiling(img, model_fn, tile_size, radius):
    """UNet/inference.py:27-136.  model_fn: float32 [1,C,h,w] -> softmax [1,h,w,K]."""
    img, pad_y, pad_x = _pad_to_factor(img)
    if pad_x > 0 or pad_y > 0:
        img = np.pad(img, pad_width=((0, pad_y), (0, pad_x), (0, 0)), mode="reflect")
    height, width = img.shape[0], img.shape[1]
    mask = np.zeros((height, width), dtype=np.int32)
    for t in tile_plan(height, width, tile_size, radius):
        tile = img[t["y_st"]:t["y_end"], t["x_st"]:t["x_end"]]
        batch = np.ascontiguousarray(tile.transpose((2, 0, 1))[None])
        sm = np.asarray(model_fn(batch))
        sm = sm.reshape(sm.shape[-3], sm.shape[-2], sm.shape[-1])
        pred = np.argmax(sm, axis=-1).astype(np.int32)
        if t["pre_x"] > 0:
            pred = pred[:, t["pre_x"]:]
        if t["pre_y"] > 0:
            pred = pred[t["pre_y"]:, :]
        if t["post_x"] > 0:
            pred = pred[:, :-t["post_x"]]
        if t["post_y"] > 0:
            pred = pred[:-t["post_y"], :]
        mask[t["y_st_z"]:t["y_end_z"], t["x_st_z"]:t["x_end_z"]] = pred      # last writer wins (Q11)
    if pad_x > 0:
        mask = mask[:, 0:-pad_x]
    if pad_y > 0:
        mask = mask[0:-pad_y, :]
    return mask


def inference_whole(img, model_fn):
    """UNet/inference.py:139-173."""
    img, pad_y, pad_x = _pad_to_factor(img)
    img = np.pad(img, pad_width=((0, pad_y), (0, pad_x), (0, 0)), mode="reflect")
    batch = np.ascontiguousarray(img.transpose((2, 0, 1))[None])
    sm = np.asarray(model_fn(batch))
    sm = sm.reshape(sm.shape[-3], sm.shape[-2], sm.shape[-1])
    pred = np.argmax(sm, axis=-1).astype(np.int32)
    if pad_x > 0:
        pred = pred[:, 0:-pad_x]
    if pad_y > 0:
        pred = pred[0:-pad_y, :]
    return pred


def make_model_fn(params):
    dt = next(iter(params.values())).dtype

    def fn(batch):
        with torch.no_grad():
            sm, _ = forward(params, torch.as_tensor(np.asarray(batch), dtype=dt), False)
        return sm.numpy()
    return fn


def synthetic_batch(n, c, h, w, k, seed=0, dtype=np.float32):
    """Seeded synthetic (image, label-index) pair shaped like config 2: smooth noise image,
    z-scored per sample per channel; labels = thresholded / argmax'd smooth noise fields."""
    from scipy.ndimage import gaussian_filter
    rng = np.random.default_rng(seed)
    img = rng.normal(3045.0, 376.0, size=(n, c, h, w))
    img = gaussian_filter(img, sigma=(0, 0, 2, 2))
    img = np.clip(np.round(img), 0, 65535)
    x = np.stack([zscore_normalize(img[i]) for i in range(n)]).astype(dtype)
    fields = gaussian_filter(rng.normal(size=(n, k, h, w)), sigma=(0, 0, 4, 4))
    if k == 2:
        thr = np.quantile(fields[:, 1], 0.71)
        lab = (fields[:, 1] > thr).astype(np.uint8)
    else:
        lab = fields.argmax(1).astype(np.uint8)
    return x, lab
