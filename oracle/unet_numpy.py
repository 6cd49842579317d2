"""Second, independent derivation of the U-Net train step -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Hand-written forward AND backward formulas in numpy, NHWC layout, no autograd.
It exists to pin oracle/unet_oracle.py (torch autograd): tests/test_oracle.py requires the two to agree
in fp64 on loss, logits and every one of the 92 gradients.  It is also the blueprint of the GPU kernel
schedule (same dataflow: conv -> relu -> stats -> bn-apply, bn-bwd reduce/apply, dgrad, wgrad ...).

Reference anchors: UNet/model.py:28-63 (layer recipes), :85-146 (wiring), :204-221 (loss + grads).
PARITY UNPINNED against TensorFlow itself -- see oracle/unet_oracle.py header.
"""
from __future__ import annotations

import numpy as np

BN_EPS = 1e-3


def conv_fwd(x, w_hwio, b):
    """x [N,H,W,Ci], w [k,k,Ci,Co] cross-correlation, zero 'same' padding."""
    k = w_hwio.shape[0]
    r = k // 2
    N, H, W, _ = x.shape
    xp = np.pad(x, ((0, 0), (r, r), (r, r), (0, 0)))
    out = np.zeros((N, H, W, w_hwio.shape[3]), dtype=x.dtype)
    for dy in range(k):
        for dx in range(k):
            out += xp[:, dy:dy + H, dx:dx + W, :] @ w_hwio[dy, dx]
    return out + b


def conv_dgrad(dz, w_hwio):
    k = w_hwio.shape[0]
    r = k // 2
    N, H, W, _ = dz.shape
    dzp = np.pad(dz, ((0, 0), (r, r), (r, r), (0, 0)))
    dx = np.zeros((N, H, W, w_hwio.shape[2]), dtype=dz.dtype)
    for dy in range(k):
        for dxx in range(k):
            # x[p + (dy-r, dx-r)] contributed to z[p]  =>  dx[q] += dz[q - (dy-r, dx-r)] W[dy,dx]^T
            dx += dzp[:, 2 * r - dy:2 * r - dy + H, 2 * r - dxx:2 * r - dxx + W, :] @ w_hwio[dy, dxx].T
    return dx


def conv_wgrad(x, dz, k):
    r = k // 2
    N, H, W, Ci = x.shape
    xp = np.pad(x, ((0, 0), (r, r), (r, r), (0, 0)))
    dw = np.zeros((k, k, Ci, dz.shape[3]), dtype=x.dtype)
    d2 = dz.reshape(-1, dz.shape[3])
    for dy in range(k):
        for dx in range(k):
            dw[dy, dx] = xp[:, dy:dy + H, dx:dx + W, :].reshape(-1, Ci).T @ d2
    return dw, d2.sum(0)


def deconv_fwd(x, w, b):
    """w [2,2,Co,Ci]; out[n,2i+a,2j+c,co] = b + sum_ci x[n,i,j,ci] w[a,c,co,ci]."""
    N, h, wd, _ = x.shape
    Co = w.shape[2]
    out = np.zeros((N, 2 * h, 2 * wd, Co), dtype=x.dtype)
    for a in range(2):
        for c in range(2):
            out[:, a::2, c::2, :] = x @ w[a, c].T
    return out + b


def deconv_bwd(x, dz, w):
    dx = np.zeros_like(x)
    dw = np.zeros_like(w)
    for a in range(2):
        for c in range(2):
            d = dz[:, a::2, c::2, :]
            dx += d @ w[a, c]
            dw[a, c] = d.reshape(-1, d.shape[3]).T @ x.reshape(-1, x.shape[3])
    return dx, dw, dz.reshape(-1, dz.shape[3]).sum(0)


def bn_fwd(a, gamma, beta):
    mu = a.mean(axis=(0, 1, 2))
    var = a.var(axis=(0, 1, 2))
    rstd = 1.0 / np.sqrt(var + BN_EPS)
    return (a - mu) * rstd * gamma + beta, mu, rstd


def bn_bwd(dy, a, mu, rstd, gamma):
    M = a.shape[0] * a.shape[1] * a.shape[2]
    xh = (a - mu) * rstd
    dbeta = dy.sum(axis=(0, 1, 2))
    dgamma = (dy * xh).sum(axis=(0, 1, 2))
    da = gamma * rstd * (dy - dbeta / M - xh * dgamma / M)
    return da, dgamma, dbeta


def pool_fwd(y):
    N, H, W, C = y.shape
    v = y.reshape(N, H // 2, 2, W // 2, 2, C).transpose(0, 1, 3, 5, 2, 4).reshape(N, H // 2, W // 2, C, 4)
    idx = v.argmax(-1)                         # first max on ties; slot = 2*dy + dx
    return np.take_along_axis(v, idx[..., None], -1)[..., 0], idx


def pool_bwd(dp, idx):
    N, h, w, C = dp.shape
    d = np.zeros((N, h, w, C, 4), dtype=dp.dtype)
    np.put_along_axis(d, idx[..., None], dp[..., None], -1)
    return d.reshape(N, h, w, C, 2, 2).transpose(0, 1, 4, 2, 5, 3).reshape(N, 2 * h, 2 * w, C)


class ManualUNet:
    """Forward with caches + explicit backward.  Params use the TF layouts of oracle.unet_oracle.init_params
    (as numpy arrays)."""

    def __init__(self, params):
        self.p = {k: (v.numpy() if hasattr(v, "numpy") else np.asarray(v)) for k, v in params.items()}
        self.c = {}

    # -- blocks -------------------------------------------------------------------------------
    def _conv(self, name, x):
        p = self.p
        a = np.maximum(conv_fwd(x, p[name + "/kernel"], p[name + "/bias"]), 0)
        y, mu, rstd = bn_fwd(a, p[name + "/gamma"], p[name + "/beta"])
        self.c[name] = (x, a, mu, rstd)
        return y

    def _conv_bwd(self, name, dy, need_dx=True):
        p, g = self.p, self.g
        x, a, mu, rstd = self.c[name]
        da, g[name + "/gamma"], g[name + "/beta"] = bn_bwd(dy, a, mu, rstd, p[name + "/gamma"])
        dz = da * (a > 0)
        g[name + "/kernel"], g[name + "/bias"] = conv_wgrad(x, dz, p[name + "/kernel"].shape[0])
        return conv_dgrad(dz, p[name + "/kernel"]) if need_dx else None

    def _deconv(self, name, x):
        p = self.p
        z = deconv_fwd(x, p[name + "/kernel"], p[name + "/bias"])
        y, mu, rstd = bn_fwd(z, p[name + "/gamma"], p[name + "/beta"])
        self.c[name] = (x, z, mu, rstd)
        return y

    def _deconv_bwd(self, name, dy):
        p, g = self.p, self.g
        x, z, mu, rstd = self.c[name]
        dz, g[name + "/gamma"], g[name + "/beta"] = bn_bwd(dy, z, mu, rstd, p[name + "/gamma"])
        dx, g[name + "/kernel"], g[name + "/bias"] = deconv_bwd(x, dz, p[name + "/kernel"])
        return dx

    # -- graph --------------------------------------------------------------------------------
    def forward(self, x_nchw, dropout_masks=None):
        dm = dropout_masks or {}
        x = np.ascontiguousarray(np.transpose(x_nchw, (0, 2, 3, 1)))
        self.pool_idx = {}
        self.drop = {}
        skips = {}
        cur = x
        for lvl in (1, 2, 3, 4):
            cur = self._conv(f"enc{lvl}b", self._conv(f"enc{lvl}a", cur))
            if lvl == 4 and "drop4" in dm:
                self.drop["drop4"] = np.transpose(np.asarray(dm["drop4"]), (0, 2, 3, 1)).astype(cur.dtype) * 2.0
                cur = cur * self.drop["drop4"]
            skips[lvl] = cur
            cur, self.pool_idx[lvl] = pool_fwd(cur)
        cur = self._conv("botb", self._conv("bota", cur))
        if "dropb" in dm:
            self.drop["dropb"] = np.transpose(np.asarray(dm["dropb"]), (0, 2, 3, 1)).astype(cur.dtype) * 2.0
            cur = cur * self.drop["dropb"]
        for lvl in (4, 3, 2, 1):
            u = self._deconv(f"up{lvl}", cur)
            cur = self._conv(f"dec{lvl}b", self._conv(f"dec{lvl}a", np.concatenate([skips[lvl], u], axis=-1)))
        self.logits = self._conv("head", cur)
        e = np.exp(self.logits - self.logits.max(-1, keepdims=True))
        self.softmax = e / e.sum(-1, keepdims=True)
        return self.softmax

    def loss(self, labels_onehot, global_batch_size):
        t = np.asarray(labels_onehot, dtype=self.logits.dtype)
        z = self.logits - self.logits.max(-1, keepdims=True)
        lsm = z - np.log(np.exp(z).sum(-1, keepdims=True))
        _, H, W, _ = t.shape
        self.t = t
        self.denom = global_batch_size * H * W
        return -(t * lsm).sum() / self.denom

    def backward(self):
        self.g = {}
        d = (self.softmax - self.t) / self.denom          # dL/dlogits (App. A.7)
        d = self._conv_bwd("head", d)
        dskip = {}
        for lvl in (1, 2, 3, 4):
            d = self._conv_bwd(f"dec{lvl}a", self._conv_bwd(f"dec{lvl}b", d))
            C = d.shape[-1] // 2
            dskip[lvl] = d[..., :C]
            d = self._deconv_bwd(f"up{lvl}", d[..., C:])
        if "dropb" in self.drop:
            d = d * self.drop["dropb"]
        d = self._conv_bwd("bota", self._conv_bwd("botb", d))
        for lvl in (4, 3, 2, 1):
            d = pool_bwd(d, self.pool_idx[lvl]) + dskip[lvl]      # skip fan-out gradient sum (App. E)
            if lvl == 4 and "drop4" in self.drop:
                d = d * self.drop["drop4"]
            d = self._conv_bwd(f"enc{lvl}a", self._conv_bwd(f"enc{lvl}b", d), need_dx=(lvl != 1))
        return self.g


# ---------------------------------------------------------------------------------------------------------------------
# BatchNorm folded into the CONSUMER convolution ("the plan for it"): y = s * a + t per input channel, the conv
# reads `a` with weights W' = W * s[ci]; because 'same' padding is applied to y (zeros), the constant t contributes only
# through the taps that fall INSIDE the image: a bias that takes 9 values per output channel (3 row cases x 3 column cases).
def fold_weights(w_hwio, s, t):
    """-> (W' [k,k,Ci,Co] = W * s[ci], Tt [k,k,Co] = sum_ci W[dy,dx,ci,co] * t[ci])"""
    return w_hwio * s[None, None, :, None], np.einsum("yxio,i->yxo", w_hwio, t)


def border_case_bias(Tt, b):
    """bias[row_case][col_case][co]: case 0 = first row/column (tap -1 is outside), 1 = interior, 2 = last row/column (tap +1
    is outside).  For a 1-pixel-high/wide image both neighbours are outside: handled by `conv_fwd_folded` via explicit masks."""
    k = Tt.shape[0]
    out = np.zeros((3, 3, Tt.shape[2]), dtype=Tt.dtype)
    for rc in range(3):
        for cc in range(3):
            rows = [d for d in range(k) if not (rc == 0 and d == 0) and not (rc == 2 and d == k - 1)]
            cols = [d for d in range(k) if not (cc == 0 and d == 0) and not (cc == 2 and d == k - 1)]
            out[rc, cc] = b + sum(Tt[dy, dx] for dy in rows for dx in cols)
    return out


def conv_fwd_folded(a, w_hwio, b, s, t):
    """== conv_fwd(s * a + t, W, b) for H, W >= 2, computed from `a` alone (what the folded kernel will do)"""
    N, H, W, _ = a.shape
    assert H >= 2 and W >= 2
    Wf, Tt = fold_weights(w_hwio, s, t)
    z = conv_fwd(a, Wf, np.zeros_like(b))
    bias = border_case_bias(Tt, b)
    rc = np.ones(H, dtype=np.int64)
    rc[0], rc[-1] = 0, 2
    cc = np.ones(W, dtype=np.int64)
    cc[0], cc[-1] = 0, 2
    return z + bias[rc[:, None], cc[None, :]][None]


def border_sums(dz):
    """Sdz[dy,dx,co] = sum of dz over the output pixels whose tap (dy-1, dx-1) neighbour lies inside the image: the total minus
    the first / last row and column strips (plus the doubly-subtracted corner)."""
    N, H, W, Co = dz.shape
    tot = dz.sum((0, 1, 2))
    row = {0: dz[:, 0].sum((0, 1)), 1: np.zeros(Co, dz.dtype), 2: dz[:, H - 1].sum((0, 1))}      # excluded strip per tap row
    col = {0: dz[:, :, 0].sum((0, 1)), 1: np.zeros(Co, dz.dtype), 2: dz[:, :, W - 1].sum((0, 1))}
    cor = {(0, 0): dz[:, 0, 0].sum(0), (0, 2): dz[:, 0, W - 1].sum(0), (2, 0): dz[:, H - 1, 0].sum(0), (2, 2): dz[:, H - 1, W - 1].sum(0)}
    out = np.zeros((3, 3, Co), dtype=dz.dtype)
    for dy in range(3):
        for dx in range(3):
            out[dy, dx] = tot - row[dy] - col[dx] + cor.get((dy, dx), 0.0)
    return out


def conv_wgrad_folded(a, dz, s, t):
    """== conv_wgrad(s * a + t, dz, 3)[0] from `a`: dW[dy,dx,ci,co] = s[ci] * dW_a[dy,dx,ci,co] + t[ci] * Sdz[dy,dx,co]"""
    dw_a, _ = conv_wgrad(a, dz, 3)
    return dw_a * s[None, None, :, None] + t[None, None, :, None] * border_sums(dz)[:, :, None, :]


def bn_bwd_sums_from_wgrad(w_hwio, dw_a, sdz, mu, rstd):
    """The backward sums of the BatchNormalization that FEEDS a convolution, without a pass over its gradient tensor:
    with dy = conv_dgrad(dz, W) (gradient w.r.t. the BatchNorm output) and xhat = (a - mu) * rstd,
        dbeta[c]  = sum_p dy[p, c]           = sum_{tap, co} W[tap, c, co] * Sdz[tap, co]
        dgamma[c] = sum_p dy[p, c] xhat[p, c] = rstd[c] * sum_{tap, co} W[tap, c, co] * (dW_a[tap, c, co] - mu[c] * Sdz[tap, co])
    where dW_a = conv_wgrad(a, dz) is the consumer's weight gradient computed on the pre-BatchNorm activation (what the folded
    training step computes anyway) and Sdz = border_sums(dz).  Exact identity (exchange the order of the two sums)."""
    dbeta = np.einsum("yxco,yxo->c", w_hwio, sdz)
    dgamma = np.einsum("yxco,yxco->c", w_hwio, dw_a - mu[None, None, :, None] * sdz[:, :, None, :]) * rstd
    return dbeta, dgamma
