"""Training-time augmentation of the reference (UNet/augment.py), run on the GPU on whole raw-pixel batches.

The reference augments each example on the host inside its reader processes (UNet/imagereader.py:283-294): float32
conversion, `skimage.transform.rotate` + `.warp` (bilinear, mirror boundary) on image and mask, additive Gaussian noise
scaled by the image's dynamic range, `scipy.ndimage.gaussian_filter`, optional intensity shift, mask rounding.  At
B200 step rates (~700 images/s per GPU) that host path cannot feed one GPU, so the random PARAMETERS are drawn on the
host in the reference's order (`draw_params`) and the pixels are processed by `csrc/augment.cu` between the H2D copy of
the raw batch and the z-score: two bilinear warps (rotation; translate/scale with the flips folded into the index map),
min/max, Philox noise, separable blur.  There is no CPU fallback: `augment_image` below keeps the reference's signature
and runs the same kernels on a batch of one.
"""
from __future__ import annotations

import numpy as np
import torch

from . import _C

MAX_RADIUS = 32
_DT = {torch.uint8: 0, torch.int16: 1, torch.float32: 2}


def draw_params(rng, n, h, w, rotation_flag=False, reflection_flag=False, jitter_augmentation_severity=0, noise_augmentation_severity=0,
                scale_augmentation_severity=0, blur_augmentation_max_sigma=0, intensity_augmentation_severity=0):
    """The scalar draws of augment_image (UNet/augment.py:61-153), in the reference's order, for n examples.
    rng: numpy RandomState-like (`.rand()`); the reference uses the global np.random."""
    jit, noi, sca = jitter_augmentation_severity or 0, noise_augmentation_severity or 0, scale_augmentation_severity or 0
    blu, inten = blur_augmentation_max_sigma or 0, intensity_augmentation_severity or 0
    assert 0 <= jit < 1 and 0 <= noi < 1 and 0 <= sca < 1 and 0 <= inten < 1          # augment.py:48-52
    p = dict(orientation=np.full(n, np.nan), reflect_x=np.zeros(n, bool), reflect_y=np.zeros(n, bool), jitter_x=np.zeros(n, np.int64),
             jitter_y=np.zeros(n, np.int64), scale_x=np.ones(n), scale_y=np.ones(n), noise_factor=np.zeros(n), blur_sigma=np.zeros(n),
             shift_factor=np.zeros(n))
    for i in range(n):
        if rotation_flag:
            p["orientation"][i] = 360 * rng.rand()
        if reflection_flag:
            p["reflect_x"][i] = rng.rand() > 0.5
            p["reflect_y"][i] = rng.rand() > 0.5
        if jit > 0:
            jx = int(jit * (w * rng.rand()))
            if rng.rand() > 0.5:
                jx = -jx
            jy = int(jit * (h * rng.rand()))
            if rng.rand() > 0.5:
                jy = -jy
            p["jitter_x"][i], p["jitter_y"][i] = jx, jy
        if sca > 0:
            p["scale_x"][i] = (1 - sca) + 2 * sca * rng.rand()
            p["scale_y"][i] = (1 - sca) + 2 * sca * rng.rand()
        if noi > 0:
            p["noise_factor"][i] = noi * (2 * rng.rand() - 1)             # sigma = U(-s_max, s_max), s_max = noi * range (:118-127)
        if blu > 0:
            p["blur_sigma"][i] = max(0.0, blu * (2 * rng.rand() - 1))      # negative draws mean "no blur" (:130-139)
        if inten > 0:
            v = rng.rand() * inten
            p["shift_factor"][i] = v if rng.rand() > 0.5 else -v
    return p


def warp_matrices(p, h, w):
    """-> (rotation inverse maps [n,6] or None, affine inverse maps with the flips folded in [n,6]); rows (m0..m5) of
    src_x = m0 x + m1 y + m2, src_y = m3 x + m4 y + m5 (skimage's (col, row) convention)"""
    n = len(p["scale_x"])
    rot = None
    if not np.all(np.isnan(p["orientation"])):
        rot = np.zeros((n, 6))
        cx, cy = w / 2.0 - 0.5, h / 2.0 - 0.5
        for i in range(n):
            a = np.deg2rad(0.0 if np.isnan(p["orientation"][i]) else p["orientation"][i])
            R = np.array([[np.cos(a), -np.sin(a), 0], [np.sin(a), np.cos(a), 0], [0, 0, 1]])
            M = np.array([[1, 0, cx], [0, 1, cy], [0, 0, 1.0]]) @ R @ np.array([[1, 0, -cx], [0, 1, -cy], [0, 0, 1.0]])
            rot[i] = M[:2].reshape(-1)
    aff = np.zeros((n, 6))
    for i in range(n):
        inv = np.linalg.inv(np.array([[p["scale_x"][i], 0, p["jitter_x"][i]], [0, p["scale_y"][i], p["jitter_y"][i]], [0, 0, 1.0]]))
        F = np.eye(3)
        if p["reflect_x"][i]:              # np.fliplr after the warp == sampling the warp at the mirrored column
            F = F @ np.array([[-1, 0, w - 1], [0, 1, 0], [0, 0, 1.0]])
        if p["reflect_y"][i]:
            F = F @ np.array([[1, 0, 0], [0, -1, h - 1], [0, 0, 1.0]])
        aff[i] = (inv @ F)[:2].reshape(-1)
    return rot, aff


def gaussian_taps(sigma, truncate=4.0):
    """scipy.ndimage _gaussian_kernel1d: radius = int(truncate * sigma + 0.5), normalised exp(-x^2 / 2 sigma^2); -> (radius, w[0..MAX_RADIUS])"""
    w = np.zeros(MAX_RADIUS + 1)
    if sigma <= 0:
        w[0] = 1.0
        return 0, w
    radius = int(truncate * float(sigma) + 0.5)
    if radius > MAX_RADIUS:
        raise ValueError(f"blur sigma {sigma} needs radius {radius} > {MAX_RADIUS}")
    x = np.arange(-radius, radius + 1)
    phi = np.exp(-0.5 / (sigma * sigma) * x ** 2)
    phi = phi / phi.sum()
    w[:radius + 1] = phi[radius:]
    return radius, w


def channel_mix(radius, w, C):
    """the 1-D filter along an axis of length C with scipy's 'reflect' (edge-repeating mirror) boundary, as a C x C matrix"""
    M = np.zeros((C, C))
    for c in range(C):
        for k in range(-radius, radius + 1):
            j = c + k
            while j < 0 or j >= C:
                j = -j - 1 if j < 0 else 2 * C - 1 - j
            M[c, j] += w[abs(k)]
    return M


class DeviceAugmenter:
    """raw [N,C,H,W] uint8 / uint16-as-int16 / float32 device batch + uint8 [N,H,W] labels -> float32 batch + uint8 labels"""

    def __init__(self, device, seed=0):
        self.device = torch.device(device)
        self.seed = int(seed)
        self.calls = 0
        self._buf = {}

    def _b(self, name, numel, dtype):
        t = self._buf.get(name)
        if t is None or t.numel() < numel or t.dtype != dtype:
            t = torch.empty(int(numel), dtype=dtype, device=self.device)
            self._buf[name] = t
        return t[:numel]

    def _up(self, a, dtype):
        return torch.as_tensor(np.ascontiguousarray(a, dtype=dtype)).to(self.device, non_blocking=True)

    def __call__(self, raw, labels, p):
        N, C, H, W = raw.shape
        st = torch.cuda.current_stream(self.device).cuda_stream
        src_code = _DT.get(raw.dtype)
        if src_code is None:
            raise IOError(f"unsupported pixel dtype {raw.dtype}")
        raw = raw.contiguous()
        rot, aff = warp_matrices(p, H, W)
        out = torch.empty((N, C, H, W), dtype=torch.float32, device=self.device)
        tmp = self._b("tmp", N * C * H * W, torch.float32).view(N, C, H, W)
        aff_d = self._up(aff, np.float64)
        if rot is not None:
            _C.call("ub_aug_warp", raw, src_code, tmp, 2, self._up(rot, np.float64), N, C, H, W, st)
            _C.call("ub_aug_warp", tmp, 2, out, 2, aff_d, N, C, H, W, st)
        else:
            _C.call("ub_aug_warp", raw, src_code, out, 2, aff_d, N, C, H, W, st)
        lab_out = None
        if labels is not None:
            labels = labels.contiguous()
            lab_out = torch.empty((N, H, W), dtype=torch.uint8, device=self.device)
            if rot is not None:
                ltmp = self._b("ltmp", N * H * W, torch.float32)
                _C.call("ub_aug_warp", labels, 0, ltmp, 2, self._up(rot, np.float64), N, 1, H, W, st)
                _C.call("ub_aug_warp", ltmp, 2, lab_out, 0, aff_d, N, 1, H, W, st)
            else:
                _C.call("ub_aug_warp", labels, 0, lab_out, 0, aff_d, N, 1, H, W, st)
        per = C * H * W
        mm = self._b("minmax", N * 64 * 2, torch.float32)
        if np.any(p["noise_factor"] != 0):
            fac = np.stack([p["noise_factor"], np.zeros(N)], axis=1)
            _C.call("ub_aug_minmax", out, mm, N, per, st)
            _C.call("ub_aug_noise", out, mm, self._up(fac, np.float32), N, per, self.seed, self.calls * (1 << 32), st)
        if np.any(p["blur_sigma"] > 0):
            taps = [gaussian_taps(s) for s in p["blur_sigma"]]
            rad = self._up([t[0] for t in taps], np.int32)
            wts = self._up(np.stack([t[1] for t in taps]), np.float64)
            _C.call("ub_aug_blur_axis", out, tmp, wts, rad, 0, N, C, H, W, st)
            _C.call("ub_aug_blur_axis", tmp, out, wts, rad, 1, N, C, H, W, st)
            if C > 1:
                mix = np.stack([channel_mix(t[0], t[1], C) for t in taps])
                _C.call("ub_aug_chanmix", out, self._up(mix, np.float64), N, C, H * W, st)
        if np.any(p["shift_factor"] != 0):
            fac = np.stack([np.zeros(N), p["shift_factor"]], axis=1)
            _C.call("ub_aug_minmax", out, mm, N, per, st)
            _C.call("ub_aug_noise", out, mm, self._up(fac, np.float32), N, per, self.seed, 0, st)
        self.calls += 1
        return out, lab_out


_default = None


def augment_image(img, mask=None, rotation_flag=False, reflection_flag=False, jitter_augmentation_severity=0, noise_augmentation_severity=0,
                  scale_augmentation_severity=0, blur_augmentation_max_sigma=0, intensity_augmentation_severity=0):
    """UNet/augment.py:19 -- same arguments and return values ([H,W,C] float32 image, rounded float32 mask), one example."""
    global _default
    img = np.asarray(img, dtype=np.float32)
    assert img.ndim == 3
    h, w, c = img.shape
    if mask is not None:
        mask = np.asarray(mask)
        assert mask.shape[0] == h and mask.shape[1] == w
    if _default is None:
        _default = DeviceAugmenter(torch.device("cuda", torch.cuda.current_device()), seed=int(np.random.randint(0, 2 ** 31 - 1)))
    p = draw_params(np.random, 1, h, w, rotation_flag, reflection_flag, jitter_augmentation_severity, noise_augmentation_severity,
                    scale_augmentation_severity, blur_augmentation_max_sigma, intensity_augmentation_severity)
    raw = torch.as_tensor(np.ascontiguousarray(img.transpose(2, 0, 1)[None])).to(_default.device)
    lab = torch.as_tensor(np.ascontiguousarray(mask, dtype=np.uint8)[None]).to(_default.device) if mask is not None else None
    out, lab_out = _default(raw, lab, p)
    out = out[0].permute(1, 2, 0).contiguous().cpu().numpy()
    if mask is not None:
        return out, lab_out[0].cpu().numpy().astype(np.float32)
    return out
