"""CPU restatement of the reference's augmentation (UNet/augment.py) -- TEST INFRASTRUCTURE ONLY (imported by tests/).

The arithmetic lives in third-party code that is not installed here: scikit-image (`skimage.transform.rotate`, `.warp`,
`.AffineTransform`; requirements.txt leaves the version open) and SciPy (`scipy.ndimage.gaussian_filter`, installed).  The
skimage pieces are restated from their published algorithm: `warp(order=1, mode='reflect')` = bilinear interpolation in
double precision of the source pixel at `inverse_map @ (col, row, 1)` with out-of-range indices mirrored WITHOUT repeating
the edge sample (numpy.pad 'reflect'; skimage/_shared/interpolation.pxd coord_map 'R').  Pinned by tests/test_augment_cpu.py
against scipy.ndimage.map_coordinates(order=1, mode='mirror'), an independent implementation of the same sampling rule.
Parity with skimage itself is unpinned (not installable)."""
import numpy as np
import scipy.ndimage


def coord_map_reflect(dim, coord):
    """skimage coord_map(dim, coord, 'R') on an integer array"""
    coord = np.asarray(coord, dtype=np.int64)
    if dim == 1:
        return np.zeros_like(coord)
    cmax = dim - 1
    out = coord.copy()
    neg = coord < 0
    n = -coord[neg]
    out[neg] = np.where((n // cmax) % 2 != 0, cmax - (n % cmax), n % cmax)
    big = coord > cmax
    b = coord[big]
    out[big] = np.where((b // cmax) % 2 != 0, cmax - (b % cmax), b % cmax)
    return out


def warp2d(img, inv):
    """skimage.transform.warp(img, inv, order=1, mode='reflect', preserve_range=True) for a 2-D image, float64"""
    img = np.asarray(img, dtype=np.float64)
    H, W = img.shape
    yy, xx = np.meshgrid(np.arange(H, dtype=np.float64), np.arange(W, dtype=np.float64), indexing="ij")
    c = inv[0, 0] * xx + inv[0, 1] * yy + inv[0, 2]
    r = inv[1, 0] * xx + inv[1, 1] * yy + inv[1, 2]
    minr, minc = np.floor(r), np.floor(c)
    maxr, maxc = np.ceil(r), np.ceil(c)
    dr, dc = r - minr, c - minc
    r0, r1 = coord_map_reflect(H, minr.astype(np.int64)), coord_map_reflect(H, maxr.astype(np.int64))
    c0, c1 = coord_map_reflect(W, minc.astype(np.int64)), coord_map_reflect(W, maxc.astype(np.int64))
    top = (1 - dc) * img[r0, c0] + dc * img[r0, c1]
    bottom = (1 - dc) * img[r1, c0] + dc * img[r1, c1]
    return (1 - dr) * top + dr * bottom


def warp(img, inv):
    img = np.asarray(img)
    if img.ndim == 2:
        return warp2d(img, inv)
    return np.stack([warp2d(img[..., k], inv) for k in range(img.shape[2])], axis=-1)


def rotation_inverse_map(angle_deg, rows, cols):
    """skimage.transform.rotate(resize=False, center=None): tform = T(center) R(angle) T(-center) handed to warp as the inverse map"""
    cx, cy = cols / 2.0 - 0.5, rows / 2.0 - 0.5
    a = np.deg2rad(angle_deg)
    R = np.array([[np.cos(a), -np.sin(a), 0], [np.sin(a), np.cos(a), 0], [0, 0, 1]])
    T1 = np.array([[1, 0, cx], [0, 1, cy], [0, 0, 1.0]])
    T3 = np.array([[1, 0, -cx], [0, 1, -cy], [0, 0, 1.0]])
    return T1 @ R @ T3


def affine_inverse_map(jitter_x, jitter_y, scale_x, scale_y):
    """AffineTransform(translation=(jx, jy), scale=(sx, sy))._inv_matrix (UNet/augment.py:163-165)"""
    return np.linalg.inv(np.array([[scale_x, 0, jitter_x], [0, scale_y, jitter_y], [0, 0, 1.0]]))


def apply_affine_transformation(I, orientation, reflect_x, reflect_y, jitter_x, jitter_y, scale_x, scale_y):
    """UNet/augment.py:157-174"""
    if orientation is not None:
        I = warp(I, rotation_inverse_map(orientation, I.shape[0], I.shape[1]))
    I = warp(I, affine_inverse_map(jitter_x, jitter_y, scale_x, scale_y))
    if reflect_x:
        I = np.fliplr(I)
    if reflect_y:
        I = np.flipud(I)
    return I


def augment_image(img, mask, p, noise_field=None):
    """UNet/augment.py:19-155 with the random draws injected.  p: dict with orientation (None = no rotation), reflect_x,
    reflect_y, jitter_x, jitter_y, scale_x, scale_y, noise_factor (= sigma / range), blur_sigma, shift_factor (= delta / range).
    noise_field: the randn(h, w, c) array (None -> zeros)."""
    img = np.asarray(img, dtype=np.float32)
    img = apply_affine_transformation(img, p["orientation"], p["reflect_x"], p["reflect_y"], p["jitter_x"], p["jitter_y"], p["scale_x"], p["scale_y"])
    if mask is not None:
        mask = apply_affine_transformation(np.asarray(mask, dtype=np.float32), p["orientation"], p["reflect_x"], p["reflect_y"], p["jitter_x"],
                                           p["jitter_y"], p["scale_x"], p["scale_y"])
    if p.get("noise_factor", 0.0) != 0.0:
        sigma = p["noise_factor"] * (np.max(img) - np.min(img))
        if noise_field is not None:
            img = img + noise_field * sigma
    if p.get("blur_sigma", 0.0) > 0:
        img = scipy.ndimage.gaussian_filter(img, p["blur_sigma"], mode="reflect")
    if p.get("shift_factor", 0.0) != 0.0:
        img = img + p["shift_factor"] * (np.max(img) - np.min(img))
    img = np.asarray(img, dtype=np.float32)
    if mask is not None:
        return img, np.round(np.asarray(mask, dtype=np.float32))
    return img
