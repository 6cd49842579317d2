"""Host side of the device augmentation (unetb200/augment.py) and the pinning of its oracle (oracle/augment_oracle.py)."""
import numpy as np
import pytest
import scipy.ndimage

from oracle import augment_oracle as AO
from unetb200 import augment as UA


def _coords(M, H, W):
    yy, xx = np.meshgrid(np.arange(float(H)), np.arange(float(W)), indexing="ij")
    return M[1, 0] * xx + M[1, 1] * yy + M[1, 2], M[0, 0] * xx + M[0, 1] * yy + M[0, 2]


@pytest.mark.parametrize("M", [AO.rotation_inverse_map(33.3, 37, 53), AO.rotation_inverse_map(271.0, 37, 53),
                               AO.affine_inverse_map(-5, 3, 0.9, 1.1), AO.affine_inverse_map(200, -333, 0.31, 0.27)])
def test_oracle_warp_matches_an_independent_mirror_interpolator(M):
    """skimage's bilinear + coord_map('R') restated == scipy map_coordinates(order=1, mode='mirror'), incl. many wraps"""
    img = np.random.default_rng(0).normal(size=(37, 53))
    r, c = _coords(M, 37, 53)
    ref = scipy.ndimage.map_coordinates(img, [r, c], order=1, mode="mirror")
    assert np.abs(AO.warp2d(img, M) - ref).max() < 1e-12


def test_coord_map_reflect_is_numpy_pad_reflect():
    for dim in (2, 3, 7):
        base = np.arange(dim)
        padded = np.pad(base, (3 * dim, 3 * dim), mode="reflect")
        idx = np.arange(-3 * dim, dim + 3 * dim)
        assert np.array_equal(AO.coord_map_reflect(dim, idx), padded)
    assert np.array_equal(AO.coord_map_reflect(1, np.array([-4, 0, 9])), [0, 0, 0])


class ScriptedRng:
    def __init__(self, vals):
        self.vals, self.i = list(vals), 0

    def rand(self):
        v = self.vals[self.i]
        self.i += 1
        return v


def test_draw_params_follows_the_reference_order():
    # UNet/augment.py:61-153: orientation, reflect_x, reflect_y, (jitter_x, sign), (jitter_y, sign), scale_x, scale_y, noise, blur
    vals = [0.25, 0.6, 0.4, 0.5, 0.7, 0.9, 0.2, 0.0, 1.0 - 1e-12, 0.75, 0.1]
    rng = ScriptedRng(vals)
    p = UA.draw_params(rng, 1, 100, 200, True, True, 0.1, 0.02, 0.1, 2, None)
    assert rng.i == len(vals)
    assert p["orientation"][0] == 90.0 and p["reflect_x"][0] and not p["reflect_y"][0]
    assert p["jitter_x"][0] == -int(0.1 * 200 * 0.5) and p["jitter_y"][0] == int(0.1 * 100 * 0.9)
    assert abs(p["scale_x"][0] - 0.9) < 1e-12 and abs(p["scale_y"][0] - 1.1) < 1e-9
    assert abs(p["noise_factor"][0] - 0.01) < 1e-12 and p["blur_sigma"][0] == 0.0          # negative sigma draw = no blur
    # everything off: no draws, identity parameters
    rng = ScriptedRng([])
    p = UA.draw_params(rng, 2, 64, 64)
    rot, aff = UA.warp_matrices(p, 64, 64)
    assert rot is None and np.allclose(aff, [[1, 0, 0, 0, 1, 0]] * 2)
    with pytest.raises(AssertionError):
        UA.draw_params(ScriptedRng([0.5] * 20), 1, 8, 8, jitter_augmentation_severity=1.0)


def test_flips_fold_into_the_second_warp():
    rng = np.random.default_rng(1)
    img = rng.normal(size=(24, 40))
    for rx in (False, True):
        for ry in (False, True):
            p = dict(orientation=np.array([np.nan]), reflect_x=np.array([rx]), reflect_y=np.array([ry]), jitter_x=np.array([3]),
                     jitter_y=np.array([-2]), scale_x=np.array([1.07]), scale_y=np.array([0.94]))
            _, aff = UA.warp_matrices(p, 24, 40)
            M = np.vstack([aff[0].reshape(2, 3), [0, 0, 1]])
            ref = AO.apply_affine_transformation(img, None, rx, ry, 3, -2, 1.07, 0.94)
            assert np.abs(AO.warp2d(img, M) - ref).max() < 1e-10
    p["orientation"] = np.array([17.0])
    rot, _ = UA.warp_matrices(p, 24, 40)
    assert np.allclose(np.vstack([rot[0].reshape(2, 3), [0, 0, 1]]), AO.rotation_inverse_map(17.0, 24, 40))


@pytest.mark.parametrize("sigma", [0.3, 1.0, 1.3, 2.0])
def test_gaussian_taps_and_channel_mix_match_scipy(sigma):
    radius, w = UA.gaussian_taps(sigma)
    delta = np.zeros(101)
    delta[50] = 1.0
    k = scipy.ndimage.gaussian_filter1d(delta, sigma, mode="reflect")
    assert radius == int(4 * sigma + 0.5)
    assert np.abs(k[50:50 + radius + 1] - w[:radius + 1]).max() < 1e-15 and k[50 + radius + 1] == 0
    for C in (1, 2, 3, 4):
        x = np.random.default_rng(C).normal(size=(5, C))
        ref = scipy.ndimage.gaussian_filter1d(x, sigma, axis=1, mode="reflect")
        assert np.abs(x @ UA.channel_mix(radius, w, C).T - ref).max() < 1e-12
    assert UA.gaussian_taps(0.0)[0] == 0
