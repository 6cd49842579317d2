"""Mask files (unetb200/tiffio.py): BigTIFF, 1024 x 1024 deflate tiles, as UNet/inference.py:221-222 asks of skimage."""
import numpy as np
import pytest

from unetb200 import tiffio


@pytest.mark.parametrize("dtype,shape", [(np.uint8, (1500, 2100)), (np.uint16, (1024, 1024)), (np.int32, (33, 47)), (np.uint8, (2048, 1030))])
def test_tiled_bigtiff_is_read_back_by_libtiff(tmp_path, dtype, shape):
    rng = np.random.default_rng(0)
    hi = min(np.iinfo(dtype).max, 70000)
    mask = rng.integers(0, hi, size=shape).astype(dtype)
    mask[:200, :300] = 1                                              # compressible region
    p = str(tmp_path / "mask.tif")
    tiffio.write_tiled_bigtiff(p, mask)
    st = tiffio.read_tiff_structure(p)
    assert st["bigtiff"] and st[256] == [shape[1]] and st[257] == [shape[0]]
    assert st[259] == [8] and st[322] == [1024] and st[323] == [1024]                  # deflate, 1024 x 1024 tiles
    ntiles = -(-shape[0] // 1024) * -(-shape[1] // 1024)
    assert len(st[324]) == len(st[325]) == ntiles
    assert st[258] == [8 * np.dtype(dtype).itemsize] and st[339] == [1 if np.dtype(dtype).kind == "u" else 2]
    from PIL import Image
    Image.MAX_IMAGE_PIXELS = None
    back = np.asarray(Image.open(p))
    assert back.shape == mask.shape and np.array_equal(back.astype(np.int64), mask.astype(np.int64))
    if dtype != np.int32:                                             # OpenCV has no int32 TIFF reader
        cv2 = pytest.importorskip("cv2")
        back2 = cv2.imread(p, cv2.IMREAD_UNCHANGED)
        assert back2 is not None and np.array_equal(back2, mask)


def test_inference_imsave_uses_the_reference_container(tmp_path):
    from unetb200.inference import imsave, imread
    mask = (np.arange(1200 * 1100).reshape(1200, 1100) % 3).astype(np.uint8)
    p = str(tmp_path / "m.tif")
    imsave(p, mask, "tif")
    st = tiffio.read_tiff_structure(p)
    assert st["bigtiff"] and st[322] == [1024] and st[259] == [8]
    assert np.array_equal(imread(p), mask)
    q = str(tmp_path / "m.png")
    imsave(q, mask, "png")
    assert np.array_equal(imread(q), mask)
