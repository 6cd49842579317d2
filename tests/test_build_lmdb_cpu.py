"""unetb200.build_lmdb: the reference's database builder (UNet/build_lmdb.py) on synthetic image / mask folders."""
import os

import numpy as np
import pytest
from PIL import Image

import unetb200.build_lmdb as B
import unetb200.imagereader as R
from unetb200 import lmdbfile


def _folders(tmp_path, n, h, w, rng):
    os.makedirs(tmp_path / "images")
    os.makedirs(tmp_path / "masks")
    data = {}
    for i in range(n):
        img = rng.integers(0, 65535, size=(h, w)).astype(np.uint16)
        msk = (rng.random((h, w)) > 0.7).astype(np.uint8) * (1 + i % 2)
        Image.fromarray(img).save(tmp_path / "images" / f"img{i:02d}.tif")
        Image.fromarray(msk).save(tmp_path / "masks" / f"img{i:02d}.tif")
        data[f"img{i:02d}"] = (img, msk)
    return data


def test_enforce_size_multiple_crops_centred():
    a = np.arange(37 * 53).reshape(37, 53)
    c = B.enforce_size_multiple(a)
    assert c.shape == (32, 48) and c[0, 0] == a[int(5 / 2), int(5 / 2)]          # build_lmdb.py:63-80
    assert B.enforce_size_multiple(np.zeros((64, 32, 3))).shape == (64, 32, 3)


def test_tile_origins_cover_the_image_and_end_at_the_edges():
    for h, w, t in [(1024, 1024, 512), (600, 1300, 512), (512, 512, 512), (2000, 777, 256)]:
        o = B.tile_origins(h, w, t)
        cover = np.zeros((h, w), bool)
        for y, x in o:
            assert 0 <= y <= h - t and 0 <= x <= w - t
            cover[y:y + t, x:x + t] = True
        assert cover.all() and max(y for y, _ in o) == h - t and max(x for _, x in o) == w - t
        assert len(o) == len(range(0, w, t - 96)) * len(range(0, h, t - 96))      # stride = tile - RADIUS (build_lmdb.py:87)
    with pytest.raises(IOError):
        B.tile_origins(100, 600, 512)


def test_whole_image_databases(tmp_path, capsys):
    rng = np.random.default_rng(0)
    data = _folders(tmp_path, 10, 70, 100, rng)
    out = tmp_path / "out"
    B.main(str(tmp_path / "images"), str(tmp_path / "masks"), str(out), "HES", 0.8, ".tif", 0, 512, seed=3)
    assert sorted(os.listdir(out)) == ["test-HES.lmdb", "train-HES.lmdb"]
    assert "data.mdb" in os.listdir(out / "train-HES.lmdb") and "img_filenames.csv" in os.listdir(out / "train-HES.lmdb")
    train_files = open(out / "train-HES.lmdb" / "img_filenames.csv").read().split()
    test_files = open(out / "test-HES.lmdb" / "img_filenames.csv").read().split()
    assert len(train_files) == 8 and len(test_files) == 2 and sorted(train_files + test_files) == sorted(f + ".tif" for f in data)
    db = lmdbfile.Reader(str(out / "train-HES.lmdb"))
    keys = [k.decode() for k in db.keys()]
    assert len(keys) == 8
    for k in keys:
        name, classes = k.split(":")
        img, msk = data[name]
        d = R.decode_pair(db.get(k.encode()))
        assert (d["img_height"], d["img_width"], d["channels"]) == (64, 96, 1)          # cropped to multiples of 16
        assert d["img_type"] == "<u2" and d["mask_type"] == "|u1"
        assert bytes(d["image"]) == B.enforce_size_multiple(img).tobytes() and bytes(d["mask"]) == B.enforce_size_multiple(msk).tobytes()
        assert classes == ",".join(str(c) for c in np.unique(B.enforce_size_multiple(msk)))
    capsys.readouterr()
    rd = R.ImageReader(str(out / "train-HES.lmdb"), use_augmentation=False, shuffle=False, number_classes=3)
    assert rd.get_image_count() == 8 and rd.get_image_size() == [64, 96, 1]
    # rebuilding replaces the databases
    B.main(str(tmp_path / "images"), str(tmp_path / "masks"), str(out), "HES", 0.5, "tif", 0, 512, seed=4)
    assert len(list(lmdbfile.Reader(str(out / "train-HES.lmdb")).keys())) == 5
    with pytest.raises(RuntimeError):
        B.main(str(tmp_path / "images"), str(tmp_path / "masks"), str(out), "HES", 0.8, "png", 0, 512)


def test_tiled_databases(tmp_path, capsys):
    rng = np.random.default_rng(1)
    data = _folders(tmp_path, 2, 300, 420, rng)
    out = tmp_path / "out"
    B.main(str(tmp_path / "images"), str(tmp_path / "masks"), str(out), "T", 1.0, "tif", 1, 128, seed=0)
    db = lmdbfile.Reader(str(out / "train-T.lmdb"))
    keys = [k.decode() for k in db.keys()]
    per_image = len(set(B.tile_origins(300, 420, 128)))          # slid-back tiles repeat; a repeated key overwrites (txn.put)
    assert len(keys) == 2 * per_image
    k = next(k for k in keys if k.startswith("img00_i172_j292:"))                    # the slid-back corner tile
    d = R.decode_pair(db.get(k.encode()))
    assert bytes(d["image"]) == data["img00"][0][172:300, 292:420].tobytes() and d["img_height"] == d["img_width"] == 128
    with pytest.raises(AssertionError):
        B.main(str(tmp_path / "images"), str(tmp_path / "masks"), str(out), "T", 1.0, "tif", 1, 100)
    capsys.readouterr()
