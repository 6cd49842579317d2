"""CPU (-m "not gpu"): the data surface -- LMDB file format, ImageMaskPair wire format, ImageReader sampling."""
import os

import numpy as np
import pytest

import unetb200.imagereader as R
import unetb200.lmdbfile as L


def _pairs(n, h=32, w=48, c=1, dtype=np.uint16, k=3, seed=0):
    rng = np.random.default_rng(seed)
    out = []
    for i in range(n):
        img = rng.integers(0, 60000 if dtype == np.uint16 else 255, size=(h, w, c)).astype(dtype)
        mask = np.zeros((h, w), dtype=np.uint8)
        if i % 2:
            mask[: h // 2] = 1
        if i % 5 == 0:
            mask[:, : w // 4] = 2 if k > 2 else 1
        out.append((f"img{i:03d}", img, mask))
    return out


def test_lmdb_roundtrip_large_values_and_multilevel_tree(tmp_path):
    rng = np.random.default_rng(1)
    items = [(f"k{i:05d}:0,1".encode(), rng.bytes(int(rng.choice([5, 300, 2030, 2040, 9000, 70000])))) for i in range(1500)]
    L.write(str(tmp_path / "db"), items)
    with L.Reader(str(tmp_path / "db")) as r:
        assert len(r) == 1500 and r.depth >= 2 and r.psize == 4096
        assert list(r.keys()) == sorted(k for k, _ in items)
        d = dict(items)
        for k, v in r.items():
            assert d[k] == v
        for k in list(d)[::37]:
            assert r.get(k) == d[k]
        assert r.get(b"absent") is None
    # on-disk invariants of the format: two meta pages with the magic, page numbers stamped, data.mdb a whole number of pages
    raw = open(tmp_path / "db" / "data.mdb", "rb").read()
    assert len(raw) % 4096 == 0
    for pg in (0, 1):
        assert int.from_bytes(raw[pg * 4096:pg * 4096 + 8], "little") == pg
        assert int.from_bytes(raw[pg * 4096 + 16:pg * 4096 + 20], "little") == 0xBEEFC0DE
    with pytest.raises(IOError):
        L.Reader(str(tmp_path / "nope"))
    with pytest.raises(ValueError):
        L.write(str(tmp_path / "dup"), [(b"a", b"1"), (b"a", b"2")])


def test_image_mask_pair_against_the_reference_descriptor():
    """encode_pair / decode_pair against the reference's OWN message definition: the serialized descriptor embedded in
    UNet/isg_ai_pb2.py:22 (fixture tests/golden/isg_ai_descriptor.bin, extracted by tests/golden/make_descriptor.py), loaded through
    descriptor_pool.AddSerializedFile -- the records build_lmdb.py:44-59 writes and imagereader.py:269-281 parses."""
    import os
    from google.protobuf import descriptor_pool, message_factory
    raw = open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "isg_ai_descriptor.bin"), "rb").read()
    pool = descriptor_pool.DescriptorPool()
    pool.AddSerializedFile(raw)
    cls = message_factory.GetMessageClass(pool.FindMessageTypeByName("isg_ai.ImageMaskPair"))
    rng = np.random.default_rng(3)
    for img, mask in [(rng.integers(0, 65535, size=(32, 48, 1)).astype(np.uint16), rng.integers(0, 2, size=(32, 48)).astype(np.uint8)),
                      (rng.integers(0, 255, size=(16, 16, 3)).astype(np.uint8), rng.integers(0, 8, size=(16, 16)).astype(np.uint8))]:
        # the writer of the reference (build_lmdb.py:44-59)
        ref = cls()
        ref.channels, ref.img_height, ref.img_width = img.shape[2], img.shape[0], img.shape[1]
        ref.image, ref.mask = img.tobytes(), mask.tobytes()
        ref.img_type, ref.mask_type = img.dtype.str, mask.dtype.str
        ref.labels = np.unique(mask).astype(np.uint8).tobytes()
        theirs = ref.SerializeToString()
        mine = R.encode_pair(img, mask)
        assert mine == theirs                                         # byte-identical record
        d = R.decode_pair(theirs)                                     # and the reference's bytes parse with the hand-written codec
        got = np.frombuffer(d["image"], dtype=d["img_type"]).reshape(d["img_height"], d["img_width"], d["channels"])
        assert np.array_equal(got, img) and np.array_equal(np.frombuffer(d["mask"], dtype=d["mask_type"]).reshape(mask.shape), mask)
        back = cls.FromString(mine)                                   # the reader of the reference (imagereader.py:269-281)
        assert np.array_equal(np.frombuffer(back.image, dtype=back.img_type).reshape(back.img_height, back.img_width, back.channels), img)


def test_image_mask_pair_wire_format_matches_protobuf_runtime():
    """encode_pair / decode_pair against google.protobuf's own encoder for the message of UNet/isg_ai.proto:16-31
    (descriptor rebuilt here field by field; the reference's generated module is not importable under protobuf >= 4)"""
    from google.protobuf import descriptor_pb2, descriptor_pool, message_factory
    fdp = descriptor_pb2.FileDescriptorProto(name="isg_ai_test.proto", package="isg_ai_test", syntax="proto2")
    msg = fdp.message_type.add(name="ImageMaskPair")
    T = descriptor_pb2.FieldDescriptorProto
    for num, name, typ in [(1, "channels", T.TYPE_INT32), (2, "img_height", T.TYPE_INT32), (3, "img_width", T.TYPE_INT32),
                           (4, "img_type", T.TYPE_STRING), (5, "mask_type", T.TYPE_STRING), (6, "image", T.TYPE_BYTES),
                           (7, "mask", T.TYPE_BYTES), (8, "labels", T.TYPE_BYTES)]:
        msg.field.add(name=name, number=num, type=typ, label=T.LABEL_OPTIONAL)
    pool = descriptor_pool.DescriptorPool()
    pool.Add(fdp)
    cls = message_factory.GetMessageClass(pool.FindMessageTypeByName("isg_ai_test.ImageMaskPair"))
    _, img, mask = _pairs(2)[1]
    mine = R.encode_pair(img, mask)
    m = cls()
    m.ParseFromString(mine)
    assert (m.channels, m.img_height, m.img_width) == (1, 32, 48)
    assert m.img_type == "<u2" and m.mask_type == "|u1"
    assert m.image == img.tobytes() and m.mask == mask.tobytes() and m.labels == np.unique(mask).astype(np.uint8).tobytes()
    theirs = cls(channels=1, img_height=32, img_width=48, img_type="<u2", mask_type="|u1", image=img.tobytes(), mask=mask.tobytes(),
                 labels=np.unique(mask).astype(np.uint8).tobytes()).SerializeToString()
    d = R.decode_pair(theirs)
    assert (d["channels"], d["img_height"], d["img_width"], d["img_type"], d["mask_type"]) == (1, 32, 48, "<u2", "|u1")
    assert bytes(d["image"]) == img.tobytes() and bytes(d["mask"]) == mask.tobytes()
    assert theirs == mine          # same field order and encoding as the reference's writer produces


def test_record_keys_follow_build_lmdb():
    mask = np.zeros((4, 4), dtype=np.uint8)
    mask[0, 0] = 2
    assert R.record_key("a.tif", mask) == b"a.tif:0,2"
    assert R.record_key("a", mask, 512, 1024) == b"a_i512_j1024:0,2"


def test_reader_contract(tmp_path):
    db = str(tmp_path / "train.lmdb")
    R.write_database(db, _pairs(10))
    r = R.ImageReader(db, use_augmentation=False, balance_classes=True, shuffle=True, number_classes=3, seed=5)
    assert r.get_image_count() == 10 and r.get_image_size() == [32, 48, 1]
    assert r.get_image_tensor_shape() == [1, 32, 48] and r.get_label_tensor_shape() == [32, 48]
    assert [len(k) for k in r.keys] == [10, 5, 2]            # classes present per record, from the key suffix
    # balanced sampling: class drawn uniformly, so the two records holding class 2 come up ~1/3 of the time
    hits = sum(r._next_key() in r.keys[2] for _ in range(3000))
    assert 0.27 < hits / 3000 < 0.50
    I, M = r.get_example()
    assert I.dtype == np.float32 and I.shape == (1, 32, 48) and abs(float(I.mean())) < 1e-3 and abs(float(I.std()) - 1) < 1e-3
    assert M.dtype == np.int32 and M.shape == (32, 48, 3) and (M.sum(-1) == 1).all()
    x, l = r.next_raw_batch(4)
    assert tuple(x.shape) == (4, 1, 32, 48) and tuple(l.shape) == (4, 32, 48) and r.src_dtype_code() == 1
    # sequential (test) reader strides by world size from its rank
    t0 = R.ImageReader(db, use_augmentation=False, shuffle=False, number_classes=3, rank=0, world_size=2)
    t1 = R.ImageReader(db, use_augmentation=False, shuffle=False, number_classes=3, rank=1, world_size=2)
    assert [t0._next_key() for _ in range(3)] == t0.keys_flat[0:6:2] and [t1._next_key() for _ in range(3)] == t1.keys_flat[1:6:2]
    # label >= number_classes is an IndexError, as in the reference (imagereader.py:306-311)
    bad = R.ImageReader(db, use_augmentation=False, shuffle=False, number_classes=2)
    with pytest.raises(IndexError):
        for _ in range(10):
            bad.get_example()
    with pytest.raises(IOError):
        R.ImageReader(str(tmp_path / "missing"))
    R.write_database(str(tmp_path / "odd.lmdb"), _pairs(2, h=30))
    with pytest.raises(IOError):
        R.ImageReader(str(tmp_path / "odd.lmdb"))


def test_early_stopping_selection():
    from unetb200.train import select_best_epoch
    assert select_best_epoch([0.5, 0.40004, 0.4, 0.41]) == 1          # first epoch within 1e-4 of the minimum
    assert select_best_epoch([0.3]) == 0


def test_reader_draws_augmentation_with_the_reference_constants(tmp_path):
    """UNet/imagereader.py:78-85: rotation + reflection on, jitter 0.1, noise 0.02, scale 0.1, blur sigma <= 2, intensity off"""
    db = str(tmp_path / "aug.lmdb")
    R.write_database(db, _pairs(4))
    a = R.ImageReader(db, use_augmentation=True, shuffle=True, number_classes=3, seed=3)
    b = R.ImageReader(db, use_augmentation=True, shuffle=True, number_classes=3, seed=3)
    p, q = a.draw_augmentation(64), b.draw_augmentation(64)
    assert all(np.array_equal(p[k], q[k], equal_nan=True) for k in p)             # seeded: reproducible
    assert ((0 <= p["orientation"]) & (p["orientation"] < 360)).all()
    assert 5 < p["reflect_x"].sum() < 59 and 5 < p["reflect_y"].sum() < 59
    assert np.abs(p["jitter_x"]).max() <= int(0.1 * 48) and np.abs(p["jitter_y"]).max() <= int(0.1 * 32)
    assert (p["scale_x"] >= 0.9).all() and (p["scale_x"] <= 1.1).all() and (p["scale_y"] >= 0.9).all() and (p["scale_y"] <= 1.1).all()
    assert np.abs(p["noise_factor"]).max() <= 0.02 and (p["noise_factor"] < 0).any() and (p["noise_factor"] > 0).any()
    assert (p["blur_sigma"] >= 0).all() and p["blur_sigma"].max() <= 2 and (p["blur_sigma"] == 0).sum() > 10     # negative draws = no blur
    assert (p["shift_factor"] == 0).all()
    r1 = R.ImageReader(db, use_augmentation=True, shuffle=True, number_classes=3, seed=3, rank=1, world_size=2)
    assert not np.array_equal(r1.draw_augmentation(8)["orientation"], a.draw_augmentation(8)["orientation"])      # ranks draw differently
