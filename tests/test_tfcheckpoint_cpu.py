"""TensorBundle checkpoint files (unetb200/tfcheckpoint.py): the files `tf.train.Checkpoint.write` leaves behind
(UNet/train.py:96, :181-184), written and read without TensorFlow.  CPU only: the library is loaded for its host CRC-32C."""
import os
import struct

import numpy as np
import pytest

from unetb200 import tfcheckpoint as T

BLOCKS = ["enc1a", "enc1b", "enc2a", "enc2b", "enc3a", "enc3b", "enc4a", "enc4b", "bota", "botb"]
KINDS = ["first"] + ["conv"] * 9
for lvl in (4, 3, 2, 1):
    BLOCKS += [f"up{lvl}", f"dec{lvl}a", f"dec{lvl}b"]
    KINDS += ["deconv", "conv", "conv"]
BLOCKS.append("head")
KINDS.append("head")


def test_crc32c_known_answers():
    assert T.crc32c(b"123456789") == 0xE3069283                     # the CRC-32C check value
    assert T.crc32c(b"") == 0
    assert T.crc32c(bytes(32)) == 0x8A9136AA                        # RFC 3720 B.4: 32 bytes of zeros
    assert T.crc32c(bytes([0xff] * 32)) == 0x62A8AB43               # RFC 3720 B.4: 32 bytes of ones
    assert T.crc32c(bytes(range(32))) == 0x46DD794E                 # RFC 3720 B.4: incrementing bytes
    b = os.urandom(4099)
    assert T.crc32c(b[1000:], T.crc32c(b[:1000])) == T.crc32c(b)    # continuation
    for c in (0, 1, 0xdeadbeef, 0xffffffff):
        assert T.unmask_crc(T.mask_crc(c)) == c
    pw = pytest.importorskip("tensorboard.compat.tensorflow_stub.pywrap_tensorflow")
    for n in (1, 7, 8, 9, 63, 64, 65, 1000):                         # independent pure-Python implementation
        b = os.urandom(n)
        assert T.crc32c(b) == pw.crc32c(b)
        assert T.mask_crc(T.crc32c(b)) == pw.masked_crc32c(b)


def test_table_round_trip_multi_block(tmp_path):
    rng = np.random.default_rng(0)
    items = [(b"", b"header")]
    for i in range(3000):
        items.append((f"key/{i:05d}/suffix".encode(), rng.bytes(int(rng.integers(0, 300)))))
    p = str(tmp_path / "t.index")
    T.write_table(p, items, block_size=4096)
    back = T.read_table(p)
    assert list(back.items()) == items
    raw = open(p, "rb").read()
    assert struct.unpack("<Q", raw[-8:])[0] == T.TABLE_MAGIC and len(raw) > 48
    with pytest.raises(ValueError):
        T.write_table(p, [(b"b", b""), (b"a", b"")])
    # a flipped byte inside a data block is caught by the block checksum
    bad = bytearray(raw)
    bad[100] ^= 0x40
    open(p, "wb").write(bad)
    with pytest.raises(IOError):
        T.read_table(p)


def test_separator_rules():
    assert T._shortest_separator(b"abcdefg", b"abzzz") == b"abd"
    assert T._shortest_separator(b"abc", b"abcd") == b"abc"           # prefix: unchanged
    assert T._shortest_separator(b"ab\xff", b"ac") == b"ab\xff"
    assert T._short_successor(b"\xff\xffa") == b"\xff\xffb"
    assert T._short_successor(b"model/x") == b"n"


def test_bundle_round_trip(tmp_path):
    rng = np.random.default_rng(1)
    tensors = {
        "a/float": rng.normal(size=(3, 3, 4, 8)).astype(np.float32),
        "a/scalar": np.asarray(7, dtype=np.int64),
        "b/empty": np.zeros((0, 5), dtype=np.float32),
        "b/u8": rng.integers(0, 255, size=(17,), dtype=np.uint8),
        T.OBJECT_GRAPH_KEY: b"\x0a\x00 some bytes \xff",
    }
    prefix = str(tmp_path / "ckpt")
    T.write_bundle(prefix, tensors)
    assert sorted(os.listdir(tmp_path)) == ["ckpt.data-00000-of-00001", "ckpt.index"]     # what checkpoint.write leaves behind
    back = T.read_bundle(prefix)
    assert sorted(back) == sorted(tensors)
    for k, v in tensors.items():
        if isinstance(v, bytes):
            assert back[k] == v
        else:
            assert back[k].dtype == v.dtype and back[k].shape == v.shape and np.array_equal(back[k], v)
    # header entry: one shard, little endian, bundle version 1
    idx = T.read_table(prefix + ".index")
    assert idx[b""] == bytes([0x08, 0x01, 0x1a, 0x02, 0x08, 0x01])
    # entry layout: float tensor -> dtype 1, dims, size = 4 * numel, data stored in key order
    e = T._parse_entry(idx[b"a/float"])
    assert e["dtype"] == 1 and e["shape"] == (3, 3, 4, 8) and e["size"] == 4 * 288
    assert T._parse_entry(idx[b"a/scalar"])["shape"] == () and T._parse_entry(idx[b"a/scalar"])["dtype"] == 9
    # corrupt one data byte -> checksum error
    data = bytearray(open(prefix + ".data-00000-of-00001", "rb").read())
    data[e["offset"] + 5] ^= 1
    open(prefix + ".data-00000-of-00001", "wb").write(data)
    with pytest.raises(IOError):
        T.read_bundle(prefix)


def test_string_tensor_layout():
    payload, crc = T._string_payload([b"hello"])
    assert payload[0] == 5 and payload[5:] == b"hello" and len(payload) == 1 + 4 + 5
    lc = struct.unpack("<I", payload[1:5])[0]
    assert lc == T.mask_crc(T.crc32c(struct.pack("<I", 5)))           # checksum of the lengths as uint32
    assert crc == T.crc32c(b"hello", T.crc32c(payload[1:5], T.crc32c(struct.pack("<I", 5))))


def test_variable_keys_follow_the_keras_layer_order():
    keys = T.variable_keys(BLOCKS, KINDS)
    names = list(keys)
    assert names[0] == "model/layer_with_weights-0/kernel" and keys[names[0]] == ("enc1a", "kernel")
    assert keys["model/layer_with_weights-1/moving_variance"] == ("enc1a", "moving_variance")
    assert keys["model/layer_with_weights-2/bias"] == ("enc1b", "bias")
    assert keys["model/layer_with_weights-20/kernel"] == ("up4", "kernel")        # 10 encoder/bottleneck blocks x (conv, bn)
    assert keys["model/layer_with_weights-45/gamma"] == ("head", "gamma")
    assert len(keys) == 23 * 6                                                     # 23 blocks x (kernel, bias, gamma, beta, mean, var)
    plan = T.keras_layer_plan(BLOCKS, KINDS)
    roles = [r for _, r, _ in plan]
    assert roles.count("pool") == 4 and roles.count("dropout") == 2 and roles.count("concat") == 4
    assert [n for n, _, _ in plan][:6] == ["input_1", "conv2d", "batch_normalization", "conv2d_1", "batch_normalization_1", "max_pooling2d"]
    i = [n for n, _, _ in plan].index("conv2d_7")                                  # enc4b -> dropout -> pool (model.py:104-106)
    assert roles[i:i + 4] == ["conv", "bn", "dropout", "pool"]


def test_object_graph_parses_with_an_independent_decoder():
    pb = pytest.importorskip("tensorboard.compat.proto.trackable_object_graph_pb2")
    g = pb.TrackableObjectGraph()
    g.ParseFromString(T.object_graph(BLOCKS, KINDS))
    root = g.nodes[0]
    assert [c.local_name for c in root.children] == ["model", "optimizer"]
    model = g.nodes[root.children[0].node_id]
    opt = g.nodes[root.children[1].node_id]
    by_name = {c.local_name: c.node_id for c in model.children}
    assert by_name["layer_with_weights-0"] == by_name["layer-1"]                  # layer-0 is the InputLayer
    assert "layer_with_weights-45" in by_name and "layer_with_weights-46" not in by_name
    conv0 = g.nodes[by_name["layer_with_weights-0"]]
    assert [c.local_name for c in conv0.children] == ["kernel", "bias"]
    kern = g.nodes[conv0.children[0].node_id]
    assert kern.attributes[0].name == "VARIABLE_VALUE" and kern.attributes[0].full_name == "conv2d/kernel"
    assert kern.attributes[0].checkpoint_key == "model/layer_with_weights-0/kernel/.ATTRIBUTES/VARIABLE_VALUE"
    assert sorted(c.local_name for c in opt.children) == ["beta_1", "beta_2", "decay", "iter", "learning_rate"]
    assert len(opt.slot_variables) == 2 * 23 * 4                                   # m and v for kernel, bias, gamma, beta of 23 blocks
    s0 = opt.slot_variables[0]
    assert s0.slot_name == "m" and s0.original_variable_node_id == conv0.children[0].node_id
    assert g.nodes[s0.slot_variable_node_id].attributes[0].checkpoint_key == \
        "model/layer_with_weights-0/kernel/.OPTIMIZER_SLOT/optimizer/m/.ATTRIBUTES/VARIABLE_VALUE"
    # every checkpoint key named in the graph is unique
    keys = [a.checkpoint_key for n in g.nodes for a in n.attributes]
    assert len(keys) == len(set(keys)) == 23 * 6 + 5 + 2 * 23 * 4


def _fake_params(rng, nc=1, K=2, b=4):
    """TF-layout parameter dict of a narrow U-Net (base width b) -- the file format does not care about widths"""
    chans = {"enc1a": (nc, b), "enc1b": (b, b), "enc2a": (b, 2 * b), "enc2b": (2 * b, 2 * b), "enc3a": (2 * b, 4 * b), "enc3b": (4 * b, 4 * b),
             "enc4a": (4 * b, 8 * b), "enc4b": (8 * b, 8 * b), "bota": (8 * b, 16 * b), "botb": (16 * b, 16 * b), "head": (b, K)}
    for lvl in (4, 3, 2, 1):
        c = b << (lvl - 1)
        chans[f"up{lvl}"] = (2 * c, c)
        chans[f"dec{lvl}a"] = (2 * c, c)
        chans[f"dec{lvl}b"] = (c, c)
    p = {}
    for n, kind in zip(BLOCKS, KINDS):
        ci, co = chans[n]
        shape = (2, 2, co, ci) if kind == "deconv" else ((1, 1, ci, co) if kind == "head" else (3, 3, ci, co))
        p[n + "/kernel"] = rng.normal(size=shape).astype(np.float32)
        for part in ("bias", "gamma", "beta", "moving_mean", "moving_var"):
            p[f"{n}/{part}"] = rng.normal(size=(co,)).astype(np.float32)
    return p


def test_unet_checkpoint_round_trip(tmp_path):
    rng = np.random.default_rng(2)
    p, m, v = _fake_params(rng), _fake_params(rng), _fake_params(rng)
    prefix = str(tmp_path / "checkpoint" / "ckpt")
    os.makedirs(os.path.dirname(prefix))
    T.save_unet(prefix, BLOCKS, KINDS, p, m, v, step=1234, learning_rate=3e-4)
    ck = T.load_unet(prefix, BLOCKS, KINDS)
    assert ck["step"] == 1234 and abs(ck["learning_rate"] - 3e-4) < 1e-10
    for k in p:
        assert np.array_equal(ck["params"][k], p[k]), k
        if "moving" not in k:
            assert np.array_equal(ck["adam_m"][k], m[k]) and np.array_equal(ck["adam_v"][k], v[k])
    bundle = T.read_bundle(prefix)
    assert bundle["optimizer/iter/.ATTRIBUTES/VARIABLE_VALUE"].dtype == np.int64
    assert bundle["model/layer_with_weights-20/kernel/.ATTRIBUTES/VARIABLE_VALUE"].shape == p["up4/kernel"].shape
    assert len(bundle) == 1 + 5 + 23 * 6 + 2 * 23 * 4
    # weights-only checkpoint (what expect_partial tolerates): optimizer state comes back as None
    T.save_unet(prefix, BLOCKS, KINDS, p)
    ck = T.load_unet(prefix, BLOCKS, KINDS)
    assert ck["adam_m"] is None and np.array_equal(ck["params"]["head/kernel"], p["head/kernel"])
    # a checkpoint that lacks model variables is an error, not a silent partial load
    t = {k: v for k, v in T.read_bundle(prefix).items() if "layer_with_weights-3/" not in k}
    T.write_bundle(prefix, t)
    with pytest.raises(IOError):
        T.load_unet(prefix, BLOCKS, KINDS)
