"""TensorFlow checkpoint files (`tf.train.Checkpoint(optimizer=..., model=...).write(prefix)`, UNet/train.py:96, :181-184;
read back by `checkpoint.read(...).expect_partial()`, UNet/model.py:81-83) without TensorFlow.

What the reference leaves on disk is a TensorBundle:
  <prefix>.index                  an SSTable (the LevelDB table format TensorFlow vendors in core/lib/io): key "" ->
                                  BundleHeaderProto, every tensor key -> BundleEntryProto (dtype, shape, offset, size, crc32c)
  <prefix>.data-00000-of-00001    the raw little-endian tensor bytes, back to back
and the keys follow the TF2 object graph: `model/layer_with_weights-<i>/kernel/.ATTRIBUTES/VARIABLE_VALUE`,
`optimizer/iter/...`, Adam slots `<variable>/.OPTIMIZER_SLOT/optimizer/{m,v}/...`, plus the serialized
`TrackableObjectGraph` under `_CHECKPOINTABLE_OBJECT_GRAPH`.

TensorFlow is not installable in this environment, so this module is written from the published file formats and is
validated against itself, against an independent protobuf decoder for the object graph and against the CRC-32C test
vector (tests/test_tfcheckpoint_cpu.py); byte-compatibility with a checkpoint produced by TensorFlow is UNVERIFIED.
Host-side only: nothing here is on the training or inference hot path.
"""
from __future__ import annotations

import re
import struct
from collections import OrderedDict

import numpy as np

from . import _C

TABLE_MAGIC = 0xdb4775248b80fb57
BLOCK_SIZE = 262144               # tensorflow/core/lib/io/table_options.h default
RESTART_INTERVAL = 16
MASK_DELTA = 0xa282ead8
HEADER_KEY = b""
OBJECT_GRAPH_KEY = "_CHECKPOINTABLE_OBJECT_GRAPH"
ATTR = "/.ATTRIBUTES/VARIABLE_VALUE"

# tensorflow DataType enum values <-> numpy
_DT = {1: np.float32, 2: np.float64, 3: np.int32, 4: np.uint8, 5: np.int16, 6: np.int8, 9: np.int64, 10: np.bool_, 17: np.uint16,
       19: np.float16, 22: np.uint32, 23: np.uint64}
_DT_OF = {np.dtype(v): k for k, v in _DT.items()}
DT_STRING = 7


# ------------------------------------------------------------------------------------------------ checksums, varints
def crc32c(data, crc=0):
    a = np.frombuffer(data, dtype=np.uint8) if not isinstance(data, np.ndarray) else data.reshape(-1).view(np.uint8)
    if a.size == 0:
        return crc
    return int(_C.lib.ub_host_crc32c(a.ctypes.data, a.size, crc))


def mask_crc(crc):
    return (((crc >> 15) | (crc << 17)) + MASK_DELTA) & 0xffffffff


def unmask_crc(masked):
    rot = (masked - MASK_DELTA) & 0xffffffff
    return ((rot >> 17) | (rot << 15)) & 0xffffffff


def _varint(n):
    out = bytearray()
    while True:
        b = n & 0x7f
        n >>= 7
        if n:
            out.append(b | 0x80)
        else:
            out.append(b)
            return bytes(out)


def _read_varint(buf, pos):
    shift = val = 0
    while True:
        b = buf[pos]
        pos += 1
        val |= (b & 0x7f) << shift
        if not b & 0x80:
            return val, pos
        shift += 7


# ------------------------------------------------------------------------------------------------ protobuf (wire level)
def _pb_varint(field, v):
    return _varint(field << 3) + _varint(v)


def _pb_bytes(field, b):
    return _varint(field << 3 | 2) + _varint(len(b)) + b


def _pb_fixed32(field, v):
    return _varint(field << 3 | 5) + struct.pack("<I", v)


def _pb_parse(buf):
    """-> {field: [values]}; varint -> int, length-delimited -> bytes, fixed32/64 -> int"""
    out, pos = {}, 0
    while pos < len(buf):
        tag, pos = _read_varint(buf, pos)
        field, wt = tag >> 3, tag & 7
        if wt == 0:
            v, pos = _read_varint(buf, pos)
        elif wt == 2:
            n, pos = _read_varint(buf, pos)
            v = bytes(buf[pos:pos + n])
            pos += n
        elif wt == 5:
            v = struct.unpack_from("<I", buf, pos)[0]
            pos += 4
        elif wt == 1:
            v = struct.unpack_from("<Q", buf, pos)[0]
            pos += 8
        else:
            raise IOError(f"unsupported protobuf wire type {wt}")
        out.setdefault(field, []).append(v)
    return out


def _header_proto():
    # BundleHeaderProto{num_shards = 1; endianness = LITTLE (0, omitted); version = VersionDef{producer = 1}}
    return _pb_varint(1, 1) + _pb_bytes(3, _pb_varint(1, 1))


def _entry_proto(dtype, shape, offset, size, crc_masked):
    # BundleEntryProto{dtype = 1; shape = 2; shard_id = 3 (0, omitted); offset = 4; size = 5; crc32c = 6 (fixed32)}
    dims = b"".join(_pb_bytes(2, _pb_varint(1, int(d)) if d else b"") for d in shape)       # TensorShapeProto.dim{size}
    out = _pb_varint(1, dtype) + _pb_bytes(2, dims)
    if offset:
        out += _pb_varint(4, offset)
    if size:
        out += _pb_varint(5, size)
    return out + _pb_fixed32(6, crc_masked)


def _parse_entry(buf):
    f = _pb_parse(buf)
    shape = []
    for d in _pb_parse(f.get(2, [b""])[0]).get(2, []):
        shape.append(_pb_parse(d).get(1, [0])[0])
    return dict(dtype=f.get(1, [0])[0], shape=tuple(shape), shard=f.get(3, [0])[0], offset=f.get(4, [0])[0], size=f.get(5, [0])[0],
                crc=f.get(6, [0])[0])


# ------------------------------------------------------------------------------------------------ table (SSTable) files
class _BlockBuilder:
    def __init__(self, restart_interval):
        self.interval = restart_interval
        self.buf = bytearray()
        self.restarts = [0]
        self.counter = 0
        self.last = b""

    def add(self, key, value):
        shared = 0
        if self.counter < self.interval:
            n = min(len(self.last), len(key))
            while shared < n and self.last[shared] == key[shared]:
                shared += 1
        else:
            self.restarts.append(len(self.buf))
            self.counter = 0
        self.buf += _varint(shared) + _varint(len(key) - shared) + _varint(len(value)) + key[shared:] + value
        self.last = key
        self.counter += 1

    def size_estimate(self):
        return len(self.buf) + 4 * len(self.restarts) + 4

    def empty(self):
        return not self.buf

    def finish(self):
        return bytes(self.buf) + b"".join(struct.pack("<I", r) for r in self.restarts) + struct.pack("<I", len(self.restarts))


def _shortest_separator(start, limit):
    n = min(len(start), len(limit))
    i = 0
    while i < n and start[i] == limit[i]:
        i += 1
    if i < n and start[i] < 0xff and start[i] + 1 < limit[i]:
        return start[:i] + bytes([start[i] + 1])
    return start


def _short_successor(key):
    for i, b in enumerate(key):
        if b != 0xff:
            return key[:i] + bytes([b + 1])
    return key


def _handle(offset, size):
    return _varint(offset) + _varint(size)


def write_table(path, items, block_size=BLOCK_SIZE):
    """items: (key bytes, value bytes) in strictly increasing key order"""
    with open(path, "wb") as f:
        pos = 0

        def write_block(contents):
            nonlocal pos
            trailer = b"\x00" + struct.pack("<I", mask_crc(crc32c(contents + b"\x00")))     # kNoCompression
            f.write(contents + trailer)
            h = (pos, len(contents))
            pos += len(contents) + 5
            return h

        data, index = _BlockBuilder(RESTART_INTERVAL), _BlockBuilder(1)
        pending = None
        last_key = None
        for key, value in items:
            if last_key is not None and not key > last_key:
                raise ValueError("table keys must be strictly increasing")
            if pending is not None:
                index.add(_shortest_separator(last_key, key), _handle(*pending))
                pending = None
            data.add(key, value)
            last_key = key
            if data.size_estimate() >= block_size:
                pending = write_block(data.finish())
                data = _BlockBuilder(RESTART_INTERVAL)
        if not data.empty():
            pending = write_block(data.finish())
        meta = write_block(_BlockBuilder(RESTART_INTERVAL).finish())
        if pending is not None:
            index.add(_short_successor(last_key), _handle(*pending))
        idx = write_block(index.finish())
        footer = _handle(*meta) + _handle(*idx)
        footer += b"\x00" * (40 - len(footer)) + struct.pack("<Q", TABLE_MAGIC)
        f.write(footer)


def _block_entries(block):
    nrestart = struct.unpack_from("<I", block, len(block) - 4)[0]
    end = len(block) - 4 * (nrestart + 1)
    pos, key = 0, b""
    while pos < end:
        shared, pos = _read_varint(block, pos)
        non_shared, pos = _read_varint(block, pos)
        vlen, pos = _read_varint(block, pos)
        key = key[:shared] + bytes(block[pos:pos + non_shared])
        pos += non_shared
        yield key, bytes(block[pos:pos + vlen])
        pos += vlen


def read_table(path):
    """-> OrderedDict key -> value of every entry (block checksums verified)"""
    buf = open(path, "rb").read()
    if len(buf) < 48 or struct.unpack_from("<Q", buf, len(buf) - 8)[0] != TABLE_MAGIC:
        raise IOError(f"{path}: not a TensorFlow table file (bad magic)")
    footer = buf[-48:]
    _, p = _read_varint(footer, 0)
    _, p = _read_varint(footer, p)
    ioff, p = _read_varint(footer, p)
    isize, p = _read_varint(footer, p)

    def block(off, size):
        contents, ctype = buf[off:off + size], buf[off + size]
        stored = struct.unpack_from("<I", buf, off + size + 1)[0]
        if unmask_crc(stored) != crc32c(buf[off:off + size + 1]):
            raise IOError(f"{path}: block checksum mismatch at offset {off}")
        if ctype != 0:
            raise IOError(f"{path}: compressed table blocks (type {ctype}) are not supported")
        return contents

    out = OrderedDict()
    for _, h in _block_entries(block(ioff, isize)):
        off, p = _read_varint(h, 0)
        size, p = _read_varint(h, p)
        for k, v in _block_entries(block(off, size)):
            out[k] = v
    return out


# ------------------------------------------------------------------------------------------------ tensor bundle
def _string_payload(strings):
    """DT_STRING on-disk form: [varint64 len]* [fixed32 masked crc of the lengths] [bytes]*; -> (payload, entry crc unmasked)"""
    lengths = b"".join(_varint(len(s)) for s in strings)
    crc = 0
    for s in strings:
        crc = crc32c(struct.pack("<I", len(s)) if len(s) <= 0xffffffff else struct.pack("<Q", len(s)), crc)
    lc = struct.pack("<I", mask_crc(crc))
    crc = crc32c(lc, crc)
    for s in strings:
        crc = crc32c(s, crc)
    return lengths + lc + b"".join(strings), crc


def write_bundle(prefix, tensors):
    """tensors: {key: numpy array | bytes (a scalar DT_STRING)} -> <prefix>.index, <prefix>.data-00000-of-00001"""
    items = [(HEADER_KEY, _header_proto())]
    offset = 0
    with open(prefix + ".data-00000-of-00001", "wb") as f:
        for key in sorted(tensors, key=lambda k: k.encode()):
            t = tensors[key]
            if isinstance(t, (bytes, bytearray)):
                payload, crc = _string_payload([bytes(t)])
                dtype, shape = DT_STRING, ()
            else:
                shape0 = np.shape(t)
                a = np.ascontiguousarray(t)          # (0-d arrays come back 1-d: the shape is taken before)
                if a.dtype.byteorder == ">":
                    a = a.astype(a.dtype.newbyteorder("<"))
                if a.dtype not in _DT_OF:
                    raise TypeError(f"{key}: dtype {a.dtype} has no TensorFlow DataType mapping here")
                payload, crc = a.reshape(-1).view(np.uint8), crc32c(a)
                dtype, shape = _DT_OF[a.dtype], shape0
            f.write(payload)
            size = len(payload)
            items.append((key.encode(), _entry_proto(dtype, shape, offset, size, mask_crc(crc))))
            offset += size
    write_table(prefix + ".index", items)


def read_bundle(prefix, verify=True):
    """-> OrderedDict key -> numpy array (bytes for a scalar DT_STRING)"""
    index = read_table(prefix + ".index")
    hdr = _pb_parse(index.get(HEADER_KEY, b""))
    if hdr.get(1, [1])[0] != 1:
        raise IOError("multi-shard TensorBundles are not supported (the reference writes one shard)")
    if hdr.get(2, [0])[0] != 0:
        raise IOError("big-endian TensorBundle")
    data = np.fromfile(prefix + ".data-00000-of-00001", dtype=np.uint8)
    out = OrderedDict()
    for k, v in index.items():
        if k == HEADER_KEY:
            continue
        e = _parse_entry(v)
        raw = data[e["offset"]:e["offset"] + e["size"]]
        if len(raw) != e["size"]:
            raise IOError(f"{k.decode()}: data file is truncated")
        if e["dtype"] == DT_STRING:
            n = int(np.prod(e["shape"])) if e["shape"] else 1
            b = raw.tobytes()
            pos, lens = 0, []
            for _ in range(n):
                ln, pos = _read_varint(b, pos)
                lens.append(ln)
            pos += 4
            strs = []
            for ln in lens:
                strs.append(b[pos:pos + ln])
                pos += ln
            if verify and mask_crc(_string_payload(strs)[1]) != e["crc"]:
                raise IOError(f"{k.decode()}: checksum mismatch")
            out[k.decode()] = strs[0] if not e["shape"] else strs
            continue
        if e["dtype"] not in _DT:
            raise IOError(f"{k.decode()}: unsupported DataType {e['dtype']}")
        if verify and mask_crc(crc32c(raw)) != e["crc"]:
            raise IOError(f"{k.decode()}: checksum mismatch")
        out[k.decode()] = raw.view(_DT[e["dtype"]]).reshape(e["shape"]).copy()
    return out


# ------------------------------------------------------------------------------------------------ the U-Net's object graph
_PARTS = {"conv": ("kernel", "bias"), "bn": ("gamma", "beta", "moving_mean", "moving_variance")}


def keras_layer_plan(block_names, kinds):
    """The Keras layer list of UNet._build_model (UNet/model.py:85-146) in `model.layers` order.
    block_names / kinds: this package's weighted blocks in forward order (kind 'first'|'conv'|'deconv'|'head').
    -> [(keras_name, role, block)] with role in {'input','conv','bn','pool','dropout','concat','permute','softmax'}"""
    plan = [("input_1", "input", None)]
    counts = {}

    def nm(base):
        i = counts.get(base, 0)
        counts[base] = i + 1
        return base if i == 0 else f"{base}_{i}"

    for b, kind in zip(block_names, kinds):
        plan.append((nm("conv2d_transpose" if kind == "deconv" else "conv2d"), "conv", b))
        plan.append((nm("batch_normalization"), "bn", b))
        if kind == "deconv":
            plan.append((nm("concatenate"), "concat", None))
        if b in ("enc4b", "botb"):
            plan.append((nm("dropout"), "dropout", None))
        if b in ("enc1b", "enc2b", "enc3b", "enc4b"):
            plan.append((nm("max_pooling2d"), "pool", None))
    plan += [("permute", "permute", None), ("softmax", "softmax", None)]
    return plan


def variable_keys(block_names, kinds):
    """-> OrderedDict checkpoint key (without the /.ATTRIBUTES suffix) -> (block, part) for every model variable"""
    out = OrderedDict()
    wi = 0
    for _, role, b in keras_layer_plan(block_names, kinds):
        if role in _PARTS:
            for part in _PARTS[role]:
                out[f"model/layer_with_weights-{wi}/{part}"] = (b, part)
            wi += 1
    return out


def object_graph(block_names, kinds):
    """Serialized TrackableObjectGraph of Checkpoint(optimizer=Adam, model=<functional Keras model>): breadth-first node
    numbering from the root, children of the model named layer_with_weights-<i> / layer-<j>, Adam slot references."""
    nodes = []          # [children [(local_name, id)], attributes [(name, full_name, key)], slots [(orig, slot, id)]]

    def new():
        nodes.append(([], [], []))
        return len(nodes) - 1

    root = new()
    model, opt = new(), new()
    nodes[root][0].extend([("model", model), ("optimizer", opt)])            # Checkpoint sorts its keyword arguments
    plan = keras_layer_plan(block_names, kinds)
    layer_ids = [new() for _ in plan]
    wi = 0
    for li, (kname, role, _) in enumerate(plan):
        if role in _PARTS:
            nodes[model][0].append((f"layer_with_weights-{wi}", layer_ids[li]))
            wi += 1
        nodes[model][0].append((f"layer-{li}", layer_ids[li]))
    hyper = OrderedDict()
    for h in ("iter", "beta_1", "beta_2", "decay", "learning_rate"):
        hid = new()
        hyper[h] = hid
        nodes[opt][0].append((h, hid))
        nodes[hid][1].append(("VARIABLE_VALUE", f"Adam/{h}", f"optimizer/{h}{ATTR}"))
    trainable = []
    wi = 0
    for li, (kname, role, _) in enumerate(plan):
        if role not in _PARTS:
            continue
        for part in _PARTS[role]:
            vid = new()
            nodes[layer_ids[li]][0].append((part, vid))
            nodes[vid][1].append(("VARIABLE_VALUE", f"{kname}/{part}", f"model/layer_with_weights-{wi}/{part}{ATTR}"))
            if not part.startswith("moving_"):
                trainable.append((vid, f"{kname}/{part}", f"model/layer_with_weights-{wi}/{part}"))
        wi += 1
    for vid, full, key in trainable:
        for slot in ("m", "v"):
            sid = new()
            nodes[sid][1].append(("VARIABLE_VALUE", f"Adam/{full}/{slot}", f"{key}/.OPTIMIZER_SLOT/optimizer/{slot}{ATTR}"))
            nodes[opt][2].append((vid, slot, sid))
    out = bytearray()
    for children, attrs, slots in nodes:
        body = bytearray()
        for name, nid in children:
            body += _pb_bytes(1, (_pb_varint(1, nid) if nid else b"") + _pb_bytes(2, name.encode()))
        for name, full, key in attrs:
            body += _pb_bytes(2, _pb_bytes(1, name.encode()) + _pb_bytes(2, full.encode()) + _pb_bytes(3, key.encode()))
        for orig, slot, sid in slots:
            body += _pb_bytes(3, (_pb_varint(1, orig) if orig else b"") + _pb_bytes(2, slot.encode()) + _pb_varint(3, sid))
        out += _pb_bytes(1, bytes(body))
    return bytes(out)


# ------------------------------------------------------------------------------------------------ U-Net <-> bundle
_PART_OF = {"kernel": "kernel", "bias": "bias", "gamma": "gamma", "beta": "beta", "moving_mean": "moving_mean", "moving_variance": "moving_var"}


def save_unet(prefix, block_names, kinds, params, adam_m=None, adam_v=None, step=0, learning_rate=3e-4, beta_1=0.9, beta_2=0.999):
    """params / adam_m / adam_v: {'<block>/<part>': array} in the TF layouts (UNet.export_params / export_flat)"""
    tensors = {OBJECT_GRAPH_KEY: object_graph(block_names, kinds)}
    for key, (b, part) in variable_keys(block_names, kinds).items():
        tensors[key + ATTR] = np.asarray(params[f"{b}/{_PART_OF[part]}"], dtype=np.float32)
        if adam_m is not None and not part.startswith("moving_"):
            tensors[f"{key}/.OPTIMIZER_SLOT/optimizer/m{ATTR}"] = np.asarray(adam_m[f"{b}/{_PART_OF[part]}"], dtype=np.float32)
            tensors[f"{key}/.OPTIMIZER_SLOT/optimizer/v{ATTR}"] = np.asarray(adam_v[f"{b}/{_PART_OF[part]}"], dtype=np.float32)
    tensors["optimizer/iter" + ATTR] = np.asarray(int(step), dtype=np.int64)
    tensors["optimizer/beta_1" + ATTR] = np.asarray(beta_1, dtype=np.float32)
    tensors["optimizer/beta_2" + ATTR] = np.asarray(beta_2, dtype=np.float32)
    tensors["optimizer/decay" + ATTR] = np.asarray(0.0, dtype=np.float32)
    tensors["optimizer/learning_rate" + ATTR] = np.asarray(learning_rate, dtype=np.float32)
    write_bundle(prefix, tensors)


def load_unet(prefix, block_names, kinds):
    """-> dict(params={...}, adam_m={...}|None, adam_v={...}|None, step=int|None, learning_rate=float|None).
    Optimizer state is optional (the reference restores with expect_partial)."""
    bundle = read_bundle(prefix)
    params, m, v = {}, {}, {}
    missing = []
    for key, (b, part) in variable_keys(block_names, kinds).items():
        name = f"{b}/{_PART_OF[part]}"
        if key + ATTR not in bundle:
            missing.append(key)
            continue
        params[name] = bundle[key + ATTR]
        mk = f"{key}/.OPTIMIZER_SLOT/optimizer/m{ATTR}"
        if mk in bundle:
            m[name] = bundle[mk]
            v[name] = bundle[f"{key}/.OPTIMIZER_SLOT/optimizer/v{ATTR}"]
    if missing:
        have = sorted(k for k in bundle if re.match(r"model/layer_with_weights-\d+/", k))
        raise IOError(f"checkpoint {prefix} lacks {len(missing)} model variables (first: {missing[0]}); it holds {len(have)} model tensors")
    n_train = sum(1 for _, (b, part) in variable_keys(block_names, kinds).items() if not part.startswith("moving_"))
    full = len(m) == n_train
    step = bundle.get("optimizer/iter" + ATTR)
    lr = bundle.get("optimizer/learning_rate" + ATTR)
    return dict(params=params, adam_m=m if full else None, adam_v=v if full else None,
                step=int(step) if step is not None else None, learning_rate=float(lr) if lr is not None else None)
