"""Mask output in the container the reference writes (UNet/inference.py:221-222):
`skimage.io.imsave(path, mask, compress=6, bigtiff=True, tile=(1024, 1024))` = a little-endian BigTIFF with one IFD,
1024 x 1024 tiles, Adobe-deflate (zlib level 6) compression, edge tiles padded to the full tile size.
skimage / tifffile are not installed here and Pillow cannot write tiles, so the few structures involved are written
directly (TIFF 6.0 + the BigTIFF 8-byte-offset variant).  Readable by libtiff (Pillow, OpenCV): tests/test_tiff_cpu.py."""
from __future__ import annotations

import struct
import zlib

import numpy as np

_TYPES = {"B": 1, "ascii": 2, "H": 3, "I": 4, "rational": 5, "Q": 16}
_SAMPLE_FORMAT = {"u": 1, "i": 2, "f": 3}


def _entry(tag, typ, values):
    """one 20-byte BigTIFF IFD entry -> (entry bytes with a zero value field, payload or None, inline value bytes or None)"""
    if typ == "ascii":
        data = values.encode() + b"\x00"
        count = len(data)
    elif typ == "rational":
        data = b"".join(struct.pack("<II", a, b) for a, b in values)
        count = len(values)
    else:
        data = struct.pack("<%d%s" % (len(values), typ), *values)
        count = len(values)
    head = struct.pack("<HHQ", tag, _TYPES[typ], count)
    if len(data) <= 8:
        return head, None, data + b"\x00" * (8 - len(data))
    return head, data, None


def write_tiled_bigtiff(path, image, tile=(1024, 1024), level=6, software="unetb200"):
    """image: [H, W] or [H, W, S] array of an 8/16/32/64-bit integer or float dtype"""
    a = np.asarray(image)
    if a.ndim == 2:
        a = a[:, :, None]
    if a.ndim != 3:
        raise IOError(f"expected [H,W] or [H,W,S], got shape {a.shape}")
    if a.dtype.kind not in _SAMPLE_FORMAT or a.dtype.itemsize not in (1, 2, 4, 8):
        raise IOError(f"unsupported sample dtype {a.dtype}")
    a = a.astype(a.dtype.newbyteorder("<"), copy=False)
    H, W, S = a.shape
    th, tw = tile
    if th % 16 or tw % 16:
        raise IOError("TIFF tile sizes must be multiples of 16")
    ty, tx = (H + th - 1) // th, (W + tw - 1) // tw
    with open(path, "wb") as f:
        f.write(struct.pack("<2sHHHQ", b"II", 43, 8, 0, 0))             # BigTIFF header; first-IFD offset patched below
        offsets, counts = [], []
        buf = np.zeros((th, tw, S), dtype=a.dtype)
        for j in range(ty):
            for i in range(tx):
                blk = a[j * th:(j + 1) * th, i * tw:(i + 1) * tw]
                if blk.shape[:2] != (th, tw):                            # edge tiles are padded (TIFF 6.0 section 15)
                    buf[:] = 0
                    buf[:blk.shape[0], :blk.shape[1]] = blk
                    blk = buf
                z = zlib.compress(np.ascontiguousarray(blk).tobytes(), level)
                offsets.append(f.tell())
                counts.append(len(z))
                f.write(z)
                if f.tell() & 1:
                    f.write(b"\x00")
        entries = [
            _entry(256, "I", [W]), _entry(257, "I", [H]), _entry(258, "H", [8 * a.dtype.itemsize] * S),
            _entry(259, "H", [8]),                                       # Adobe deflate
            _entry(262, "H", [1 if S != 3 else 2]),                      # min-is-black / RGB
            _entry(277, "H", [S]), _entry(282, "rational", [(1, 1)]), _entry(283, "rational", [(1, 1)]),
            _entry(284, "H", [1]), _entry(296, "H", [1]), _entry(305, "ascii", software),
            _entry(322, "I", [tw]), _entry(323, "I", [th]), _entry(324, "Q", offsets), _entry(325, "Q", counts),
            _entry(339, "H", [_SAMPLE_FORMAT[a.dtype.kind]] * S),
        ]
        if S > 1 and S != 3:
            entries.append(_entry(338, "H", [0] * (S - 1)))              # extra samples: unspecified
        entries.sort(key=lambda e: struct.unpack("<H", e[0][:2])[0])
        # out-of-line payloads first, then the IFD
        placed = []
        for head, payload, inline in entries:
            if payload is None:
                placed.append(head + inline)
            else:
                if f.tell() & 7:
                    f.write(b"\x00" * (8 - (f.tell() & 7)))
                placed.append(head + struct.pack("<Q", f.tell()))
                f.write(payload)
        if f.tell() & 7:
            f.write(b"\x00" * (8 - (f.tell() & 7)))
        ifd = f.tell()
        f.write(struct.pack("<Q", len(placed)) + b"".join(placed) + struct.pack("<Q", 0))
        f.seek(8)
        f.write(struct.pack("<Q", ifd))


def read_tiff_structure(path):
    """minimal BigTIFF / classic TIFF IFD reader (first directory): {tag: values}; used by the tests"""
    b = open(path, "rb").read()
    if b[:2] != b"II":
        raise IOError("only little-endian TIFF files are handled here")
    big = struct.unpack_from("<H", b, 2)[0] == 43
    ifd = struct.unpack_from("<Q", b, 8)[0] if big else struct.unpack_from("<I", b, 4)[0]
    n = struct.unpack_from("<Q" if big else "<H", b, ifd)[0]
    pos = ifd + (8 if big else 2)
    fmt = {1: "B", 2: "c", 3: "H", 4: "I", 16: "Q"}
    out = {"bigtiff": big}
    for _ in range(n):
        if big:
            tag, typ, cnt = struct.unpack_from("<HHQ", b, pos)
            vpos, room, pos = pos + 12, 8, pos + 20
        else:
            tag, typ, cnt = struct.unpack_from("<HHI", b, pos)
            vpos, room, pos = pos + 8, 4, pos + 12
        if typ == 5:
            size = 8 * cnt
        else:
            size = struct.calcsize(fmt[typ]) * cnt
        if size > room:
            vpos = struct.unpack_from("<Q" if big else "<I", b, vpos)[0]
        if typ == 5:
            out[tag] = [struct.unpack_from("<II", b, vpos + 8 * i) for i in range(cnt)]
        elif typ == 2:
            out[tag] = b[vpos:vpos + cnt].rstrip(b"\x00").decode()
        else:
            out[tag] = list(struct.unpack_from("<%d%s" % (cnt, fmt[typ]), b, vpos))
    return out
