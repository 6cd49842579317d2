"""Minimal LMDB file-format reader / bulk writer (data.mdb), written from the published on-disk layout of LMDB 0.9
(mdb.c: MDB_page / MDB_node / MDB_meta / MDB_db).  py-lmdb and liblmdb are not installed in this environment, and the
reference's data surface is "an lmdb directory of isg_ai.proto ImageMaskPair records" (UNet/build_lmdb.py:29-60,
UNet/imagereader.py:120-153), so the record store is read and written directly.

Supported: 64-bit little-endian databases with the default flags of `lmdb.open(path)` as the reference uses them (one
unnamed database, unique byte-string keys in memcmp order, no DUPSORT, any page size recorded in the meta page).
The writer bulk-loads a fresh database (sorted keys -> leaf pages -> branch levels, values larger than the node limit in
overflow pages) and stamps both meta pages; it does not modify existing files.

CAVEAT: without liblmdb the two halves can only be validated against each other and against the
format description; interoperability with files produced by the real library is untested here.
"""
from __future__ import annotations

import mmap
import os
import struct

MDB_MAGIC = 0xBEEFC0DE
MDB_DATA_VERSION = 1
P_BRANCH, P_LEAF, P_OVERFLOW, P_META = 0x01, 0x02, 0x04, 0x08
F_BIGDATA = 0x01
PAGEHDRSZ = 16
NODESIZE = 8
DB_FMT = "<IHHQQQQQ"          # md_pad, md_flags, md_depth, md_branch_pages, md_leaf_pages, md_overflow_pages, md_entries, md_root
DB_SIZE = struct.calcsize(DB_FMT)   # 48
P_INVALID = 0xFFFFFFFFFFFFFFFF


def _data_path(path):
    return os.path.join(path, "data.mdb") if os.path.isdir(path) else path


class Reader:
    """Read-only view of an LMDB environment's main database.  `Reader(path)`; `len(r)`, `r.keys()`, `r.get(key)`,
    `r.items()`.  Values are returned as bytes copies (safe to keep after close())."""

    def __init__(self, path):
        fp = _data_path(path)
        if not os.path.exists(fp):
            raise IOError("Missing Database")
        self._f = open(fp, "rb")
        self._mm = mmap.mmap(self._f.fileno(), 0, access=mmap.ACCESS_READ)
        m0 = self._read_meta(0)
        if m0 is None:
            raise IOError("not an LMDB data file (bad magic)")
        metas = [m0]
        m1 = self._read_meta(m0["psize"])            # the second meta page sits one page further
        if m1 is not None:
            metas.append(m1)
        self.meta = max(metas, key=lambda m: m["txnid"])
        self.psize = self.meta["psize"]
        self.entries = self.meta["main"]["entries"]
        self.root = self.meta["main"]["root"]
        self.depth = self.meta["main"]["depth"]

    def _read_meta(self, off):
        mm = self._mm
        if off + PAGEHDRSZ + 24 + 2 * DB_SIZE + 16 > len(mm):
            return None
        flags = struct.unpack_from("<H", mm, off + 10)[0]
        magic, version, _addr, mapsize = struct.unpack_from("<IIQQ", mm, off + PAGEHDRSZ)
        if magic != MDB_MAGIC or not (flags & P_META):
            return None
        if version != MDB_DATA_VERSION:
            raise IOError(f"unsupported LMDB data version {version}")
        dbs = []
        for d in range(2):
            pad, fl, depth, br, lf, ov, ent, root = struct.unpack_from(DB_FMT, mm, off + PAGEHDRSZ + 24 + d * DB_SIZE)
            dbs.append(dict(pad=pad, flags=fl, depth=depth, branch_pages=br, leaf_pages=lf, overflow_pages=ov, entries=ent, root=root))
        last_pg, txnid = struct.unpack_from("<QQ", mm, off + PAGEHDRSZ + 24 + 2 * DB_SIZE)
        return dict(psize=dbs[0]["pad"], mapsize=mapsize, free=dbs[0], main=dbs[1], last_pg=last_pg, txnid=txnid)

    def close(self):
        self._mm.close()
        self._f.close()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def __len__(self):
        return int(self.entries)

    # ---- page helpers
    def _page(self, pgno):
        off = pgno * self.psize
        flags = struct.unpack_from("<H", self._mm, off + 10)[0]
        lower, upper = struct.unpack_from("<HH", self._mm, off + 12)
        return off, flags, (lower - PAGEHDRSZ) // 2

    def _node(self, page_off, idx):
        ptr = struct.unpack_from("<H", self._mm, page_off + PAGEHDRSZ + 2 * idx)[0]
        no = page_off + ptr
        lo, hi, fl, ksize = struct.unpack_from("<HHHH", self._mm, no)
        return no, lo, hi, fl, ksize

    def _leaf_value(self, no, lo, hi, fl, ksize):
        dsize = lo | (hi << 16)
        doff = no + NODESIZE + ksize
        if fl & F_BIGDATA:
            pg = struct.unpack_from("<Q", self._mm, doff)[0]
            start = pg * self.psize + PAGEHDRSZ
            return bytes(self._mm[start:start + dsize])
        return bytes(self._mm[doff:doff + dsize])

    def _walk(self, pgno, want_values):
        off, flags, n = self._page(pgno)
        if flags & P_LEAF:
            for i in range(n):
                no, lo, hi, fl, ksize = self._node(off, i)
                key = bytes(self._mm[no + NODESIZE:no + NODESIZE + ksize])
                yield (key, self._leaf_value(no, lo, hi, fl, ksize)) if want_values else key
        elif flags & P_BRANCH:
            for i in range(n):
                no, lo, hi, fl, ksize = self._node(off, i)
                yield from self._walk(lo | (hi << 16) | (fl << 32), want_values)
        else:
            raise IOError(f"unexpected page flags {flags:#x} at page {pgno}")

    def keys(self):
        if self.root == P_INVALID or self.entries == 0:
            return iter(())
        return self._walk(self.root, False)

    def items(self):
        if self.root == P_INVALID or self.entries == 0:
            return iter(())
        return self._walk(self.root, True)

    def get(self, key, default=None):
        """B+tree descent: in a branch page the first node's key is implicit (-inf), as in mdb_node_search."""
        if self.root == P_INVALID or self.entries == 0:
            return default
        pgno = self.root
        while True:
            off, flags, n = self._page(pgno)
            if flags & P_LEAF:
                lo_i, hi_i = 0, n - 1
                while lo_i <= hi_i:
                    mid = (lo_i + hi_i) // 2
                    no, lo, hi, fl, ksize = self._node(off, mid)
                    k = self._mm[no + NODESIZE:no + NODESIZE + ksize]
                    if k == key:
                        return self._leaf_value(no, lo, hi, fl, ksize)
                    if k < key:
                        lo_i = mid + 1
                    else:
                        hi_i = mid - 1
                return default
            # branch: last index i >= 1 with key_i <= key, else 0
            lo_i, hi_i, pick = 1, n - 1, 0
            while lo_i <= hi_i:
                mid = (lo_i + hi_i) // 2
                no, lo, hi, fl, ksize = self._node(off, mid)
                k = self._mm[no + NODESIZE:no + NODESIZE + ksize]
                if k <= key:
                    pick = mid
                    lo_i = mid + 1
                else:
                    hi_i = mid - 1
            no, lo, hi, fl, ksize = self._node(off, pick)
            pgno = lo | (hi << 16) | (fl << 32)


def _even(n):
    return (n + 1) & ~1


def write(path, items, psize=4096, map_size=None):
    """Bulk-load a fresh LMDB environment directory `path` (data.mdb + empty lock.mdb) from (key, value) byte pairs."""
    items = sorted(((bytes(k), bytes(v)) for k, v in items), key=lambda kv: kv[0])
    for a, b in zip(items, items[1:]):
        if a[0] == b[0]:
            raise ValueError(f"duplicate key {a[0]!r}")
    maxkey = 511
    for k, _ in items:
        if not (0 < len(k) <= maxkey):
            raise ValueError("LMDB keys must be 1..511 bytes")
    nodemax = (((psize - PAGEHDRSZ) // 2) & ~1) - 2
    pages = {}                 # pgno -> bytes
    next_pg = 2
    n_leaf = n_branch = n_over = 0

    def page_bytes(pgno, flags, nodes):
        """nodes: list of raw node byte strings (even-sized); laid out from the page end downward"""
        buf = bytearray(psize)
        upper = psize
        ptrs = []
        for nd in nodes:
            upper -= len(nd)
            buf[upper:upper + len(nd)] = nd
            ptrs.append(upper)
        lower = PAGEHDRSZ + 2 * len(nodes)
        assert lower <= upper, "page overflow"
        struct.pack_into("<QHHHH", buf, 0, pgno, 0, flags, lower, upper)
        for i, pt in enumerate(ptrs):
            struct.pack_into("<H", buf, PAGEHDRSZ + 2 * i, pt)
        return bytes(buf)

    # ---- leaf level
    leaf_nodes = []            # (key, node bytes)
    for k, v in items:
        if NODESIZE + len(k) + len(v) > nodemax:
            npg = (PAGEHDRSZ + len(v) + psize - 1) // psize
            pg = next_pg
            next_pg += npg
            n_over += npg
            buf = bytearray(npg * psize)
            struct.pack_into("<QHHI", buf, 0, pg, 0, P_OVERFLOW, npg)
            buf[PAGEHDRSZ:PAGEHDRSZ + len(v)] = v
            pages[pg] = bytes(buf)
            nd = struct.pack("<HHHH", len(v) & 0xFFFF, len(v) >> 16, F_BIGDATA, len(k)) + k + struct.pack("<Q", pg)
        else:
            nd = struct.pack("<HHHH", len(v) & 0xFFFF, len(v) >> 16, 0, len(k)) + k + v
        if len(nd) & 1:
            nd += b"\0"
        leaf_nodes.append((k, nd))

    def greedy_groups(sized):
        """sized: [(item, node_bytes_len)] -> groups of items that fit one page (2-byte slot per node)"""
        groups, cur, used = [], [], PAGEHDRSZ
        for item, sz in sized:
            need = sz + 2
            if cur and used + need > psize:
                groups.append(cur)
                cur, used = [], PAGEHDRSZ
            cur.append(item)
            used += need
        if cur:
            groups.append(cur)
        return groups

    depth = 0
    root = P_INVALID
    if leaf_nodes:
        level = []
        for grp in greedy_groups([((k, nd), len(nd)) for k, nd in leaf_nodes]):
            pg = next_pg
            next_pg += 1
            n_leaf += 1
            pages[pg] = page_bytes(pg, P_LEAF, [nd for _, nd in grp])
            level.append((grp[0][0], pg))
        depth = 1
        while len(level) > 1:
            nxt = []
            for grp in greedy_groups([((k, pg), _even(NODESIZE + len(k))) for k, pg in level]):
                pg = next_pg
                next_pg += 1
                n_branch += 1
                raw = []
                for i, (k, child) in enumerate(grp):
                    kk = b"" if i == 0 else k            # the first key of a branch page is implicit
                    nd = struct.pack("<HHHH", child & 0xFFFF, (child >> 16) & 0xFFFF, (child >> 32) & 0xFFFF, len(kk)) + kk
                    if len(nd) & 1:
                        nd += b"\0"
                    raw.append(nd)
                pages[pg] = page_bytes(pg, P_BRANCH, raw)
                nxt.append((grp[0][0], pg))
            level = nxt
            depth += 1
        root = level[0][1]

    last_pg = next_pg - 1
    total_bytes = next_pg * psize
    if map_size is None:
        map_size = max(total_bytes, 10 * 1024 * 1024)

    def meta_page(pgno, txnid):
        buf = bytearray(psize)
        struct.pack_into("<QHHHH", buf, 0, pgno, 0, P_META, 0, 0)
        struct.pack_into("<IIQQ", buf, PAGEHDRSZ, MDB_MAGIC, MDB_DATA_VERSION, 0, map_size)
        # FREE_DBI: md_pad carries the page size, md_flags = mm_flags = MDB_INTEGERKEY (0x08), as mdb_env_init_meta stamps it
        struct.pack_into(DB_FMT, buf, PAGEHDRSZ + 24, psize, 0x08, 0, 0, 0, 0, 0, P_INVALID)
        struct.pack_into(DB_FMT, buf, PAGEHDRSZ + 24 + DB_SIZE, 0, 0, depth, n_branch, n_leaf, n_over, len(items), root)   # MAIN_DBI
        struct.pack_into("<QQ", buf, PAGEHDRSZ + 24 + 2 * DB_SIZE, last_pg if last_pg >= 1 else 1, txnid)
        return bytes(buf)

    os.makedirs(path, exist_ok=True)
    with open(os.path.join(path, "data.mdb"), "wb") as f:
        f.write(meta_page(0, 0))
        f.write(meta_page(1, 1))
        pg = 2
        while pg < next_pg:
            blob = pages[pg]
            f.write(blob)
            pg += len(blob) // psize
    with open(os.path.join(path, "lock.mdb"), "wb") as f:
        f.write(b"")
