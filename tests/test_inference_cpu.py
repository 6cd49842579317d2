"""CPU (-m "not gpu"): host logic of unetb200.inference against the oracle's restatement of UNet/inference.py."""
import numpy as np
import pytest

import unetb200.inference as I
from oracle import unet_oracle as O


@pytest.mark.parametrize("H,W,tile,radius", [(400, 336, 288, 96), (1248, 832, 288, 96), (2000 // 16 * 16, 1008, 512, 96),
                                             (96, 96, 288, 96), (832 * 2 + 32, 832 + 864, 1024, 96), (1200, 1136, 1024, 96),
                                             (640, 640, 256, 64), (656, 1296, 256, 64)])
def test_disjoint_zones_reproduce_last_writer_wins(H, W, tile, radius):
    """the reference composes the mask by overwriting; tile_plan clips the zones instead: same result"""
    ref = np.full((H, W), -1, dtype=np.int64)
    src_y = np.zeros((H, W), dtype=np.int64)          # which TILE-LOCAL row the surviving value came from
    for k, t in enumerate(O.tile_plan(H, W, tile, radius)):
        hh, ww = t["y_end"] - t["y_st"], t["x_end"] - t["x_st"]
        ly, lx = np.meshgrid(np.arange(hh), np.arange(ww), indexing="ij")
        sl = (slice(t["pre_y"], hh - t["post_y"] if t["post_y"] else None), slice(t["pre_x"], ww - t["post_x"] if t["post_x"] else None))
        ref[t["y_st_z"]:t["y_end_z"], t["x_st_z"]:t["x_end_z"]] = k
        src_y[t["y_st_z"]:t["y_end_z"], t["x_st_z"]:t["x_end_z"]] = (ly * 100000 + lx)[sl]
    got = np.full((H, W), -1, dtype=np.int64)
    got_src = np.zeros((H, W), dtype=np.int64)
    cover = np.zeros((H, W), dtype=np.int64)
    plan = I.tile_plan(H, W, tile, radius)
    assert len(plan) == len(O.tile_plan(H, W, tile, radius))
    for k, t in enumerate(plan):
        hh, ww = t["y1"] - t["y0"], t["x1"] - t["x0"]
        assert hh % 16 == 0 and ww % 16 == 0
        ly, lx = np.meshgrid(np.arange(hh), np.arange(ww), indexing="ij")
        zy, zx = t["cy1"] - t["cy0"], t["cx1"] - t["cx0"]
        got[t["dy"]:t["dy"] + zy, t["dx"]:t["dx"] + zx] = k
        got_src[t["dy"]:t["dy"] + zy, t["dx"]:t["dx"] + zx] = (ly * 100000 + lx)[t["cy0"]:t["cy1"], t["cx0"]:t["cx1"]]
        cover[t["dy"]:t["dy"] + zy, t["dx"]:t["dx"] + zx] += 1
    assert (cover == 1).all()                          # disjoint and complete
    assert (got == ref).all() and (got_src == src_y).all()
    # tile boxes agree with the reference's
    for a, b in zip(plan, O.tile_plan(H, W, tile, radius)):
        assert (a["y0"], a["y1"], a["x0"], a["x1"]) == (b["y_st"], b["y_end"], b["x_st"], b["x_end"])


def test_narrow_mask_dtype_quirks():
    for mx, dt in [(0, np.uint8), (255, np.uint8), (256, np.uint16), (65535, np.uint16), (65536, np.int32), (70000, np.int32)]:
        m = np.zeros((4, 4), dtype=np.int32)
        m[0, 0] = mx
        assert I.narrow_mask_dtype(m).dtype == dt == O.narrow_mask_dtype(m).dtype


def test_bad_dimensions_raise_ioerror():
    with pytest.raises(IOError):
        I._as_hwc(np.zeros((2, 2, 2, 2)))


def test_shard_plan_partitions_the_tiles():
    """tile-run sharding (unetb200.inference.shard_plan): every tile belongs to exactly one rank, counts are balanced to one tile, the
    rows a rank uploads contain its tiles, and the statistics row ranges are disjoint, cover the image and lie inside the uploaded rows"""
    from unetb200 import inference as I
    for (H, W, world) in [(20000, 20000, 8), (2208, 1520, 2), (2208, 1520, 3), (1040, 5000, 4), (4096, 4096, 1), (1040, 1040, 8)]:
        plan = I.tile_plan(H, W, 1024, 96)
        shards = I.shard_plan(H, W, 1024, 96, world)
        assert len(shards) == world
        seen = [(t["dy"], t["dx"]) for b in shards for t in b["tiles"]]
        assert seen == [(t["dy"], t["dx"]) for t in plan]                      # contiguous runs in row-major order
        counts = [len(b["tiles"]) for b in shards]
        assert max(counts) - min(counts) <= 1
        edge = 0
        for b in shards:
            if not b["tiles"]:
                assert b["sy1"] == b["sy0"]
                continue
            assert all(b["y0"] <= t["y0"] and t["y1"] <= b["y1"] for t in b["tiles"])
            if b["sy1"] > b["sy0"]:
                assert b["sy0"] == edge and b["y0"] <= b["sy0"] and b["sy1"] <= b["y1"]
                edge = b["sy1"]
        assert edge == H
    cover = np.zeros((2208, 1520), dtype=np.int8)
    for b in I.shard_plan(2208, 1520, 1024, 96, 3):
        for t in b["tiles"]:
            cover[t["dy"]:t["dy"] + t["cy1"] - t["cy0"], t["dx"]:t["dx"] + t["cx1"] - t["cx0"]] += 1
    assert (cover == 1).all()
