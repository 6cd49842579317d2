"""Tiled inference -- the counterpart of UNet/inference.py (same CLI, same functions, same outputs).

    python -m unetb200.inference --checkpoint_filepath CKPT --image_folder DIR --output_folder OUT \
           --number_classes K --number_channels C [--image_format tif]

Reference behaviour kept (file:line under /root/reference/UNet):
  * whole-image z-score before tiling (inference.py:206), per channel, std <= 1 -> subtract the mean only;
  * bottom/right reflect padding to a multiple of SIZE_FACTOR = 16 (inference.py:29-46, :143-156);
  * tile geometry: zone of responsibility = tile - 2 * radius, halo stripped only on unclamped sides, later tiles
    overwrite earlier ones where clamped zones overlap (inference.py:56-129);
  * images <= 1024 x 1024 take the single-shot path (inference.py:209-212);
  * mask dtype narrowing by three independent ifs (inference.py:215-220).
What is new underneath: the image is uploaded ONCE, normalised on the GPU (ub_zscore), equal-shaped tiles are batched
through the folded inference forward, the argmax of every tile's zone is written by the head kernel straight into a
uint8 device mask (no softmax round trip, no per-tile H2D/D2H), and tiles are sharded round-robin over the ranks of a
torchrun job (the zones are made disjoint first -- see tile_plan -- so the shards combine with one SUM reduce); segment_sharded
additionally uploads and normalises only the rows a rank's tiles read.
"""
from __future__ import annotations

import argparse
import os

import numpy as np

TILE_SIZE = 1024          # UNet/inference.py:24
SIZE_FACTOR = 16          # UNet/model.py:25
TILE_BATCH = 4            # equal-shaped tiles per forward


def tile_plan(height, width, tile_size, radius):
    """Tiles of UNet/inference.py:56-95 with DISJOINT zones.

    The reference clamps a zone to the image edge whenever the tile's halo crosses it and lets the next loop iteration
    overwrite the tail.  The value that survives at a pixel is the one written by the LAST tile (row-major
    order) whose zone contains it, which is the tile whose start i satisfies i <= y < i + zone: clipping every zone to
    [i, min(i + zone, height)) yields exactly the reference's final mask with no write-order dependence, so the
    tiles can run in any order, batched, and on different GPUs.
    Returns a list of dicts: tile box y0,y1,x0,x1; crop box cy0,cy1,cx0,cx1 inside the tile; destination dy,dx."""
    zone = tile_size - 2 * radius
    assert tile_size % SIZE_FACTOR == 0
    assert radius % SIZE_FACTOR == 0
    assert zone >= radius
    plan = []
    for i in range(0, height, zone):
        y0 = max(i - radius, 0)
        y1 = min(i + zone + radius, height)
        zy = min(i + zone, height) - i
        for j in range(0, width, zone):
            x0 = max(j - radius, 0)
            x1 = min(j + zone + radius, width)
            zx = min(j + zone, width) - j
            plan.append(dict(y0=y0, y1=y1, x0=x0, x1=x1, cy0=i - y0, cy1=i - y0 + zy, cx0=j - x0, cx1=j - x0 + zx, dy=i, dx=j))
    return plan


def narrow_mask_dtype(mask):
    """UNet/inference.py:215-220: three independent ifs; a maximum of exactly 65536 matches none and stays as it is."""
    mx = int(np.max(mask)) if mask.size else 0
    if 0 <= mx <= 255:
        mask = mask.astype(np.uint8)
    if 255 < mx < 65536:
        mask = mask.astype(np.uint16)
    if mx > 65536:
        mask = mask.astype(np.int32)
    return mask


def _as_hwc(img):
    if img.ndim not in (2, 3):
        raise IOError('Invalid number of dimensions for input image. Expecting HW or HWC dimension ordering.')
    if img.ndim == 2:
        img = img.reshape((img.shape[0], img.shape[1], 1))
    return img


def _pad_amounts(h, w):
    return (SIZE_FACTOR - h % SIZE_FACTOR) % SIZE_FACTOR, (SIZE_FACTOR - w % SIZE_FACTOR) % SIZE_FACTOR


def _to_device_chw(img_hwc, device):
    """host HWC (any dtype) -> device CHW float32 (unpadded: the tile reader mirrors past the image edge)"""
    import torch
    chw = np.ascontiguousarray(img_hwc.transpose((2, 0, 1)))
    t = torch.from_numpy(chw)
    if device.type == "cuda":
        t = t.pin_memory().to(device, non_blocking=True)
    return t.to(torch.float32) if t.dtype != torch.float32 else t


def _run_tiles(img_chw, img_h, tiles, unet_model, mask, mask_ld, tile_batch, row0=0, zrow0=0):
    """Forward `tiles` (tile_plan entries in the coordinates of the PADDED image) and write each zone's argmax into `mask`.
    img_chw [C, rows, W]: rows [row0, row0 + rows) of the normalised UNPADDED image, img_h = its total height; the tiles are read
    in place (UNet.predict_tiles_from), grouped by shape so that equal-sized tiles share a forward.  zrow0: image row of mask row 0."""
    import torch
    if not tiles:
        return
    dev = img_chw.device
    groups = {}
    for t in tiles:
        groups.setdefault((t["y1"] - t["y0"], t["x1"] - t["x0"]), []).append(t)
    order = [t for shape in groups for t in groups[shape]]
    geo = torch.from_numpy(np.array([[t["cy0"], t["cy1"], t["cx0"], t["cx1"], t["dy"] - zrow0, t["dx"]] for t in order], dtype=np.int32)).to(dev)
    org = torch.from_numpy(np.array([[t["y0"] - row0, t["x0"]] for t in order], dtype=np.int32)).to(dev)
    k = 0
    for (h, w), ts in groups.items():
        for s in range(0, len(ts), tile_batch):
            n = min(tile_batch, len(ts) - s)
            unet_model.predict_tiles_from(img_chw, img_h - row0, img_chw.shape[2], org[k:k + n], n, h, w, geo[k:k + n], mask, mask_ld)
            k += n


def segment_device(img_chw, unet_model, tile_size=TILE_SIZE, radius=None, tile_batch=TILE_BATCH, dist=None):
    """img_chw: NORMALISED float32 device tensor [C, H, W] (unpadded).  Returns the uint8 device mask [Hp, Wp] of the image
    reflect-padded bottom/right to multiples of 16 (UNet/inference.py:29-46; callers crop [:H, :W]); the padding is never
    materialised.  tile_size None = one forward of the whole image.  With `dist` (unetb200.dist.DataParallel, world > 1) the tiles
    are sharded round-robin and the mask is SUM-reduced (zones are disjoint, unowned pixels are zero)."""
    import torch
    C, H, W = img_chw.shape
    pad_y, pad_x = _pad_amounts(H, W)
    Hp, Wp = H + pad_y, W + pad_x
    dev = img_chw.device
    mask = torch.zeros((Hp, Wp), dtype=torch.uint8, device=dev)
    if tile_size is None:                       # single-shot path (UNet/inference.py:139-173)
        plan = [dict(y0=0, y1=Hp, x0=0, x1=Wp, cy0=0, cy1=Hp, cx0=0, cx1=Wp, dy=0, dx=0)]
    else:
        if radius is None:
            radius = unet_model.estimate_radius()
        plan = tile_plan(Hp, Wp, tile_size, radius)
    rank, world = (dist.rank, dist.world_size) if dist is not None else (0, 1)
    _run_tiles(img_chw, H, plan[rank::world], unet_model, mask, Wp, tile_batch)
    if world > 1:
        import torch.distributed as td
        td.all_reduce(mask, op=td.ReduceOp.SUM)
    return mask


def shard_plan(height, width, tile_size, radius, world):
    """Sharding of the tile plan over `world` ranks: rank r owns a contiguous RUN of tiles in row-major order (equal counts to within
    one tile: 625 tiles over 8 ranks = 78 or 79 each -- whole tile rows would give 3 or 4 rows of 25, a 28 % imbalance).  Per rank:
      tiles          its tiles (tile_plan entries, coordinates of the padded image)
      y0, y1         image rows its tiles read (zones plus halo): the only rows it uploads and normalises
      sy0, sy1       full-width rows whose pixels it contributes to the whole-image z-score statistics: the zone rows of the tile rows
                     whose FIRST tile lies in its run (disjoint over ranks, covering the image, inside [y0, y1))"""
    plan = tile_plan(height, width, tile_size, radius)
    out = []
    for r in range(world):
        lo, hi = (len(plan) * r) // world, (len(plan) * (r + 1)) // world
        tiles = plan[lo:hi]
        if not tiles:
            out.append(dict(tiles=[], y0=0, y1=0, sy0=0, sy1=0))
            continue
        starts = [t for t in tiles if t["dx"] == 0]
        sy0 = min((t["dy"] for t in starts), default=0)
        sy1 = max((t["dy"] + t["cy1"] - t["cy0"] for t in starts), default=0)
        out.append(dict(tiles=tiles, y0=min(t["y0"] for t in tiles), y1=max(t["y1"] for t in tiles), sy0=sy0, sy1=sy1))
    return out


def segment_sharded(raw_host, unet_model, dist, tile_size=TILE_SIZE, radius=None, tile_batch=TILE_BATCH, out_host=None):
    """Tiled inference of one image with the TILES sharded over the ranks of `dist` (shard_plan).  raw_host: pinned host tensor
    [C, H, W] of raw pixels (uint8 / uint16-as-int16 bits / float32), the same on every rank.  Each rank uploads only the rows its
    tiles read; the per-channel z-score statistics of the UNPADDED image (UNet/inference.py:206) are summed over disjoint row ranges
    and all-reduced (2 doubles per channel); the rows are normalised and the tiles segmented in place (the reflect padding of the
    bottom / right edge, inference.py:46, is mirror indexing inside the tile reader); zones are disjoint, so one SUM all-reduce of
    the uint8 mask over NVLink assembles the result on every rank.
    Returns the uint8 device mask [H, W]; if `out_host` (pinned uint8 [H, W]) is given, rank 0 copies it there."""
    import torch
    import torch.distributed as td
    from . import _C
    C, H, W = raw_host.shape
    dev = unet_model.device
    rank, world = dist.rank, dist.world_size
    pad_y, pad_x = _pad_amounts(H, W)
    Hp, Wp = H + pad_y, W + pad_x
    if radius is None:
        radius = unet_model.estimate_radius()
    b = shard_plan(Hp, Wp, tile_size, radius, world)[rank]
    code = {torch.uint8: 0, torch.int16: 1, torch.float32: 2}[raw_host.dtype]
    st = unet_model._stream()
    sums = torch.zeros((C, 2), dtype=torch.float64, device=dev)
    mask = torch.zeros((Hp, Wp), dtype=torch.uint8, device=dev)
    if b["tiles"]:
        y0, y1 = b["y0"], min(b["y1"], H)                       # rows that exist in the unpadded image
        raw = torch.empty((C, y1 - y0, W), dtype=raw_host.dtype, device=dev)
        for c in range(C):                                       # each channel's rows are one contiguous chunk of the host image
            raw[c].copy_(raw_host[c, y0:y1], non_blocking=True)
        s0, s1 = b["sy0"], min(b["sy1"], H)                      # rows of the unpadded image that carry this rank's share of the statistics
        scratch = torch.empty(C * _C.UB_ZSCORE_BLOCKS * 2, dtype=torch.float64, device=dev)
        if s1 > s0:
            _C.call("ub_zscore_sums", raw[:, s0 - y0:], code, sums, scratch, C, (s1 - s0) * W, (y1 - y0) * W, st)
    if world > 1:
        td.all_reduce(sums, op=td.ReduceOp.SUM)
    if b["tiles"]:
        x = torch.empty((C, y1 - y0, W), dtype=torch.float32, device=dev)
        _C.call("ub_zscore_apply_sums", raw, code, x, sums, float(H) * float(W), C, (y1 - y0) * W, st)
        _run_tiles(x, H, b["tiles"], unet_model, mask, Wp, tile_batch, row0=y0)
    if world > 1:
        td.all_reduce(mask, op=td.ReduceOp.SUM)                   # disjoint zones, zeros elsewhere
    mask = mask[:H, :W]
    if out_host is not None and rank == 0:
        out_host.copy_(mask, non_blocking=True)
    return mask


def zscore_device(img_chw_raw, unet_model):
    """whole-image per-channel z-score (UNet/imagereader.py:33-66, called at inference.py:206) on the GPU.
    img_chw_raw: device tensor [C,H,W] of dtype uint8 / uint16-as-int16-bits / float32."""
    import torch
    from . import _C
    C, H, W = img_chw_raw.shape
    code = {torch.uint8: 0, torch.int16: 1, torch.uint16: 1, torch.float32: 2}[img_chw_raw.dtype]
    out = torch.empty((C, H, W), dtype=torch.float32, device=img_chw_raw.device)
    scratch = torch.empty(C * _C.UB_ZSCORE_BLOCKS * 2, dtype=torch.float64, device=img_chw_raw.device)
    _C.call("ub_zscore", img_chw_raw, code, out, scratch, C, H * W, torch.cuda.current_stream(out.device).cuda_stream)
    return out


def _segment_normalised_host(img, unet_model, tile_size):
    import torch
    img = _as_hwc(np.asarray(img))
    h0, w0 = img.shape[0], img.shape[1]
    pad_y, pad_x = _pad_amounts(h0, w0)
    if pad_y:
        print('image height needs to be a multiple of {}, padding with reflect'.format(SIZE_FACTOR))
    if pad_x:
        print('image width needs to be a multiple of {}, padding with reflect'.format(SIZE_FACTOR))
    dev = unet_model.device
    x = _to_device_chw(img.astype(np.float32, copy=False), dev)
    mask = segment_device(x, unet_model, tile_size, dist=getattr(unet_model, "dist", None))
    out = mask.cpu().numpy().astype(np.int32)
    return out[:h0, :w0]


def _inference_tiling(img, unet_model, tile_size):
    """UNet/inference.py:27-136: img = normalised HW / HWC float array -> int32 mask [H, W]."""
    radius = unet_model.estimate_radius()
    print('Estimated radius based on ERF : "{}"'.format(radius))
    return _segment_normalised_host(img, unet_model, tile_size)


def _inference(img, unet_model):
    """UNet/inference.py:139-173: single forward of the whole (padded) image."""
    return _segment_normalised_host(img, unet_model, None)


def imread(fp):
    """UNet/imagereader.py:69-74 reads with skimage.io (not installed here); PIL handles 8/16-bit TIFF, PNG, JPEG."""
    from PIL import Image
    Image.MAX_IMAGE_PIXELS = None
    return np.asarray(Image.open(fp))


def imsave(fp, mask, image_format):
    """UNet/inference.py:221-227: 'tif' -> BigTIFF, zlib level 6, 1024 x 1024 tiles (what skimage/tifffile is asked for
    there, written by unetb200/tiffio.py); anything else -> the format's plain writer"""
    if 'tif' in image_format:
        from . import tiffio
        tiffio.write_tiled_bigtiff(fp, mask, tile=(1024, 1024), level=6)
        return
    from PIL import Image
    Image.fromarray(mask).save(fp)


def segment_file(img_filepath, unet_model):
    """read -> upload raw pixels -> GPU z-score -> tiled / single-shot inference -> narrowed host mask"""
    import torch
    img = _as_hwc(imread(img_filepath))
    h0, w0 = img.shape[0], img.shape[1]
    print('  img.shape={}'.format(img.shape))
    dev = unet_model.device
    raw = img
    if raw.dtype == np.uint16:
        raw = raw.view(np.int16)              # torch uploads the bits; ub_zscore reads them as uint16
    elif raw.dtype not in (np.uint8, np.float32):
        raw = raw.astype(np.float32)
    chw = np.array(raw.transpose((2, 0, 1)), order="C")      # writable copy (PIL hands out read-only buffers)
    t = torch.from_numpy(chw).pin_memory().to(dev, non_blocking=True)
    x = zscore_device(t, unet_model)          # statistics of the UNPADDED image, as the reference (inference.py:206)
    if h0 > TILE_SIZE or w0 > TILE_SIZE:       # inference.py:209
        mask = segment_device(x, unet_model, TILE_SIZE, dist=getattr(unet_model, "dist", None))
    else:
        mask = segment_device(x, unet_model, None, dist=getattr(unet_model, "dist", None))
    return narrow_mask_dtype(mask.cpu().numpy()[:h0, :w0])


def inference(checkpoint_filepath, image_folder, output_folder, number_classes, number_channels, image_format):
    """UNet/inference.py:176-227"""
    from . import model
    print('Arguments:')
    print('checkpoint_filepath = {}'.format(checkpoint_filepath))
    print('image_folder = {}'.format(image_folder))
    print('output_folder = {}'.format(output_folder))
    print('image_format = {}'.format(image_format))
    print('number_classes = {}'.format(number_classes))
    print('number_channels = {}'.format(number_channels))

    if not os.path.exists(output_folder):
        os.mkdir(output_folder)
    img_filepath_list = [os.path.join(image_folder, fn) for fn in os.listdir(image_folder) if fn.endswith('.{}'.format(image_format))]

    dp = None
    if int(os.environ.get("WORLD_SIZE", "1")) > 1:
        from .dist import DataParallel
        dp = DataParallel()
    unet_model = model.UNet(number_classes, 1, number_channels, 1e-4, dist=dp)
    unet_model.load_checkpoint(checkpoint_filepath)

    print('Starting inference of file list')
    for i, img_filepath in enumerate(img_filepath_list):
        _, slide_name = os.path.split(img_filepath)
        print('{}/{} : {}'.format(i, len(img_filepath_list), slide_name))
        print('Loading image: {}'.format(img_filepath))
        segmented_mask = segment_file(img_filepath, unet_model)
        if dp is None or dp.rank == 0:
            imsave(os.path.join(output_folder, slide_name), segmented_mask, image_format)
    if dp is not None:
        dp.shutdown()


def main(argv=None):
    parser = argparse.ArgumentParser(prog='inference', description='Script to detect stars with the selected unet model')
    parser.add_argument('--checkpoint_filepath', dest='checkpoint_filepath', type=str, help='Checkpoint filepath to the  model to use', required=True)
    parser.add_argument('--image_folder', dest='image_folder', type=str, help='filepath to the folder containing tif images to inference (Required)', required=True)
    parser.add_argument('--output_folder', dest='output_folder', type=str, required=True)
    parser.add_argument('--number_classes', dest='number_classes', type=int, required=True)
    parser.add_argument('--number_channels', dest='number_channels', type=int, required=True)
    parser.add_argument('--image_format', dest='image_format', type=str, help='format (extension) of the input images. E.g {tif, jpg, png)', default='tif')
    args = parser.parse_args(argv)
    inference(args.checkpoint_filepath, args.image_folder, args.output_folder, args.number_classes, args.number_channels, args.image_format)


if __name__ == "__main__":
    main()
