"""Training-data reader -- the counterpart of UNet/imagereader.py behind the same constructor and record format.

Records: an LMDB directory whose values are `isg_ai.ImageMaskPair` protobuf messages (UNet/isg_ai.proto:16-31) and
whose keys are "{name}[_i{y}_j{x}]:{c0,c1,...}" (UNet/build_lmdb.py:44-59, :123, :178).  The protobuf wire format of
that one message is decoded / encoded by hand here (8 fields, varints and length-delimited bytes), and the LMDB file is
read by unetb200.lmdbfile -- neither py-lmdb nor a compatible generated isg_ai_pb2 is importable in this environment.

Reference semantics kept (file:line under /root/reference/UNet/imagereader.py):
  * ctor `(img_db, use_augmentation, balance_classes, shuffle, num_workers, number_classes)` (:87); `IOError("Missing
    Database")` (:112-115); tile height/width must be multiples of 16 (:136-139);
  * sampling (:209-243): shuffle -> a uniformly random key per example (with replacement); balance_classes -> a uniformly
    random CLASS first, then a random key whose suffix lists that class; no shuffle -> sequential keys;
  * per-tile z-score of the float32 CHW image (:300, :33-66) and one-hot labels with IndexError on a label >= K (:302-312).
What is different underneath ("device-side fast reader contract"): a batch is shipped to the GPU as RAW
pixels (uint8 / uint16 / float32) plus uint8 class indices from pinned memory -- 2-3 bytes per pixel instead of the
reference's 4 + 4K -- and the z-score (ub_zscore, one plane per image and channel) runs on the device; the train step
consumes the class index directly.  `get_example()` still yields the reference's (float32 CHW, int32 one-hot HWK)
pair for callers that want it.  Augmentation (UNet/augment.py, applied per example at imagereader.py:283-294 with the
class constants :78-85) runs on the DEVICE on the raw batch (unetb200/augment.py + csrc/augment.cu): a reader built with
use_augmentation=True draws the per-example parameters in the reference's order (`draw_augmentation`) and
`device_batch()` applies them between the upload and the z-score.
"""
from __future__ import annotations

import os
import random

import numpy as np

from . import lmdbfile

SIZE_FACTOR = 16


# ---------------------------------------------------------------------------------------------- protobuf (ImageMaskPair)
def _varint(buf, pos):
    v = 0
    shift = 0
    while True:
        b = buf[pos]
        pos += 1
        v |= (b & 0x7F) << shift
        if not b & 0x80:
            return v, pos
        shift += 7


def _enc_varint(v):
    out = bytearray()
    v &= (1 << 64) - 1
    while True:
        b = v & 0x7F
        v >>= 7
        if v:
            out.append(b | 0x80)
        else:
            out.append(b)
            return bytes(out)


_FIELDS = {1: "channels", 2: "img_height", 3: "img_width", 4: "img_type", 5: "mask_type", 6: "image", 7: "mask", 8: "labels"}


def decode_pair(value):
    """serialized ImageMaskPair -> dict (ints, str dtype codes, memoryviews of the pixel bytes)"""
    mv = memoryview(value)
    out = {}
    pos = 0
    n = len(mv)
    while pos < n:
        tag, pos = _varint(mv, pos)
        field, wt = tag >> 3, tag & 7
        if wt == 0:
            v, pos = _varint(mv, pos)
            if v >= 1 << 63:
                v -= 1 << 64
        elif wt == 2:
            ln, pos = _varint(mv, pos)
            v = mv[pos:pos + ln]
            pos += ln
        elif wt == 1:
            v = bytes(mv[pos:pos + 8])
            pos += 8
        elif wt == 5:
            v = bytes(mv[pos:pos + 4])
            pos += 4
        else:
            raise IOError(f"unsupported protobuf wire type {wt}")
        name = _FIELDS.get(field)
        if name is None:
            continue
        if name in ("img_type", "mask_type"):
            v = bytes(v).decode("ascii")
        out[name] = v
    return out


def encode_pair(image_hwc, mask_hw):
    """numpy image [H,W,C] (or [H,W]) + mask [H,W] -> serialized ImageMaskPair (as UNet/build_lmdb.py:29-60 fills it)"""
    img = np.ascontiguousarray(image_hwc)
    if img.ndim == 2:
        img = img.reshape(img.shape[0], img.shape[1], 1)
    mask = np.ascontiguousarray(mask_hw)
    labels = np.unique(mask).astype(np.uint8)

    def f_int(field, v):
        return _enc_varint(field << 3) + _enc_varint(int(v))

    def f_bytes(field, b):
        return _enc_varint((field << 3) | 2) + _enc_varint(len(b)) + b

    return b"".join([f_int(1, img.shape[2]), f_int(2, img.shape[0]), f_int(3, img.shape[1]),
                     f_bytes(4, img.dtype.str.encode("ascii")), f_bytes(5, mask.dtype.str.encode("ascii")),
                     f_bytes(6, img.tobytes()), f_bytes(7, mask.tobytes()), f_bytes(8, labels.tobytes())])


def record_key(name, mask_hw, i=None, j=None):
    """'{name}[_i{y}_j{x}]:{classes present}' (UNet/build_lmdb.py:44-59, :123)"""
    present = ",".join(str(int(c)) for c in np.unique(mask_hw))
    base = name if i is None else "{}_i{}_j{}".format(name, i, j)
    return "{}:{}".format(base, present).encode("ascii")


def write_database(path, named_pairs):
    """named_pairs: iterable of (name, image_hwc, mask_hw[, i, j]) -> LMDB directory at `path`"""
    items = []
    for rec in named_pairs:
        name, img, mask = rec[0], rec[1], rec[2]
        mask = np.asarray(mask).astype(np.uint8)            # build_lmdb.py:151 forces uint8 masks
        ij = rec[3:5] if len(rec) >= 5 else (None, None)
        items.append((record_key(name, mask, *ij), encode_pair(img, mask)))
    lmdbfile.write(path, items)


def zscore_normalize(image_data, channels_first=True):
    """UNet/imagereader.py:33-66 (host version, for get_example(); the training path uses ub_zscore on the device)"""
    image_data = np.asarray(image_data).astype(np.float32)
    if image_data.ndim == 3:
        if not channels_first:
            image_data = image_data.transpose((2, 0, 1))
        image_data = image_data.copy()
        for c in range(image_data.shape[0]):
            std = np.std(image_data[c])
            mv = np.mean(image_data[c])
            image_data[c] = (image_data[c] - mv) if std <= 1.0 else (image_data[c] - mv) / std
        if not channels_first:
            image_data = image_data.transpose((1, 2, 0))
        return image_data
    if image_data.ndim == 2:
        std = np.std(image_data)
        mv = np.mean(image_data)
        return (image_data - mv) if std <= 1.0 else (image_data - mv) / std
    raise IOError("Input to Z-Score normalization needs to be either a 2D or 3D image [HW, or CHW]")


def imread(fp):
    from PIL import Image
    Image.MAX_IMAGE_PIXELS = None
    return np.asarray(Image.open(fp))


# ---------------------------------------------------------------------------------------------- reader
class ImageReader:
    # augmentation parameters of the reference reader (UNet/imagereader.py:78-85)
    _reflection_flag = True
    _rotation_flag = True
    _jitter_augmentation_severity = 0.1
    _noise_augmentation_severity = 0.02
    _scale_augmentation_severity = 0.1
    _blur_max_sigma = 2
    _intensity_augmentation_severity = None

    def __init__(self, img_db, use_augmentation=True, balance_classes=False, shuffle=True, num_workers=1, number_classes=2,
                 seed=None, rank=0, world_size=1):
        self.image_db = img_db
        self.use_augmentation = bool(use_augmentation)
        self.balance_classes = bool(balance_classes)
        self.shuffle = bool(shuffle)
        self.nb_workers = num_workers
        self.nb_classes = number_classes
        self.rank, self.world_size = rank, world_size
        self._rng = random.Random(seed)
        if not os.path.exists(self.image_db):
            print('Could not load database file: ')
            print(self.image_db)
            raise IOError("Missing Database")
        self._db = lmdbfile.Reader(self.image_db)
        print('Initializing image database')
        self.keys_flat = list(self._db.keys())
        if not self.keys_flat:
            raise IOError("Empty Database")
        first = decode_pair(self._db.get(self.keys_flat[0]))
        self.image_size = [first["img_height"], first["img_width"], first["channels"]]
        self.img_dtype = np.dtype(first["img_type"])
        for d in self.image_size[:2]:
            if d % SIZE_FACTOR != 0:
                raise IOError('Input Image tile height needs to be a multiple of 16 to allow integer sized downscaled feature maps. '
                              'Input images should be either HW or HWC dimension ordering')
        self.keys = [[]]
        if self.balance_classes:
            for key in self.keys_flat:
                for k in key.decode('ascii').split(':')[1].split(','):
                    k = int(k)
                    while len(self.keys) <= k:
                        self.keys.append([])
                    self.keys[k].append(key)
        print('Dataset has {} examples'.format(len(self.keys_flat)))
        if self.balance_classes:
            print('Dataset Example Count by Class:')
            for i in range(len(self.keys)):
                print('  class: {} count: {}'.format(i, len(self.keys[i])))
        self._np_rng = np.random.RandomState(seed if seed is None else (int(seed) + 7919 * rank) % (1 << 32))
        self._augmenter = None
        self.key_idx = rank          # un-shuffled readers stride through the keys (imagereader.py:237-241)
        self._pinned = None

    # reference API -------------------------------------------------------------------------------------------
    def get_image_count(self):
        return int(len(self.keys_flat))

    def get_image_size(self):
        return self.image_size

    def get_image_tensor_shape(self):
        return [self.image_size[2], self.image_size[0], self.image_size[1]]

    def get_label_tensor_shape(self):
        return [self.image_size[0], self.image_size[1]]

    def startup(self):          # the reference forks reader processes here; batches are assembled in-process instead
        self.key_idx = self.rank

    def shutdown(self):
        pass

    def _next_key(self):
        if self.shuffle:
            if self.balance_classes:
                while True:
                    label_idx = self._rng.randint(0, self.nb_classes - 1)
                    try:
                        nb_examples = len(self.keys[label_idx])
                    except IndexError as e:
                        print('ImageReader Error: Number of classes specified differs from number of observed classes in data')
                        raise e
                    if nb_examples > 0:
                        return self.keys[label_idx][self._rng.randint(0, nb_examples - 1)]
            return self.keys_flat[self._rng.randint(0, len(self.keys_flat) - 1)]
        fn = self.keys_flat[self.key_idx % len(self.keys_flat)]
        self.key_idx = (self.key_idx + max(self.world_size, 1)) % len(self.keys_flat)
        return fn

    def _load(self, key):
        d = decode_pair(self._db.get(key))
        h, w, c = d["img_height"], d["img_width"], d["channels"]
        img = np.frombuffer(d["image"], dtype=d["img_type"]).reshape(h, w, c)
        mask = np.frombuffer(d["mask"], dtype=d["mask_type"]).reshape(h, w)
        return img, mask

    def get_example(self):
        """the reference's queue item (imagereader.py:296-316): (float32 CHW z-scored, int32 one-hot HWK)"""
        img, mask = self._load(self._next_key())
        I = zscore_normalize(img.transpose((2, 0, 1)).astype(np.float32))
        M = mask.astype(np.int32).reshape(-1)
        fM = np.zeros((len(M), self.nb_classes), dtype=np.int32)
        try:
            fM[np.arange(len(M)), M] = 1
        except IndexError as e:
            print('ImageReader Error: Number of classes specified differs from number of observed classes in data')
            raise e
        return I, fM.reshape(mask.shape[0], mask.shape[1], self.nb_classes)

    def generator(self):
        while True:
            yield self.get_example()

    # B200 path -----------------------------------------------------------------------------------------------
    def next_raw_batch(self, batch_size, slot=0):
        """-> (images [B,C,H,W] in the stored dtype, labels uint8 [B,H,W]) in pinned host memory.  The buffers of a `slot` are
        reused; if an upload from them is still in flight (`mark_uploaded`) the host waits for it before overwriting them."""
        import torch
        H, W, C = self.image_size
        dt = self.img_dtype
        tdt = {np.dtype(np.uint8): torch.uint8, np.dtype(np.uint16): torch.int16, np.dtype(np.float32): torch.float32}.get(dt)
        if self._pinned is None:
            self._pinned, self._uploaded = {}, {}
        buf = self._pinned.get(slot)
        if buf is None or buf[0].shape[0] != batch_size:
            pin = torch.cuda.is_available()
            buf = (torch.empty((batch_size, C, H, W), dtype=tdt or torch.float32, pin_memory=pin),
                   torch.empty((batch_size, H, W), dtype=torch.uint8, pin_memory=pin))
            self._pinned[slot] = buf
        ev = self._uploaded.pop(slot, None)
        if ev is not None:
            ev.synchronize()
        xi, li = buf
        xv = xi.numpy()
        if tdt is torch.int16:
            xv = xv.view(np.uint16)
        lv = li.numpy()
        for b in range(batch_size):
            img, mask = self._load(self._next_key())
            if mask.size and int(mask.max()) >= self.nb_classes:
                print('ImageReader Error: Number of classes specified differs from number of observed classes in data')
                raise IndexError("label {} outside [0, {})".format(int(mask.max()), self.nb_classes))
            xv[b] = img.transpose((2, 0, 1)) if tdt is not None else img.transpose((2, 0, 1)).astype(np.float32)
            lv[b] = mask
        return xi, li

    def mark_uploaded(self, slot, event):
        """`event` (recorded after the async H2D copies of this slot's pinned buffers) guards their reuse"""
        self._uploaded[slot] = event

    def draw_augmentation(self, batch_size):
        """per-example augmentation parameters of one batch (UNet/imagereader.py:287-294 -> augment.py:61-153)"""
        from . import augment
        H, W, _ = self.image_size
        return augment.draw_params(self._np_rng, batch_size, H, W, self._rotation_flag, self._reflection_flag, self._jitter_augmentation_severity,
                                   self._noise_augmentation_severity, self._scale_augmentation_severity, self._blur_max_sigma,
                                   self._intensity_augmentation_severity)

    def device_batch(self, batch_size, unet_model, slot=0):
        """one training / test batch on the model's device, on the CURRENT stream: upload raw pixels + class indices, augment
        (if enabled), z-score.  -> (float32 [B,C,H,W] normalised images, uint8 [B,H,W] labels), the pair UNet.train_step
        consumes.  `slot` names the pinned staging buffers and the device output buffer (callers that keep two batches alive,
        unetb200.train's one-ahead prefetch, alternate slots)."""
        import torch
        xi, li = self.next_raw_batch(batch_size, slot)
        dev = unet_model.device
        raw, lab = xi.to(dev, non_blocking=True), li.to(dev, non_blocking=True)
        if dev.type == "cuda":
            ev = torch.cuda.Event()
            ev.record(torch.cuda.current_stream(dev))
            self.mark_uploaded(slot, ev)
        if self.use_augmentation:
            from . import augment
            if self._augmenter is None:
                self._augmenter = augment.DeviceAugmenter(dev, seed=int(self._np_rng.randint(0, 2 ** 31 - 1)))
            raw, lab = self._augmenter(raw, lab, self.draw_augmentation(batch_size))
        return unet_model.normalize_batch(raw, slot="{}:{}".format(id(self), slot)), lab

    def src_dtype_code(self):
        """ub_zscore source code of the stored pixels: 0 = u8, 1 = u16, 2 = f32"""
        return {np.dtype(np.uint8): 0, np.dtype(np.uint16): 1}.get(self.img_dtype, 2)
