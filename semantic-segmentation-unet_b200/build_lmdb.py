"""`build_lmdb` of the reference (UNet/build_lmdb.py): two folders of images and masks -> `train-<name>.lmdb` and
`test-<name>.lmdb` holding `ImageMaskPair` records (isg_ai.proto) under the keys the reader expects.

  python -m unetb200.build_lmdb --image_folder DIR --mask_folder DIR --output_folder DIR --dataset_name HES
         [--train_fraction 0.8] [--image_format tif] [--use_tiling 0|1] [--tile_size 512]

Same arguments, file names and record layout as the reference; underneath, the records are encoded by
`imagereader.encode_pair` and the database file is written by `lmdbfile` (py-lmdb / liblmdb are not available here).
Host-side tooling: nothing in this module touches the GPU."""
from __future__ import annotations

import argparse
import os
import random
import shutil

import numpy as np

from . import imagereader, lmdbfile

SIZE_FACTOR = 16          # UNet/model.py:25
RADIUS = 96               # UNet/model.py:26


def read_image(fp):
    return imagereader.imread(fp)


def enforce_size_multiple(img):
    """crop (centred, UNet/build_lmdb.py:63-80) to height / width that are multiples of 16"""
    h, w = img.shape[0], img.shape[1]
    th, tw = (h // SIZE_FACTOR) * SIZE_FACTOR, (w // SIZE_FACTOR) * SIZE_FACTOR
    y0, x0 = int((h - th) / 2), int((w - tw) / 2)
    return img[y0:y0 + th, x0:x0 + tw]


def tile_origins(height, width, tile_size):
    """(y, x) of every tile of UNet/build_lmdb.py:83-130: a grid of stride tile_size - RADIUS whose last row / column is slid
    back so that the tile ends at the image edge.  (The reference shifts its column variable in place, so once a column has
    been slid back it stays there for the rest of that column -- reproduced here by clamping per column.)"""
    delta = int(tile_size - RADIUS)
    if delta <= 0:
        raise ValueError("tile_size must exceed the network radius ({})".format(RADIUS))
    if height < tile_size or width < tile_size:
        raise IOError("image {}x{} is smaller than the {} tile".format(height, width, tile_size))
    out = []
    for x in range(0, width, delta):
        xs = min(x, width - tile_size)
        for y in range(0, height, delta):
            out.append((min(y, height - tile_size), xs))
    return out


def process_slide_tiling(img, msk, tile_size, block_key):
    imgs, msks, keys = [], [], []
    for y, x in tile_origins(img.shape[0], img.shape[1], tile_size):
        m = msk[y:y + tile_size, x:x + tile_size]
        imgs.append(img[y:y + tile_size, x:x + tile_size])
        msks.append(m)
        keys.append(imagereader.record_key(block_key, m, y, x).decode("ascii"))
    return imgs, msks, keys


def generate_database(img_list, database_name, image_filepath, mask_filepath, output_folder, tile_size):
    out = os.path.join(output_folder, database_name)
    if os.path.exists(out):
        print('Deleting existing database')
        shutil.rmtree(out)
    items = {}          # key -> record; a repeated key replaces the earlier record, as txn.put does in the reference
    for i, fn in enumerate(img_list):
        print('  {}/{}'.format(i, len(img_list)))
        block_key = fn.replace('.tif', '')
        img = read_image(os.path.join(image_filepath, fn))
        msk = read_image(os.path.join(mask_filepath, fn)).astype(np.uint8)
        assert img.shape[0] == msk.shape[0], 'Image and Mask must be the same Height, input images should be either HW or HWC dimension ordering'
        assert img.shape[1] == msk.shape[1], 'Image and Mask must be the same Width, input images should be either HW or HWC dimension ordering'
        if tile_size > 0:
            for ti, tm, key in zip(*process_slide_tiling(img, msk, tile_size, block_key)):
                items[key.encode('ascii')] = imagereader.encode_pair(ti, tm)
        else:
            img, msk = enforce_size_multiple(img), enforce_size_multiple(msk)
            items[imagereader.record_key(block_key, msk)] = imagereader.encode_pair(img, msk)
    lmdbfile.write(out, list(items.items()))
    with open(os.path.join(out, 'img_filenames.csv'), 'w') as csvfile:
        for fn in img_list:
            csvfile.write(fn + '\n')


def main(image_folder, mask_folder, output_folder, dataset_name, train_fraction, image_format, use_tiling, tile_size, seed=None):
    if not use_tiling:
        tile_size = 0
    else:
        assert tile_size % SIZE_FACTOR == 0, 'UNet requires tiles with shapes that are multiples of 16'
    if image_format.startswith('.'):
        image_format = image_format[1:]
    image_folder, mask_folder, output_folder = (os.path.abspath(p) for p in (image_folder, mask_folder, output_folder))
    if not os.path.exists(output_folder):
        os.mkdir(output_folder)
    img_files = sorted(f for f in os.listdir(mask_folder) if f.endswith('.{}'.format(image_format)))
    if len(img_files) == 0:
        raise RuntimeError('Found no images')
    random.Random(seed).shuffle(img_files)           # the reference shuffles unseeded (build_lmdb.py:218); seed=None does the same
    idx = int(train_fraction * len(img_files))
    print('building train database')
    generate_database(img_files[:idx], 'train-{}.lmdb'.format(dataset_name), image_folder, mask_folder, output_folder, tile_size)
    print('building test database')
    generate_database(img_files[idx:], 'test-{}.lmdb'.format(dataset_name), image_folder, mask_folder, output_folder, tile_size)


if __name__ == "__main__":
    parser = argparse.ArgumentParser(prog='build_lmdb', description='Script which converts two folders of images and masks into a pair of lmdb databases for training.')
    parser.add_argument('--image_folder', dest='image_folder', type=str, help='filepath to the folder containing the images', default='../data/images/')
    parser.add_argument('--mask_folder', dest='mask_folder', type=str, help='filepath to the folder containing the masks', default='../data/masks/')
    parser.add_argument('--output_folder', dest='output_folder', type=str, help='filepath to the folder where the outputs will be placed', default='../data/')
    parser.add_argument('--dataset_name', dest='dataset_name', type=str, help='name of the dataset to be used in creating the lmdb files', default='HES')
    parser.add_argument('--train_fraction', dest='train_fraction', type=float, help='what fraction of the dataset to use for training (0.0, 1.0)', default=0.8)
    parser.add_argument('--image_format', dest='image_format', type=str, help='format (extension) of the input images. E.g {tif, jpg, png)', default='tif')
    parser.add_argument('--use_tiling', dest='use_tiling', type=int, help='Whether to shard the image into tiles [0 = False, 1 = True]', default=0)
    parser.add_argument('--tile_size', dest='tile_size', type=int, help='The size of the tiles to crop out of the source images, striding across all available pixels in the source images', default=512)
    a = parser.parse_args()
    main(a.image_folder, a.mask_folder, a.output_folder, a.dataset_name, a.train_fraction, a.image_format, a.use_tiling, a.tile_size)
