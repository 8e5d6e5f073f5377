#!/usr/bin/env python
"""Golden vectors for the GPU geometric transforms (TEST INFRASTRUCTURE; run in the build container only).

Imports the UNMODIFIED reference transforms (/root/reference/code/dataset/transform.py: Resize :163-177,
RandomScaleCrop :180-204, RandomCrop :16-44) and records what they do to seeded synthetic samples:
the RNG state they leave behind, checksums and a 4x-subsampled copy of the outputs.  tests/test_transforms_gpu.py
rebuilds the same inputs and compares the CUDA transforms with these records bit for bit.

    python oracle/gen_transform_golden.py        ->  tests/golden/transforms.npz
"""
import hashlib
import os
import random
import sys

import numpy as np
from PIL import Image

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, "/root/reference/code")
sys.path.insert(0, os.path.join(ROOT, "tests"))
from dataset import transform as T   # noqa: E402  (the reference)
import fixture_data as FX              # noqa: E402

CASES = [(0, 300, 280), (1, 256, 256), (2, 200, 240), (3, 512, 400), (4, 180, 170), (5, 800, 800), (6, 257, 255)]


def make_sample(seed, h, w):
    rs = np.random.RandomState(1000 + seed)
    img = (FX._blobs(rs, h, w, 3) * 255).round().astype(np.uint8).transpose(1, 2, 0)
    yy, xx = np.mgrid[:h, :w]
    r = np.hypot(yy - h / 2, xx - w / 2)
    mask = np.full((h, w), 255, np.uint8)
    mask[r < min(h, w) * 0.30] = 128
    mask[r < min(h, w) * 0.15] = 0
    return np.ascontiguousarray(img), mask


def digest(a):
    return np.frombuffer(hashlib.sha1(np.ascontiguousarray(a).tobytes()).digest(), np.uint8)


def main():
    out = {}
    for seed, h, w in CASES:
        img, mask = make_sample(seed, h, w)
        for name, tf in (("resize", T.Resize((256, 256))), ("scalecrop", T.RandomScaleCrop((256, 256))),
                         ("both", None)):
            random.seed(500 + seed)
            s = {"img": Image.fromarray(img), "mask": Image.fromarray(mask)}
            if tf is None:                                   # train.py:541: Resize then RandomScaleCrop
                s = T.RandomScaleCrop((256, 256))(T.Resize((256, 256))(s))
            else:
                s = tf(s)
            oi, om = np.array(s["img"]), np.array(s["mask"])
            key = "%s_%d" % (name, seed)
            out[key + "_img_sha"], out[key + "_mask_sha"] = digest(oi), digest(om)
            out[key + "_img_sub"], out[key + "_mask_sub"] = oi[::4, ::4].copy(), om[::4, ::4].copy()
            out[key + "_shape"] = np.array(oi.shape[:2])
            out[key + "_after"] = np.array([random.random()])
    out["cases"] = np.array(CASES)
    path = os.path.join(ROOT, "tests", "golden", "transforms.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
