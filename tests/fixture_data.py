"""Deterministic synthetic on-disk fixtures in the reference's dataset layouts.

fundus:   <root>/DomainK_train.list          lines "DomainK/train/ROIs/image/x.png DomainK/train/ROIs/mask/x.png"  (fundus.py:143,164)
          <root>/DomainK/train.list          same paths relative to the domain dir                                 (fundus.py:206-209)
          images RGB 8-bit (varying sizes), masks 8-bit with {0, 128, 255} regions                                 (fundus.py:227-235)
prostate: <root>/DomainK/image/*.npy  H x W x 3 float32 in [-1, 1];  <root>/DomainK/mask/*.npy  H x W int         (prostate.py:134,149,153)
Used by oracle/gen_golden.py (build container, reference mounted) and by the tests (anywhere).
"""
import os

import numpy as np
from PIL import Image

FUNDUS_DOMAINS = {"Domain1": 3, "Domain2": 4, "Domain3": 3, "Domain4": 2}     # images per domain
PROSTATE_DOMAINS = {"Domain%d" % (i + 1): 3 for i in range(6)}


def _blobs(rs, h, w, c):
    z = rs.standard_normal((c, h, w))
    fy = np.fft.fftfreq(h)[:, None]
    fx = np.fft.fftfreq(w)[None, :]
    y = np.real(np.fft.ifft2(np.fft.fft2(z) / (1.0 + (np.hypot(fy, fx) / 0.03) ** 2)))
    y = (y - y.min()) / (y.max() - y.min())
    return y


def make_fundus(root, counts=None):
    """`counts` (domain -> images) defaults to FUNDUS_DOMAINS, the fixture the dataset goldens were recorded on."""
    rs = np.random.RandomState(20221017)
    for dom, n in (FUNDUS_DOMAINS if counts is None else counts).items():
        os.makedirs(os.path.join(root, dom, "train", "ROIs", "image"), exist_ok=True)
        os.makedirs(os.path.join(root, dom, "train", "ROIs", "mask"), exist_ok=True)
        top, inner = [], []
        for i in range(n):
            size = int(rs.choice([256, 300, 384]))
            img = (_blobs(rs, size, size, 3) * 255).round().astype(np.uint8).transpose(1, 2, 0)
            yy, xx = np.mgrid[:size, :size]
            r = np.hypot(yy - size / 2 + rs.randint(-20, 20), xx - size / 2 + rs.randint(-20, 20))
            mask = np.full((size, size), 255, np.uint8)
            mask[r < size * 0.30] = 128
            mask[r < size * 0.15] = 0
            name = "%s_%02d.png" % (dom.lower(), i)
            Image.fromarray(img).save(os.path.join(root, dom, "train", "ROIs", "image", name))
            Image.fromarray(mask).save(os.path.join(root, dom, "train", "ROIs", "mask", name))
            top.append("%s/train/ROIs/image/%s %s/train/ROIs/mask/%s" % (dom, name, dom, name))
            inner.append("train/ROIs/image/%s train/ROIs/mask/%s" % (name, name))
        with open(os.path.join(root, dom + "_train.list"), "w") as f:
            f.write("\n".join(top) + "\n")
        with open(os.path.join(root, dom, "train.list"), "w") as f:
            f.write("\n".join(inner) + "\n")
    return root


def make_prostate(root, size=384):
    rs = np.random.RandomState(20221018)
    for dom, n in PROSTATE_DOMAINS.items():
        os.makedirs(os.path.join(root, dom, "image"), exist_ok=True)
        os.makedirs(os.path.join(root, dom, "mask"), exist_ok=True)
        for i in range(n):
            img = (_blobs(rs, size, size, 3) * 2 - 1).astype(np.float32).transpose(1, 2, 0)
            yy, xx = np.mgrid[:size, :size]
            mask = (np.hypot(yy - size / 2, xx - size / 2) < size * 0.2).astype(np.int64)
            np.save(os.path.join(root, dom, "image", "s%02d.npy" % i), img)
            np.save(os.path.join(root, dom, "mask", "s%02d.npy" % i), mask)
    return root


class ResizeJitter:
    """Stand-in for the reference's PIL transforms (Resize + RandomScaleCrop, train.py:541 -- out of scope):
    consumes draws from Python's `random` BEFORE the RAM block, like the real transforms do, then resizes
    the sample to 256 x 256 (image bilinear, mask nearest)."""

    def __call__(self, sample):
        import random
        img, mask = sample["img"], sample["mask"]
        if random.random() > 0.5:
            s = random.uniform(1.0, 1.2)
            w, h = int(img.size[0] * s), int(img.size[1] * s)
            img, mask = img.resize((w, h), Image.BILINEAR), mask.resize((w, h), Image.NEAREST)
        return {"img": img.resize((256, 256), Image.BILINEAR), "mask": mask.resize((256, 256), Image.NEAREST)}
