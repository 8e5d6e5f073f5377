"""GPU versions of the reference loader's geometric transforms, bit-exact with Pillow.

    Resize(target_size)        reference: code/dataset/transform.py:163-177
    RandomScaleCrop(size)      reference: code/dataset/transform.py:180-204
    RandomCrop(output_size)    reference: code/dataset/transform.py:16-44
    resize_u8(...)             the partner's Image.open(...).resize((256, 256), BILINEAR), fundus.py:209

A sample is {'img': uint8 CUDA tensor [H, W, 3] (np.array(PIL RGB image) layout), 'mask': uint8 CUDA tensor [H, W]}.
The classes keep the reference's constructor arguments and consume draws from Python's `random` in exactly the
reference's order (random.random(); 2 x random.uniform when it exceeds 0.5; 2 x random.randint in the crop), so the
RNG stream that RAM's own draw comes from afterwards is unchanged.  The pixel work -- Pillow's fixed-point bilinear
resampling, its nearest-neighbour index table, the pad and crop -- runs in libramdsir.so (ramdsir_resize_u8,
ramdsir_crop_pad_chw_u8) and reproduces Pillow's bytes exactly.
"""
from __future__ import annotations

import random

import torch

from . import _lib

__all__ = ["resize_u8", "crop_pad_u8", "Resize", "RandomCrop", "RandomScaleCrop", "to_chw"]

NEAREST, BILINEAR = _lib.FILTER_NEAREST, _lib.FILTER_BILINEAR


def _check_img(t, name):
    if not (isinstance(t, torch.Tensor) and t.is_cuda and t.dtype == torch.uint8 and t.dim() in (2, 3)):
        raise TypeError("%s must be a CUDA uint8 tensor [H, W] or [H, W, C]" % name)
    return t.contiguous()


def resize_u8(img: torch.Tensor, size, resample: int = BILINEAR, stream=None) -> torch.Tensor:
    """PIL semantics: `size` is (width, height).  img [H, W] or [H, W, C] uint8 CUDA -> same rank, resized."""
    img = _check_img(img, "img")
    wout, hout = int(size[0]), int(size[1])
    hin, win = img.shape[0], img.shape[1]
    c = 1 if img.dim() == 2 else img.shape[2]
    lib = _lib.load()
    out = torch.empty((hout, wout) if img.dim() == 2 else (hout, wout, c), dtype=torch.uint8, device=img.device)
    with torch.cuda.device(img.device):
        s = torch.cuda.current_stream() if stream is None else stream
        with torch.cuda.stream(s):
            nbytes = int(lib.ramdsir_resize_workspace_bytes(hin, win, c, hout, wout))
            ws = torch.empty(nbytes, dtype=torch.uint8, device=img.device)
            _lib.check(lib.ramdsir_resize_u8(img.data_ptr(), hin, win, c, out.data_ptr(), hout, wout, int(resample),
                                            ws.data_ptr(), nbytes, int(s.cuda_stream)))
    return out


def crop_pad_u8(img: torch.Tensor, x: int, y: int, out_w: int, out_h: int, fill: int = 0, chw: bool = False, stream=None):
    """Crop the window [x, x + out_w) x [y, y + out_h); pixels outside the image are `fill` (RandomCrop pads right and
    bottom before cropping, transform.py:27-31).  Returns [out_h, out_w(, C)], or planar [C, out_h, out_w] with chw=True."""
    img = _check_img(img, "img")
    h, w = img.shape[0], img.shape[1]
    c = 1 if img.dim() == 2 else img.shape[2]
    out = torch.empty((c, out_h, out_w), dtype=torch.uint8, device=img.device)
    with torch.cuda.device(img.device):
        s = torch.cuda.current_stream() if stream is None else stream
        _lib.check(_lib.load().ramdsir_crop_pad_chw_u8(img.data_ptr(), h, w, c, int(x), int(y), int(out_h), int(out_w),
                                                      int(fill), out.data_ptr(), int(s.cuda_stream)))
    if chw:
        return out
    return out[0] if img.dim() == 2 else out.permute(1, 2, 0).contiguous()


def to_chw(img: torch.Tensor) -> torch.Tensor:
    """[H, W, C] -> planar [C, H, W] (what the RAM kernels read; fundus.py:211-212 transposes on the host)."""
    img = _check_img(img, "img")
    return crop_pad_u8(img, 0, 0, img.shape[1], img.shape[0], chw=True)


class Resize:
    def __init__(self, target_size):
        self.target_size = target_size

    def __call__(self, sample):
        size = (self.target_size[0], self.target_size[1])
        return {"img": resize_u8(sample["img"], size, BILINEAR), "mask": resize_u8(sample["mask"], size, NEAREST)}


class RandomCrop:
    def __init__(self, output_size):
        self.output_size = output_size

    def __call__(self, sample):
        img, mask = sample["img"], sample["mask"]
        h, w = img.shape[0], img.shape[1]
        padw = self.output_size[0] - w if w < self.output_size[0] else 0
        padh = self.output_size[1] - h if h < self.output_size[1] else 0
        w, h = w + padw, h + padh                                   # ImageOps.expand(border=(0, 0, padw, padh))
        x = random.randint(0, w - self.output_size[0])
        y = random.randint(0, h - self.output_size[1])
        ow, oh = self.output_size[0], self.output_size[1]
        return {"img": crop_pad_u8(img, x, y, ow, oh, fill=0), "mask": crop_pad_u8(mask, x, y, ow, oh, fill=255)}


class RandomScaleCrop:
    def __init__(self, size):
        self.size = size
        self.crop = RandomCrop(self.size)

    def __call__(self, sample):
        img, mask = sample["img"], sample["mask"]
        assert img.shape[0] == mask.shape[0] and img.shape[1] == mask.shape[1]
        seed = random.random()
        if seed > 0.5:
            w = int(random.uniform(1, 1.5) * img.shape[1])
            h = int(random.uniform(1, 1.5) * img.shape[0])
            sample = {"img": resize_u8(img, (w, h), BILINEAR), "mask": resize_u8(mask, (w, h), NEAREST)}
        return self.crop(sample)
