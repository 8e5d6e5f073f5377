"""Per-sample drop-ins with the reference's exact signatures.

    extract_amp_spectrum(img_np)                       fundus.py:13-19  (= prostate.py:10-16)
    low_freq_mutate_np(amp_src, amp_trg, L=0.1)        fundus.py:21-39
    source_to_target_freq(src_img, amp_trg, L=0.1)     fundus.py:41-61

numpy in, numpy out, new arrays returned, inputs not mutated, and -- like the reference --
each mix draws exactly one `random.randint(1, 10)` from Python's GLOBAL `random` module
(fundus.py:35) so downstream RNG streams are unchanged.  Host arrays are staged through the
device per call: this is the strict-compatibility shim (use it with num_workers=0); the fast path
is `ramdsir_b200.api.ram_mix_batch`.  All arithmetic runs in the CUDA library in float32; results
are returned in the input's precision class (float64 in -> float64 out), as numpy 1.x did.
"""
from __future__ import annotations

import random

import numpy as np
import torch

from . import _lib
from .api import NO_EPILOGUE, band_halfwidth, ram_mix_batch

__all__ = ["extract_amp_spectrum", "low_freq_mutate_np", "source_to_target_freq"]


def _out_dtype(a):
    return np.float64 if np.asarray(a).dtype == np.float64 else np.float32


def _dev():
    if not torch.cuda.is_available():
        raise RuntimeError("ramdsir_b200 needs a CUDA device; there is no CPU fallback")
    return torch.device("cuda", torch.cuda.current_device())


def _to_dev(a):
    return torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).to(_dev())


def extract_amp_spectrum(img_np):
    """|fft2(img, axes=(-2, -1))| of a CxHxW array; full un-shifted spectrum like the reference."""
    img = np.asarray(img_np)
    if img.ndim != 3:
        raise ValueError("extract_amp_spectrum expects a CxHxW array, got shape %s" % (img.shape,))
    c, h, w = img.shape
    lib = _lib.load()
    x = _to_dev(img)
    amp = torch.empty_like(x)
    nbytes = int(lib.ramdsir_amp_spectrum_workspace_bytes(1, c, h, w))
    ws = torch.empty(max(nbytes, 1), dtype=torch.uint8, device=x.device)
    _lib.check(lib.ramdsir_amp_spectrum_f32(x.data_ptr(), amp.data_ptr(), 1, c, h, w, ws.data_ptr(), nbytes,
                                            int(torch.cuda.current_stream().cuda_stream)))
    return amp.cpu().numpy().astype(_out_dtype(img), copy=False)


def low_freq_mutate_np(amp_src, amp_trg, L=0.1):
    """Blend the centred low-frequency square of two CxHxW amplitude spectra; draws the ratio."""
    a_src, a_trg = np.asarray(amp_src), np.asarray(amp_trg)
    if a_src.ndim != 3 or a_src.shape != a_trg.shape:
        raise ValueError("amp_src and amp_trg must be CxHxW arrays of the same shape, got %s and %s"
                         % (a_src.shape, a_trg.shape))
    c, h, w = a_src.shape
    b = band_halfwidth(h, w, L)
    ratio_k = random.randint(1, 10)          # fundus.py:35 -- same call, same global stream
    lib = _lib.load()
    s, t = _to_dev(a_src), _to_dev(a_trg)
    out = torch.empty_like(s)
    _lib.check(lib.ramdsir_band_blend_f32(s.data_ptr(), t.data_ptr(), out.data_ptr(), c, h, w, b, ratio_k, 10,
                                          int(torch.cuda.current_stream().cuda_stream)))
    return out.cpu().numpy().astype(_out_dtype(a_src), copy=False)


def source_to_target_freq(src_img, amp_trg, L=0.1):
    """RAM of one HxWxC image against a partner amplitude spectrum CxHxW; returns HxWxC, unclipped."""
    src, amp = np.asarray(src_img), np.asarray(amp_trg)
    if src.ndim != 3:
        raise ValueError("source_to_target_freq expects an HxWxC image, got shape %s" % (src.shape,))
    h, w, c = src.shape
    if amp.shape != (c, h, w):
        raise ValueError("operands could not be broadcast together: src CHW %s vs amp_trg %s"
                         % ((c, h, w), amp.shape))
    b = band_halfwidth(h, w, L)
    ch, cw = h // 2, w // 2                   # floor(h / 2.0), fundus.py:27-28
    if b > ch or b > cw or 2 * b + 1 > min(h, w):
        raise ValueError("band half-width %d does not fit a %dx%d spectrum" % (b, h, w))
    ratio_k = random.randint(1, 10)          # fundus.py:35
    # the only part of amp_trg the reference reads: fftshift(amp)[:, h1:h2, w1:w2] (fundus.py:23,30-37)
    patch = np.fft.fftshift(amp, axes=(-2, -1))[:, ch - b:ch + b + 1, cw - b:cw + b + 1]
    x = _to_dev(src.transpose(2, 0, 1))[None]
    bank = _to_dev(patch)[None]
    y = ram_mix_batch(x, bank, [0], [ratio_k], L=L, clip=NO_EPILOGUE[:2])
    return y[0].cpu().numpy().astype(_out_dtype(src), copy=False).transpose(1, 2, 0)
