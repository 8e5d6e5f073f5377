"""CPU oracle for Random Amplitude Mixup (RAM) -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

This module is a numpy restatement of the reference's per-sample RAM path
(`/root/reference/code/dataset/fundus.py:13-61`, byte-identical in
`code/dataset/prostate.py:10-62`) plus the caller epilogues
(`fundus.py:215,222-223`, `prostate.py:188`).  It exists only to CHECK the CUDA
path.  Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s `cpu_baseline` /
`--impl reference` legs may import it; nothing under `ramdsir_b200/` does, and the
product path raises when the CUDA library is missing instead of coming here.

Parity pin: the reference ships no tests or golden vectors for this path (SURVEY.md
section 4), so the oracle is pinned against OUTPUTS OF THE REFERENCE ITSELF: the
script `oracle/gen_golden.py` imports the unmodified reference functions from
`/root/reference/code` in the build container and writes `tests/golden/*.npz`;
`tests/test_oracle_golden.py` replays those fixtures through this file.

The arithmetic itself lives in numpy (pocketfft `np.fft.fft2/ifft2`, `np.abs`,
`np.angle`, `np.exp`), which the reference neither vendors nor pins; the call sites
are `fundus.py:16,17,22,23,38,46,49,55,58,59`.  numpy >= 2 keeps float32 input in
complex64, numpy 1.x (what the reference was written against) promoted to
complex128; `precision="f64"` (default) reproduces the latter by casting inputs up,
which is the "numpy float64 path" that BASELINE.json's tolerance is stated against.
"""
from __future__ import annotations

import math
import random as _py_random

import numpy as np

__all__ = [
    "band_halfwidth", "band_window", "amp_spectrum", "blend_low_band", "mix_sample",
    "mix_sample_closed_form", "low_band_patch", "draw_ratio_k", "mix_batch",
    "build_bank", "epilogue", "FUNDUS_EPILOGUE", "PROSTATE_EPILOGUE",
]

# (clip_lo, clip_hi, scale, shift): y = clip(v, lo, hi) * scale + shift
FUNDUS_EPILOGUE = (0.0, 255.0, 1.0 / 127.5, -1.0)   # fundus.py:215,222-223
PROSTATE_EPILOGUE = (-1.0, 1.0, 1.0, 0.0)            # prostate.py:188


def band_halfwidth(h: int, w: int, L: float = 0.1) -> int:
    """b = floor(min(h, w) * L), evaluated in float64 like fundus.py:26."""
    return int(np.floor(np.amin((h, w)) * L))


def band_window(h: int, w: int, L: float = 0.1):
    """Inclusive-exclusive slice bounds of the centred square in SHIFTED coordinates.

    fundus.py:25-33: centre = floor(n / 2), rows/cols [c - b, c + b + 1).
    """
    b = band_halfwidth(h, w, L)
    ch = int(np.floor(h / 2.0))
    cw = int(np.floor(w / 2.0))
    return ch - b, ch + b + 1, cw - b, cw + b + 1


def _cast(a, precision):
    a = np.asarray(a)
    if precision == "f64":
        return a.astype(np.float64, copy=False)
    if precision == "native":
        return a
    raise ValueError("precision must be 'f64' or 'native'")


def amp_spectrum(img_chw, precision="f64"):
    """|fft2| over the last two axes of a CHW array (fundus.py:13-19).

    The reference also evaluates np.angle and discards it; that has no effect on the
    result and is omitted here.
    """
    return np.abs(np.fft.fft2(_cast(img_chw, precision), axes=(-2, -1)))


def blend_low_band(amp_src, amp_trg, L=0.1, ratio=1.0):
    """Low-frequency amplitude blend with an explicit `ratio` (fundus.py:21-39).

    The reference draws `ratio = random.randint(1, 10) / 10` inside this function
    (fundus.py:35); the oracle takes it as an argument so both arms of a parity test
    see the same value (use `draw_ratio_k` for the draw).  `ratio` weights the SOURCE
    amplitude, `1 - ratio` the partner's (fundus.py:37).
    """
    s = np.fft.fftshift(amp_src, axes=(-2, -1))   # new array; inputs are not mutated
    t = np.fft.fftshift(amp_trg, axes=(-2, -1))
    _, h, w = s.shape
    h1, h2, w1, w2 = band_window(h, w, L)
    s[:, h1:h2, w1:w2] = s[:, h1:h2, w1:w2] * ratio + t[:, h1:h2, w1:w2] * (1 - ratio)
    return np.fft.ifftshift(s, axes=(-2, -1))


def draw_ratio_k(rng=_py_random) -> int:
    """The reference's only RNG draw on this path: random.randint(1, 10) (fundus.py:35)."""
    return rng.randint(1, 10)


def mix_sample(src_hwc, amp_trg_chw, L=0.1, ratio=1.0, precision="f64"):
    """source_to_target_freq with explicit ratio (fundus.py:41-61).  HWC in, HWC out."""
    x = _cast(src_hwc, precision).transpose(2, 0, 1)
    f = np.fft.fft2(x, axes=(-2, -1))
    amp, pha = np.abs(f), np.angle(f)
    amp2 = blend_low_band(amp, _cast(amp_trg_chw, precision), L=L, ratio=ratio)
    y = np.fft.ifft2(amp2 * np.exp(1j * pha), axes=(-2, -1))
    return np.real(y).transpose(1, 2, 0)


def low_band_patch(amp_chw, L=0.1):
    """The only part of a partner amplitude the blend reads: fftshift(amp)[:, h1:h2, w1:w2].

    This is the amplitude-bank layout: [C, 2b+1, 2b+1], centre element = DC.
    """
    _, h, w = amp_chw.shape
    h1, h2, w1, w2 = band_window(h, w, L)
    return np.ascontiguousarray(np.fft.fftshift(amp_chw, axes=(-2, -1))[:, h1:h2, w1:w2])


def mix_sample_closed_form(src_chw, patch, b, ratio):
    """y = x + Re(ifft2(dF)) with dF supported on the band only.

    Independent second derivation used to cross-check `mix_sample` and as the model of
    what the CUDA kernels compute.  `patch` is [C, 2b+1, 2b+1] in the `low_band_patch`
    layout.  dF = F * (1 - ratio) * (A_t / |F| - 1); where |F| == 0 the phase is 0
    (np.angle(0) == 0) so dF = (1 - ratio) * A_t.
    """
    x = np.asarray(src_chw, dtype=np.float64)
    c, h, w = x.shape
    f = np.fft.fft2(x, axes=(-2, -1))
    k = np.arange(-b, b + 1)
    kh = k % h
    kw = k % w
    fb = f[:, kh[:, None], kw[None, :]]
    a = np.abs(fb)
    at = np.asarray(patch, dtype=np.float64)
    with np.errstate(divide="ignore", invalid="ignore"):
        d = np.where(a > 0, fb * (1 - ratio) * (at / a - 1), (1 - ratio) * at)
    df = np.zeros_like(f)
    df[:, kh[:, None], kw[None, :]] = d
    return x + np.real(np.fft.ifft2(df, axes=(-2, -1)))


def epilogue(v, clip_lo, clip_hi, scale, shift):
    """Caller-side clip + normalise (fundus.py:215,222-223; prostate.py:188), in float32
    like the reference (`astype(np.float32)` then in-place `/=` and `-=`)."""
    y = np.clip(v, clip_lo, clip_hi).astype(np.float32)
    if scale != 1.0:
        # fundus does `/= 127.5`; scale is passed as 1/127.5 and applied as a division
        # when it is that reciprocal, to stay bit-faithful to the reference
        inv = 1.0 / scale
        y /= np.float32(inv)
    if shift != 0.0:
        y += np.float32(shift)
    return y


def build_bank(imgs_nchw, L=0.1, precision="f64"):
    """Amplitude bank: per source image, the low-band patch of its amplitude spectrum."""
    return np.stack([low_band_patch(amp_spectrum(im, precision), L) for im in imgs_nchw])


def mix_batch(src_nchw, partners_nchw, partner_idx, ratio_k, L=0.1,
              epi=None, precision="f64"):
    """Batched driver: per sample exactly what `Fundus_Multi.__getitem__` does between
    fundus.py:213 and :223 -- extract_amp_spectrum(partner) -> source_to_target_freq ->
    clip -> normalise -- with partner index and ratio numerator supplied by the caller.
    Returns NCHW (float64 if `epi` is None, else float32 after the epilogue).
    """
    out = []
    for i in range(len(src_nchw)):
        amp = amp_spectrum(partners_nchw[int(partner_idx[i])], precision)
        y = mix_sample(np.asarray(src_nchw[i]).transpose(1, 2, 0), amp, L=L,
                       ratio=int(ratio_k[i]) / 10, precision=precision)
        y = y.transpose(2, 0, 1)
        if epi is not None:
            y = epilogue(y, *epi)
        out.append(y)
    return np.stack(out)


def rel_l2(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.linalg.norm((a - b).ravel()) / max(np.linalg.norm(b.ravel()), 1e-300))


def max_abs(a, b):
    return float(np.max(np.abs(np.asarray(a, dtype=np.float64) - np.asarray(b, dtype=np.float64))))
