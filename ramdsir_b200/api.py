"""Batched torch-tensor entry points of the B200 RAM path.

`build_amp_bank` replaces a loop of `extract_amp_spectrum` calls
(/root/reference/code/dataset/fundus.py:13-19) and keeps only what the blend reads: the
centred low-frequency square `fftshift(amp)[:, cH-b:cH+b+1, cW-b:cW+b+1]` (fundus.py:22-33).
`ram_mix_batch` replaces `source_to_target_freq` + `low_freq_mutate_np` + the caller's
clip/normalise (fundus.py:21-61, 215, 222-223; prostate.py:188) for a whole NCHW batch.

Partner indices and the ratio numerators are INPUTS: they are drawn on the host by the same
RNG calls the reference makes (see ramdsir_b200.rng) so selection is bit-identical.
All heavy lifting happens in libramdsir.so (hand-written sm_100a kernels); torch is only used
for device memory and streams.
"""
from __future__ import annotations

import math

import numpy as np
import torch

from . import _lib

FUNDUS_EPILOGUE = (0.0, 255.0, 1.0 / 127.5, -1.0)   # fundus.py:215,222-223
PROSTATE_EPILOGUE = (-1.0, 1.0, 1.0, 0.0)            # prostate.py:188
NO_EPILOGUE = (-math.inf, math.inf, 1.0, 0.0)

__all__ = ["band_halfwidth", "build_amp_bank", "ram_mix_batch", "kernel_path", "workspace_bytes",
           "FUNDUS_EPILOGUE", "PROSTATE_EPILOGUE", "NO_EPILOGUE"]


def band_halfwidth(h: int, w: int, L: float = 0.1) -> int:
    """b = floor(min(h, w) * L) in float64 (fundus.py:26)."""
    return int(_lib.load().ramdsir_band_halfwidth(int(h), int(w), float(L)))


def kernel_path(c: int, h: int, w: int, b: int) -> int:
    return int(_lib.load().ramdsir_path(int(c), int(h), int(w), int(b)))


def workspace_bytes(n: int, c: int, h: int, w: int, b: int) -> int:
    return int(_lib.load().ramdsir_workspace_bytes(int(n), int(c), int(h), int(w), int(b)))


def _stream_ptr(stream):
    s = torch.cuda.current_stream() if stream is None else stream
    return int(s.cuda_stream)


def _need_cuda_f32(t, name, ndim):
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise TypeError("%s must be a CUDA torch.Tensor (the RAM path has no CPU fallback)" % name)
    if t.dtype != torch.float32:
        raise TypeError("%s must be float32, got %s" % (name, t.dtype))
    if t.dim() != ndim:
        raise ValueError("%s must have %d dims, got shape %s" % (name, ndim, tuple(t.shape)))
    if not t.is_contiguous():
        raise ValueError("%s must be contiguous (NCHW)" % name)
    return t


def u8_supported(c: int, h: int, w: int, b: int) -> bool:
    """True when uint8 sources of this shape are read directly by the RAM kernel (ramdsir_u8_supported)."""
    return bool(_lib.load().ramdsir_u8_supported(int(c), int(h), int(w), int(b)))


def _need_cuda_src(t, name):
    """Sources may be float32 or uint8 (what the reference's loaders start from, fundus.py:205-214)."""
    if isinstance(t, torch.Tensor) and t.is_cuda and t.dtype == torch.uint8:
        if t.dim() != 4:
            raise ValueError("%s must have 4 dims, got shape %s" % (name, tuple(t.shape)))
        if not t.is_contiguous():
            raise ValueError("%s must be contiguous (NCHW)" % name)
        return t
    return _need_cuda_f32(t, name, 4)


def _workspace(nbytes, device, workspace):
    if nbytes == 0:
        return None, 0, 0
    if workspace is None or workspace.numel() * workspace.element_size() < nbytes:
        workspace = torch.empty(nbytes, dtype=torch.uint8, device=device)
    return workspace, workspace.data_ptr(), workspace.numel() * workspace.element_size()


def build_amp_bank(imgs: torch.Tensor, L: float = 0.1, out: torch.Tensor | None = None, stream=None,
                   workspace: torch.Tensor | None = None) -> torch.Tensor:
    """imgs [N, C, H, W] float32 (or uint8) CUDA -> bank [N, C, 2b+1, 2b+1] float32 (K3, amplitude-bank builder).

    bank[n, c] == np.fft.fftshift(extract_amp_spectrum(imgs[n]))[c, cH-b:cH+b+1, cW-b:cW+b+1].
    `out` may be a slice of a larger (all-gather) buffer so the collective needs no staging copy.
    """
    imgs = _need_cuda_src(imgs, "imgs")
    n, c, h, w = imgs.shape
    b = band_halfwidth(h, w, L)
    nb = 2 * b + 1
    if imgs.dtype == torch.uint8 and not u8_supported(c, h, w, b):
        imgs = imgs.float()          # device-side conversion; only the static fused kernels read uint8 directly
    if out is None:
        out = torch.empty((n, c, nb, nb), dtype=torch.float32, device=imgs.device)
    else:
        _need_cuda_f32(out, "out", 4)
        if tuple(out.shape) != (n, c, nb, nb):
            raise ValueError("out must have shape %s, got %s" % ((n, c, nb, nb), tuple(out.shape)))
    lib = _lib.load()
    if imgs.dtype == torch.uint8:
        with torch.cuda.device(imgs.device):
            _lib.check(lib.ramdsir_bank_build_u8(imgs.data_ptr(), out.data_ptr(), n, c, h, w, b, _stream_ptr(stream)))
        return out
    with torch.cuda.device(imgs.device):
        ws, ws_ptr, ws_n = _workspace(workspace_bytes(n, c, h, w, b), imgs.device, workspace)
        _lib.check(lib.ramdsir_bank_build_f32(imgs.data_ptr(), out.data_ptr(), n, c, h, w, b, ws_ptr, ws_n,
                                              _stream_ptr(stream)))
        if ws is not None:
            ws.record_stream(torch.cuda.current_stream() if stream is None else stream)
    return out


def _index_tensor(x, dtype, device, name, lo=None, hi=None):
    """Host-side validation (when the values live on the host) + upload."""
    if isinstance(x, torch.Tensor) and x.is_cuda:
        return x.to(dtype=dtype).contiguous()
    arr = np.asarray(x.cpu() if isinstance(x, torch.Tensor) else x)
    if arr.ndim != 1:
        raise ValueError("%s must be 1-D" % name)
    if arr.size and lo is not None and (arr.min() < lo or arr.max() > hi):
        raise ValueError("%s values must lie in [%d, %d], got [%d, %d]" % (name, lo, hi, arr.min(), arr.max()))
    npdt = np.int64 if dtype == torch.int64 else np.int32
    return torch.from_numpy(np.ascontiguousarray(arr.astype(npdt))).to(device, non_blocking=False)


def ram_mix_batch(src: torch.Tensor, bank: torch.Tensor, partner_idx, ratio_k, L: float = 0.1,
                  clip=(-math.inf, math.inf), scale: float = 1.0, shift: float = 0.0,
                  out: torch.Tensor | None = None, ratio_den: int = 10, stream=None,
                  workspace: torch.Tensor | None = None, independent: bool = False,
                  clean_out: torch.Tensor | None = None) -> torch.Tensor:
    """Random Amplitude Mixup of a batch (K1/K2).

    src [B, C, H, W] float32 CUDA, or uint8 (read directly by the kernel on the static shapes, bit-identical to
    passing src.float(); converted on the device otherwise); bank [N, C, 2b+1, 2b+1] from `build_amp_bank`;
    partner_idx [B] int (indices into the bank); ratio_k [B] int, ratio = ratio_k / ratio_den is the
    weight of the SOURCE amplitude (fundus.py:35,37; the reference draws ratio_k = randint(1, 10)).
    Returns clamp(RAM(src), clip) * scale + shift, float32 [B, C, H, W].

    clean_out (optional, same shape as src): also receives the normalised clean batch src * scale + shift
    (fundus.py:217-218), fused into the same kernel on the static shapes (ramdsir_mix_dual_f32).

    independent=True promises that this call neither reads nor overwrites anything the previous kernel on
    the stream produced (back-to-back batches with their own buffers); the launch may then overlap the tail
    of the previous RAM launch (RAMDSIR_FLAG_INDEPENDENT in include/ramdsir.h).
    """
    src = _need_cuda_src(src, "src")
    bank = _need_cuda_f32(bank, "bank", 4)
    bsz, c, h, w = src.shape
    b = band_halfwidth(h, w, L)
    nb = 2 * b + 1
    if src.dtype == torch.uint8 and not u8_supported(c, h, w, b):
        src = src.float()
    if tuple(bank.shape[1:]) != (c, nb, nb):
        raise ValueError("bank must be [N, %d, %d, %d] for src %s and L=%g, got %s"
                         % (c, nb, nb, tuple(src.shape), L, tuple(bank.shape)))
    if bank.device != src.device:
        raise ValueError("src and bank must live on the same device")
    n_bank = bank.shape[0]
    if out is None:
        out = torch.empty(src.shape, dtype=torch.float32, device=src.device)
    else:
        _need_cuda_f32(out, "out", 4)
        if out.shape != src.shape or out.data_ptr() == src.data_ptr():
            raise ValueError("out must match src's shape and must not alias it")
    if clean_out is not None:
        _need_cuda_f32(clean_out, "clean_out", 4)
        if clean_out.shape != src.shape or clean_out.data_ptr() in (src.data_ptr(), out.data_ptr()):
            raise ValueError("clean_out must match src's shape and alias neither src nor out")
    pidx = _index_tensor(partner_idx, torch.int64, src.device, "partner_idx", 0, n_bank - 1)
    rk = _index_tensor(ratio_k, torch.int32, src.device, "ratio_k", 0, int(ratio_den))
    if pidx.numel() != bsz or rk.numel() != bsz:
        raise ValueError("partner_idx and ratio_k must have one entry per image")
    lib = _lib.load()
    if src.dtype == torch.uint8:
        with torch.cuda.device(src.device):
            _lib.check(lib.ramdsir_mix_dual_u8(src.data_ptr(), bank.data_ptr(), pidx.data_ptr(), rk.data_ptr(),
                                               out.data_ptr(), None if clean_out is None else clean_out.data_ptr(),
                                               bsz, c, h, w, b, n_bank, int(ratio_den),
                                               float(clip[0]), float(clip[1]), float(scale), float(shift),
                                               _stream_ptr(stream), _lib.FLAG_INDEPENDENT if independent else 0))
            cur = torch.cuda.current_stream() if stream is None else stream
            pidx.record_stream(cur)
            rk.record_stream(cur)
        return out
    with torch.cuda.device(src.device):
        ws, ws_ptr, ws_n = _workspace(workspace_bytes(bsz, c, h, w, b), src.device, workspace)
        _lib.check(lib.ramdsir_mix_dual_f32(src.data_ptr(), bank.data_ptr(), pidx.data_ptr(), rk.data_ptr(),
                                          out.data_ptr(), None if clean_out is None else clean_out.data_ptr(),
                                          bsz, c, h, w, b, n_bank, int(ratio_den),
                                          float(clip[0]), float(clip[1]), float(scale), float(shift),
                                          ws_ptr, ws_n, _stream_ptr(stream),
                                          _lib.FLAG_INDEPENDENT if independent else 0))
        cur = torch.cuda.current_stream() if stream is None else stream
        for t in (pidx, rk, ws):
            if t is not None:
                t.record_stream(cur)
    return out
