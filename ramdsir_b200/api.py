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
import os

import numpy as np
import torch

from . import _lib

FUNDUS_EPILOGUE = (0.0, 255.0, 1.0 / 127.5, -1.0)   # fundus.py:215,222-223
PROSTATE_EPILOGUE = (-1.0, 1.0, 1.0, 0.0)            # prostate.py:188
NO_EPILOGUE = (-math.inf, math.inf, 1.0, 0.0)

_DEBUG_CHECKS = os.environ.get("RAMDSIR_DEBUG_CHECKS", "0") == "1"

__all__ = ["band_halfwidth", "build_amp_bank", "ram_mix_batch", "kernel_path", "workspace_bytes", "prepare",
           "FUNDUS_EPILOGUE", "PROSTATE_EPILOGUE", "NO_EPILOGUE"]


def band_halfwidth(h: int, w: int, L: float = 0.1) -> int:
    """b = floor(min(h, w) * L) in float64 (fundus.py:26)."""
    return int(_lib.load().ramdsir_band_halfwidth(int(h), int(w), float(L)))


def kernel_path(c: int, h: int, w: int, b: int) -> int:
    return int(_lib.load().ramdsir_path(int(c), int(h), int(w), int(b)))


def workspace_bytes(n: int, c: int, h: int, w: int, b: int) -> int:
    return int(_lib.load().ramdsir_workspace_bytes(int(n), int(c), int(h), int(w), int(b)))


def _stream_of(stream):
    return torch.cuda.current_stream() if stream is None else stream


class _on_stream:
    """Make `stream` torch's current stream for the whole call, so that every upload, dtype conversion and
    allocation the call performs is ordered on (and owned by) the stream the library launches on."""

    def __init__(self, device, stream):
        self.dev = torch.cuda.device(device)
        self.ctx = None if stream is None else torch.cuda.stream(stream)

    def __enter__(self):
        self.dev.__enter__()
        if self.ctx is not None:
            self.ctx.__enter__()
        return torch.cuda.current_stream()

    def __exit__(self, *exc):
        if self.ctx is not None:
            self.ctx.__exit__(*exc)
        self.dev.__exit__(*exc)


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


def prepare(c: int, h: int, w: int, L: float = 0.1) -> None:
    """Build the per-shape twiddle tables now (ramdsir_prepare) instead of inside the first RAM call."""
    _lib.check(_lib.load().ramdsir_prepare(int(c), int(h), int(w), band_halfwidth(h, w, L)))


def _need_cuda_src(t, name):
    """Sources may be float32 or uint8 (what the reference's loaders start from, fundus.py:205-214)."""
    if isinstance(t, torch.Tensor) and t.is_cuda and t.dtype == torch.uint8:
        if t.dim() != 4:
            raise ValueError("%s must have 4 dims, got shape %s" % (name, tuple(t.shape)))
        if not t.is_contiguous():
            raise ValueError("%s must be contiguous (NCHW)" % name)
        return t
    return _need_cuda_f32(t, name, 4)


def _overlap(a, b):
    """Do the memory ranges of two dense tensors intersect?"""
    a0, b0 = a.data_ptr(), b.data_ptr()
    return a0 < b0 + b.numel() * b.element_size() and b0 < a0 + a.numel() * a.element_size()


def _workspace(nbytes, device, workspace):
    if nbytes == 0:
        return None, 0, 0
    if workspace is None or workspace.numel() * workspace.element_size() < nbytes:
        workspace = torch.empty(nbytes, dtype=torch.uint8, device=device)
    return workspace, workspace.data_ptr(), workspace.numel() * workspace.element_size()


def _path_flags(path):
    if path is None:
        return 0
    try:
        return {"generic": _lib.FLAG_PATH_GENERIC, "runtime": _lib.FLAG_RUNTIME_SHAPE}[path]
    except KeyError:
        raise ValueError("path must be None, 'generic' or 'runtime'") from None


def build_amp_bank(imgs: torch.Tensor, L: float = 0.1, out: torch.Tensor | None = None, stream=None,
                   workspace: torch.Tensor | None = None, path: str | None = None) -> torch.Tensor:
    """imgs [N, C, H, W] float32 (or uint8) CUDA -> bank [N, C, 2b+1, 2b+1] float32 (K3, amplitude-bank builder).

    bank[n, c] == np.fft.fftshift(extract_amp_spectrum(imgs[n]))[c, cH-b:cH+b+1, cW-b:cW+b+1].
    `out` may be a slice of a larger (all-gather) buffer so the collective needs no staging copy.
    `path` ('generic' / 'runtime') selects another kernel family for cross-checks (RAMDSIR_FLAG_PATH_*).
    """
    imgs = _need_cuda_src(imgs, "imgs")
    n, c, h, w = imgs.shape
    b = band_halfwidth(h, w, L)
    nb = 2 * b + 1
    flags = _path_flags(path)
    if out is not None:
        _need_cuda_f32(out, "out", 4)
        if tuple(out.shape) != (n, c, nb, nb):
            raise ValueError("out must have shape %s, got %s" % ((n, c, nb, nb), tuple(out.shape)))
    lib = _lib.load()
    with _on_stream(imgs.device, stream) as cur:
        if imgs.dtype == torch.uint8 and (flags or not u8_supported(c, h, w, b)):
            imgs = imgs.float()      # device-side conversion; only the static fused kernels read uint8 directly
        if out is None:
            out = torch.empty((n, c, nb, nb), dtype=torch.float32, device=imgs.device)
        if imgs.dtype == torch.uint8:
            _lib.check(lib.ramdsir_bank_build_u8(imgs.data_ptr(), out.data_ptr(), n, c, h, w, b, int(cur.cuda_stream)))
            return out
        ws_bytes = int(lib.ramdsir_workspace_bytes_ex(n, c, h, w, b, flags))
        ws, ws_ptr, ws_n = _workspace(ws_bytes, imgs.device, workspace)
        _lib.check(lib.ramdsir_bank_build_f32_ex(imgs.data_ptr(), out.data_ptr(), n, c, h, w, b, ws_ptr, ws_n,
                                                 int(cur.cuda_stream), flags))
        if ws is not None:
            ws.record_stream(cur)
    return out


def _index_tensor(x, dtype, device, name, n, lo, hi):
    """One entry per image, validated on the host when the values live there, then uploaded (on the current stream).
    Device-resident values are only shape-checked (a range check would force a sync; set RAMDSIR_DEBUG_CHECKS=1 to
    pay for it): the kernel itself poisons the image of an out-of-range partner index with NaN."""
    if isinstance(x, torch.Tensor) and x.is_cuda:
        if x.dim() != 1 or x.numel() != n:
            raise ValueError("%s must be 1-D with one entry per image (%d), got shape %s" % (name, n, tuple(x.shape)))
        if x.device != device:
            raise ValueError("%s lives on %s, the batch on %s" % (name, x.device, device))
        if _DEBUG_CHECKS and n and (int(x.min()) < lo or int(x.max()) > hi):
            raise ValueError("%s values must lie in [%d, %d]" % (name, lo, hi))
        return x.to(dtype=dtype).contiguous()
    arr = np.asarray(x.cpu() if isinstance(x, torch.Tensor) else x)
    if arr.ndim != 1 or arr.size != n:
        raise ValueError("%s must be 1-D with one entry per image (%d)" % (name, n))
    if arr.size and (arr.min() < lo or arr.max() > hi):
        raise ValueError("%s values must lie in [%d, %d], got [%d, %d]" % (name, lo, hi, arr.min(), arr.max()))
    npdt = np.int64 if dtype == torch.int64 else np.int32
    return torch.from_numpy(np.ascontiguousarray(arr.astype(npdt))).to(device, non_blocking=False)


def ram_mix_batch(src: torch.Tensor, bank: torch.Tensor, partner_idx, ratio_k, L: float = 0.1,
                  clip=(-math.inf, math.inf), scale: float = 1.0, shift: float = 0.0,
                  out: torch.Tensor | None = None, ratio_den: int = 10, stream=None,
                  workspace: torch.Tensor | None = None, independent: bool = False, serial: bool = False,
                  clean_out: torch.Tensor | None = None, path: str | None = None) -> torch.Tensor:
    """Random Amplitude Mixup of a batch (K1/K2).

    src [B, C, H, W] float32 CUDA, or uint8 (read directly by the kernel on the static shapes, bit-identical to
    passing src.float(); converted on the device otherwise); bank [N, C, 2b+1, 2b+1] from `build_amp_bank`;
    partner_idx [B] int (indices into the bank); ratio_k [B] int, ratio = ratio_k / ratio_den is the
    weight of the SOURCE amplitude (fundus.py:35,37; the reference draws ratio_k = randint(1, 10)).
    Returns clamp(RAM(src), clip) * scale + shift, float32 [B, C, H, W].

    clean_out (optional, same shape as src): also receives the normalised clean batch src * scale + shift
    (fundus.py:217-218), fused into the same kernel on the static shapes (ramdsir_mix_dual_f32).

    Everything the call enqueues (index upload, dtype conversion, the RAM kernel) runs on `stream` (default: torch's
    current stream).  Consecutive calls on one stream overlap their tails automatically when their buffers are
    disjoint (see RAMDSIR_FLAG_* in include/ramdsir.h); independent=True skips that check, serial=True forbids the
    overlap.  path ('generic' / 'runtime') selects another kernel family for cross-checks.
    """
    src = _need_cuda_src(src, "src")
    bank = _need_cuda_f32(bank, "bank", 4)
    bsz, c, h, w = src.shape
    b = band_halfwidth(h, w, L)
    nb = 2 * b + 1
    if tuple(bank.shape[1:]) != (c, nb, nb):
        raise ValueError("bank must be [N, %d, %d, %d] for src %s and L=%g, got %s"
                         % (c, nb, nb, tuple(src.shape), L, tuple(bank.shape)))
    if bank.device != src.device:
        raise ValueError("src and bank must live on the same device")
    n_bank = bank.shape[0]
    if out is not None:
        _need_cuda_f32(out, "out", 4)
        if out.shape != src.shape or _overlap(out, src):
            raise ValueError("out must match src's shape and must not overlap it")
    if clean_out is not None:
        _need_cuda_f32(clean_out, "clean_out", 4)
        if clean_out.shape != src.shape or _overlap(clean_out, src) or (out is not None and _overlap(clean_out, out)):
            raise ValueError("clean_out must match src's shape and overlap neither src nor out")
    flags = _path_flags(path) | (_lib.FLAG_INDEPENDENT if independent else 0) | (_lib.FLAG_SERIAL if serial else 0)
    lib = _lib.load()
    with _on_stream(src.device, stream) as cur:
        keep = []                    # temporaries the kernel reads: owned by `cur` because they were made under it
        if src.dtype == torch.uint8 and (_path_flags(path) or not u8_supported(c, h, w, b)):
            src = src.float()
            keep.append(src)
        if out is None:
            out = torch.empty(src.shape, dtype=torch.float32, device=src.device)
        pidx = _index_tensor(partner_idx, torch.int64, src.device, "partner_idx", bsz, 0, n_bank - 1)
        rk = _index_tensor(ratio_k, torch.int32, src.device, "ratio_k", bsz, 0, int(ratio_den))
        keep += [pidx, rk]
        clean_ptr = None if clean_out is None else clean_out.data_ptr()
        if src.dtype == torch.uint8:
            _lib.check(lib.ramdsir_mix_dual_u8(src.data_ptr(), bank.data_ptr(), pidx.data_ptr(), rk.data_ptr(),
                                               out.data_ptr(), clean_ptr, bsz, c, h, w, b, n_bank, int(ratio_den),
                                               float(clip[0]), float(clip[1]), float(scale), float(shift),
                                               int(cur.cuda_stream), flags))
        else:
            ws_bytes = int(lib.ramdsir_workspace_bytes_ex(bsz, c, h, w, b, flags))
            ws, ws_ptr, ws_n = _workspace(ws_bytes, src.device, workspace)
            keep.append(ws)
            _lib.check(lib.ramdsir_mix_dual_f32(src.data_ptr(), bank.data_ptr(), pidx.data_ptr(), rk.data_ptr(),
                                              out.data_ptr(), clean_ptr, bsz, c, h, w, b, n_bank, int(ratio_den),
                                              float(clip[0]), float(clip[1]), float(scale), float(shift),
                                              ws_ptr, ws_n, int(cur.cuda_stream), flags))
        for t in keep:
            if t is not None:
                t.record_stream(cur)
    return out


class MixCall:
    """A validated, argument-bound `ram_mix_batch`: everything that call does before its launch -- shape / dtype /
    overlap checks, index upload, workspace -- happens once, here; `call()` afterwards is one C-ABI call
    (ramdsir_mix_dual_f32 / _u8) with the bound pointers.  For loops that re-launch on the same buffers (a ring of
    prefetched batches): the host side of a launch drops from tens of microseconds of Python to a few, which is what
    keeps the automatic tail overlap of consecutive 256-image launches (88 us each) fed when the host is busy.

    The stream is fixed at construction (`stream`, default: torch's current stream then).  The object keeps every tensor
    it was given alive; the caller keeps the OBJECT alive until the launches it enqueued have completed, and must not
    free or reshape the tensors underneath it.  Results are those of ram_mix_batch with the same arguments, bit for bit.
    """

    def __init__(self, src, bank, partner_idx, ratio_k, out, L=0.1, ratio_den=10, clip=(-math.inf, math.inf), scale=1.0, shift=0.0,
                 stream=None, workspace=None, independent=False, serial=False, clean_out=None):
        src = _need_cuda_src(src, "src")
        bank = _need_cuda_f32(bank, "bank", 4)
        bsz, c, h, w = src.shape
        b = band_halfwidth(h, w, L)
        nb = 2 * b + 1
        if tuple(bank.shape[1:]) != (c, nb, nb) or bank.device != src.device:
            raise ValueError("bank must be [N, %d, %d, %d] on %s, got %s on %s"
                             % (c, nb, nb, src.device, tuple(bank.shape), bank.device))
        _need_cuda_f32(out, "out", 4)
        if out.shape != src.shape or _overlap(out, src):
            raise ValueError("out must match src's shape and must not overlap it")
        if clean_out is not None:
            _need_cuda_f32(clean_out, "clean_out", 4)
            if clean_out.shape != src.shape or _overlap(clean_out, src) or _overlap(clean_out, out):
                raise ValueError("clean_out must match src's shape and overlap neither src nor out")
        if src.dtype == torch.uint8 and not u8_supported(c, h, w, b):
            raise ValueError("uint8 sources of shape %s are not read directly; pass src.float()" % (tuple(src.shape),))
        flags = (_lib.FLAG_INDEPENDENT if independent else 0) | (_lib.FLAG_SERIAL if serial else 0)
        lib = _lib.load()
        lo, hi = float(clip[0]), float(clip[1])
        with _on_stream(src.device, stream) as cur:
            pidx = _index_tensor(partner_idx, torch.int64, src.device, "partner_idx", bsz, 0, bank.shape[0] - 1)
            rk = _index_tensor(ratio_k, torch.int32, src.device, "ratio_k", bsz, 0, int(ratio_den))
            clean_ptr = None if clean_out is None else clean_out.data_ptr()
            head = (src.data_ptr(), bank.data_ptr(), pidx.data_ptr(), rk.data_ptr(), out.data_ptr(), clean_ptr,
                    bsz, c, h, w, b, bank.shape[0], int(ratio_den), lo, hi, float(scale), float(shift))
            if src.dtype == torch.uint8:
                ws = None
                self._fn, self._args = lib.ramdsir_mix_dual_u8, head + (int(cur.cuda_stream), flags)
            else:
                ws, ws_ptr, ws_n = _workspace(int(lib.ramdsir_workspace_bytes_ex(bsz, c, h, w, b, flags)), src.device, workspace)
                self._fn, self._args = lib.ramdsir_mix_dual_f32, head + (ws_ptr, ws_n, int(cur.cuda_stream), flags)
            self.stream = cur
        self.out = out
        self._keep = (src, bank, pidx, rk, out, clean_out, ws)

    def call(self):
        """Enqueue the launch on the stream chosen at construction; returns `out`."""
        rc = self._fn(*self._args)
        if rc:
            _lib.check(rc)
        return self.out

    __call__ = call
