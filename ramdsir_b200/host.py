"""Host-buffer entry point: RAM on batches that live in HOST memory (what a data-loader process holds).

Wraps `ramdsir_pipe_*` / `ramdsir_mix_host_f32` from include/ramdsir.h: the batch is cut into chunks and
host->device copy, the RAM kernel and device->host copy of consecutive chunks overlap on three streams.
This is the call the reference's loader would make where it runs `source_to_target_freq` today
(fundus.py:213-215, prostate.py:186-188), with the clip/normalise epilogue fused in.
"""
from __future__ import annotations

import ctypes

import numpy as np
import torch

from . import _lib
from .api import NO_EPILOGUE, band_halfwidth

__all__ = ["HostPipeline"]


def _host_ptr(t, name, dtype):
    if isinstance(t, np.ndarray):
        if t.dtype != dtype or not t.flags["C_CONTIGUOUS"]:
            raise TypeError("%s must be a C-contiguous %s array" % (name, np.dtype(dtype).name))
        return t.ctypes.data, t.size
    if isinstance(t, torch.Tensor) and not t.is_cuda:
        if t.numpy().dtype != dtype or not t.is_contiguous():
            raise TypeError("%s must be a contiguous %s CPU tensor" % (name, np.dtype(dtype).name))
        return t.data_ptr(), t.numel()
    raise TypeError("%s must be a numpy array or a CPU torch tensor" % name)


class HostPipeline:
    def __init__(self, chunk_images: int, C: int, H: int, W: int, L: float = 0.1, device=None):
        if not torch.cuda.is_available():
            raise RuntimeError("ramdsir_b200 needs a CUDA device; there is no CPU fallback")
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.C, self.H, self.W, self.L = int(C), int(H), int(W), float(L)
        self.b = band_halfwidth(H, W, L)
        self.chunk = int(chunk_images)
        self._lib = _lib.load()
        self._h = ctypes.c_void_p()
        with torch.cuda.device(self.device):
            _lib.check(self._lib.ramdsir_pipe_create(ctypes.byref(self._h), self.chunk, self.C, self.H, self.W, self.b))

    def mix(self, src_host, bank: torch.Tensor, partner_idx, ratio_k, out_host, epilogue=NO_EPILOGUE, ratio_den: int = 10):
        """src_host: [B, C, H, W] float32 -- or uint8 (fundus loaders hold uint8 images: a quarter of the H2D bytes,
        widened on the device); out_host: float32 [B, C, H, W].  Pin both for async copies.  Returns out_host."""
        if self._h is None:
            raise RuntimeError("pipeline is closed")
        is_u8 = (src_host.dtype == np.uint8) if isinstance(src_host, np.ndarray) else (src_host.dtype == torch.uint8)
        sp, sn = _host_ptr(src_host, "src_host", np.uint8 if is_u8 else np.float32)
        op, on = _host_ptr(out_host, "out_host", np.float32)
        per = self.C * self.H * self.W
        if sn % per or sn != on:
            raise ValueError("src_host/out_host must both hold B x %d x %d x %d float32" % (self.C, self.H, self.W))
        bsz = sn // per
        nb = 2 * self.b + 1
        if not (isinstance(bank, torch.Tensor) and bank.is_cuda and bank.dtype == torch.float32 and bank.is_contiguous()
                and tuple(bank.shape[1:]) == (self.C, nb, nb)):
            raise ValueError("bank must be a contiguous CUDA float32 tensor [N, %d, %d, %d]" % (self.C, nb, nb))
        if bank.device != self.device:
            raise ValueError("bank lives on %s, the pipeline on %s" % (bank.device, self.device))
        pidx = np.ascontiguousarray(np.asarray(partner_idx), dtype=np.int64)
        rk = np.ascontiguousarray(np.asarray(ratio_k), dtype=np.int32)
        if pidx.shape != (bsz,) or rk.shape != (bsz,):
            raise ValueError("partner_idx and ratio_k must have one entry per image")
        lo, hi, scale, shift = epilogue
        with torch.cuda.device(self.device):
            torch.cuda.current_stream().synchronize()      # the bank may still be in flight on torch's stream
            fn = self._lib.ramdsir_mix_host_u8 if is_u8 else self._lib.ramdsir_mix_host_f32
            _lib.check(fn(self._h, sp, bank.data_ptr(), pidx.ctypes.data, rk.ctypes.data, op, bsz, bank.shape[0],
                          int(ratio_den), float(lo), float(hi), float(scale), float(shift)))
        return out_host

    def close(self):
        if self._h is not None:
            self._lib.ramdsir_pipe_destroy(self._h)
            self._h = None

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
