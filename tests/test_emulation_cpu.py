"""CPU coverage of the CUDA kernels' index algebra.

tools/emulate/emul_fused.cpp compiles the very same __host__ __device__ phase bodies the kernels call
(ram_fused_core/static.cuh, ram_multipass.cuh) for the host and runs each CTA pipeline task by task with a
heap buffer standing in for shared memory.  Here every pipeline variant is checked against the numpy oracle, so
a container without a GPU still validates twiddle tables, band indexing, packing and the butterfly stages.
(The emulator is a test tool: it is never imported by the package and is not a fallback.)"""
import ctypes
import os
import subprocess

import numpy as np
import pytest

from oracle import ram_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tools", "emulate", "emul_fused.cpp")
LIB = os.path.join(ROOT, "tools", "emulate", "libemul.so")
CSRC = os.path.join(ROOT, "ramdsir_b200", "csrc")


@pytest.fixture(scope="module")
def emu():
    deps = [SRC] + [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    if not os.path.exists(LIB) or os.path.getmtime(LIB) < max(os.path.getmtime(d) for d in deps):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-I", CSRC, SRC, "-o", LIB])
    lib = ctypes.CDLL(LIB)
    F = ctypes.POINTER(ctypes.c_float)
    sig = [F, F, ctypes.POINTER(ctypes.c_longlong), ctypes.POINTER(ctypes.c_int), ctypes.c_int, F] + \
          [ctypes.c_int] * 5 + [ctypes.c_float] * 4 + [F]
    for name in ("emul_fused", "emul_fused_static", "emul_multipass"):
        getattr(lib, name).argtypes = sig
    return lib


def _field(rs, shape, lo, hi, integer):
    n, c, h, w = shape
    z = rs.standard_normal(shape)
    fy = np.fft.fftfreq(h)[:, None]
    fx = np.fft.fftfreq(w)[None, :]
    y = np.real(np.fft.ifft2(np.fft.fft2(z) / (1.0 + (np.hypot(fy, fx) / 0.04) ** 2)))
    y = (y - y.min()) / (y.max() - y.min())
    y = lo + (hi - lo) * y
    return (np.round(y) if integer else y).astype(np.float32)


def _ptr(a, t=ctypes.POINTER(ctypes.c_float)):
    return a.ctypes.data_as(t)


CASES = [  # (entry point, C, H, W, integer-valued, constant plane?, bank tolerance)
    ("emul_fused", 2, 64, 128, True, True),             # runtime-shape kernel (direct sums + exact degenerate rule)
    ("emul_fused", 1, 256, 256, True, False),
    ("emul_fused", 2, 128, 192, False, True),
    ("emul_fused_static", 2, 256, 256, True, True),     # default kernels, incl. a constant plane
    ("emul_fused_static", 1, 384, 384, False, True),
    ("emul_multipass", 1, 512, 512, True, False),       # K2
]


@pytest.mark.parametrize("case", CASES, ids=lambda c: "%s_%dx%dx%d" % (c[0].replace("emul_", ""), c[1], c[2], c[3]))
def test_emulated_pipeline_matches_oracle(emu, case):
    name, c, h, w, integer, const_plane = case
    fn = getattr(emu, name)
    rs = np.random.RandomState(h + 7 * c)
    lo, hi = (0, 255) if integer else (-1, 1)
    src = _field(rs, (2, c, h, w), lo, hi, integer)
    par = _field(rs, (2, c, h, w), lo, hi, integer)
    if const_plane:
        src[1, 0] = 77.0 if integer else 0.25
    b = O.band_halfwidth(h, w, 0.1)
    epi = O.FUNDUS_EPILOGUE if integer else O.PROSTATE_EPILOGUE
    bank = np.zeros((2, c, 2 * b + 1, 2 * b + 1), np.float32)
    dummy = np.zeros(1, np.float32)
    assert fn(_ptr(par), _ptr(dummy), None, None, 10, _ptr(dummy), 2, c, h, w, b, 0, 0, 0, 0, _ptr(bank)) == 0
    assert O.rel_l2(bank, O.build_bank(par, 0.1)) <= 1e-6
    pidx = np.array([1, 0], np.int64)
    rk = np.array([3, 8], np.int32)
    out = np.zeros_like(src)
    assert fn(_ptr(src), _ptr(bank), _ptr(pidx, ctypes.POINTER(ctypes.c_longlong)), _ptr(rk, ctypes.POINTER(ctypes.c_int)),
              10, _ptr(out), 2, c, h, w, b, *[float(e) for e in epi], None) == 0
    want = O.mix_batch(src, par, pidx, rk, 0.1, epi=epi)
    assert O.rel_l2(out, want) <= 1e-5 and O.max_abs(out, want) <= 1e-4
