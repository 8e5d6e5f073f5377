"""Compare the CPU emulation of the fused CUDA pipeline with the numpy oracle (dev tool)."""
import ctypes, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ram_oracle as O

lib = ctypes.CDLL(os.path.join(ROOT, "tools/emulate/libemul.so"))
F = ctypes.POINTER(ctypes.c_float)
SIG = [F, F, ctypes.POINTER(ctypes.c_longlong), ctypes.POINTER(ctypes.c_int), ctypes.c_int, F,
       ctypes.c_int] * 1 + [ctypes.c_int] * 4 + [ctypes.c_float] * 4 + [F]
lib.emul_fused.argtypes = SIG
lib.emul_fused_static.argtypes = SIG
lib.emul_multipass.argtypes = SIG

def ptr(a, t=F): return a.ctypes.data_as(t)

def smooth(rs, c, h, w, lo, hi, integer):
    z = rs.standard_normal((c, h, w))
    fy = np.fft.fftfreq(h)[:, None]; fx = np.fft.fftfreq(w)[None, :]
    y = np.real(np.fft.ifft2(np.fft.fft2(z) / (1 + (np.hypot(fy, fx) / 0.04) ** 2)))
    y = (y - y.min()) / (y.max() - y.min()); y = lo + (hi - lo) * y
    return (np.round(y) if integer else y).astype(np.float32)

def run(C, H, W, L, epi, rng, B=2, N=3, integer=True, static=False, const_plane=False):
    fn = {False: lib.emul_fused, True: lib.emul_fused_static, "mp": lib.emul_multipass}[static]
    smem = {False: lib.emul_fused_smem_bytes, True: lib.emul_fused_static_smem_bytes, "mp": lambda *a: 0}[static]
    lo, hi = (0, 255) if integer else (-1, 1)
    rs = np.random.RandomState(7)
    src = np.stack([smooth(rs, C, H, W, lo, hi, integer) for _ in range(B)])
    par = np.stack([smooth(rs, C, H, W, lo, hi, integer) for _ in range(N)])
    b = O.band_halfwidth(H, W, L)
    if const_plane:
        src[0, 0] = 77.0
    print({False: "runtime", True: "static", "mp": "multipass"}[static], "C,H,W,b", C, H, W, b, "smem", smem(H, W, b))
    bank = np.zeros((N, C, 2*b+1, 2*b+1), np.float32)
    dummy = np.zeros(1, np.float32)
    rc = fn(ptr(par), ptr(dummy), None, None, 10, ptr(dummy), N, C, H, W, b, 0, 0, 0, 0, ptr(bank))
    assert rc == 0, rc
    want_bank = O.build_bank(par, L)
    print("  bank rel_l2 %.3e" % O.rel_l2(bank, want_bank))
    pidx = np.array([rng.randint(0, N-1) for _ in range(B)], np.int64)
    rk = np.array([rng.randint(1, 9) for _ in range(B)], np.int32)
    out = np.zeros_like(src)
    rc = fn(ptr(src), ptr(bank), ptr(pidx, ctypes.POINTER(ctypes.c_longlong)), ptr(rk, ctypes.POINTER(ctypes.c_int)), 10,
                        ptr(out), B, C, H, W, b, *[float(e) for e in epi], None)
    assert rc == 0
    want = O.mix_batch(src, par, pidx, rk, L, epi=epi)
    print("  mix  rel_l2 %.3e max_abs %.3e  (k=%s)" % (O.rel_l2(out, want), O.max_abs(out, want), rk))
    return O.rel_l2(out, want), O.max_abs(out, want)

import random
rng = random.Random(3)
run(3, 256, 256, 0.1, O.FUNDUS_EPILOGUE, rng)
run(1, 384, 384, 0.1, O.PROSTATE_EPILOGUE, rng, integer=False)
run(2, 64, 128, 0.1, O.FUNDUS_EPILOGUE, rng)
run(1, 128, 64, 0.05, O.PROSTATE_EPILOGUE, rng, integer=False)
run(1, 512, 256, 0.1, O.FUNDUS_EPILOGUE, rng)

run(3, 256, 256, 0.1, O.FUNDUS_EPILOGUE, rng, static=True)
run(1, 384, 384, 0.1, O.PROSTATE_EPILOGUE, rng, integer=False, static=True)
run(3, 256, 256, 0.1, O.FUNDUS_EPILOGUE, rng, static=True, const_plane=True)
run(3, 256, 256, 0.1, O.FUNDUS_EPILOGUE, rng, static=False, const_plane=True)
run(1, 384, 384, 0.1, O.PROSTATE_EPILOGUE, rng, integer=False, static=True, const_plane=True)


run(2, 512, 512, 0.1, O.FUNDUS_EPILOGUE, rng, static="mp", B=1, N=2)
run(1, 1024, 1024, 0.1, O.FUNDUS_EPILOGUE, rng, static="mp", B=1, N=1)
run(1, 512, 512, 0.1, O.PROSTATE_EPILOGUE, rng, static="mp", B=1, N=1, integer=False, const_plane=True)

