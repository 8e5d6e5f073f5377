"""ctypes binding of libramdsir.so (the C ABI declared in include/ramdsir.h).

There is deliberately no fallback here: if the CUDA library is missing or a call fails,
the caller gets an exception.  The numpy oracle under oracle/ is test infrastructure and is
never imported from this package.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("RAMDSIR_LIB") or os.path.join(_HERE, "libramdsir.so")   # RAMDSIR_LIB: dev builds of the same ABI
CSRC = os.path.join(_HERE, "csrc")
SOURCES = ["ramdsir_abi.cu", "ram_fused.cu", "ram_generic.cu", "ram_multipass.cu", "ram_select.cu", "ram_transform.cu"]

FILTER_NEAREST, FILTER_BILINEAR = 0, 2
FLAG_INDEPENDENT, FLAG_SERIAL, FLAG_PATH_GENERIC, FLAG_RUNTIME_SHAPE = 1, 2, 4, 8
OK, ERR_INVALID, ERR_UNSUPPORTED, ERR_CUDA, ERR_WORKSPACE = 0, 1, 2, 3, 4
PATH_NONE, PATH_FUSED, PATH_MULTIPASS, PATH_GENERIC = 0, 1, 2, 3

# every symbol include/ramdsir.h declares (tests/test_abi_cpu.py checks the .so exports them all)
SYMBOLS = [
    "ramdsir_version", "ramdsir_last_error", "ramdsir_band_halfwidth", "ramdsir_path", "ramdsir_prepare",
    "ramdsir_workspace_bytes", "ramdsir_workspace_bytes_ex", "ramdsir_bank_build_f32", "ramdsir_bank_build_f32_ex",
    "ramdsir_mix_f32", "ramdsir_mix_f32_ex", "ramdsir_mix_dual_f32",
    "ramdsir_u8_supported", "ramdsir_bank_build_u8", "ramdsir_mix_dual_u8",
    "ramdsir_amp_spectrum_workspace_bytes", "ramdsir_amp_spectrum_f32", "ramdsir_band_blend_f32",
    "ramdsir_pipe_create", "ramdsir_pipe_destroy", "ramdsir_mix_host_f32", "ramdsir_mix_host_u8", "ramdsir_launch_count",
    "ramdsir_rng_state_words", "ramdsir_rng_seed_host", "ramdsir_rng_create", "ramdsir_rng_destroy", "ramdsir_rng_draw",
    "ramdsir_rng_get_state", "ramdsir_encode_fundus_mask_u8",
    "ramdsir_resize_workspace_bytes", "ramdsir_resize_u8", "ramdsir_crop_pad_chw_u8", "ramdsir_resize_tables_host",
]


class RamdsirError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("ramdsir error %d: %s" % (code, msg))
        self.code = code


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile libramdsir.so in-tree for sm_100a with nvcc (cross-compiles without a GPU)."""
    srcs = [os.path.join(CSRC, s) for s in SOURCES]
    deps = srcs + [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".h", ".cuh"))]
    deps.append(os.path.join(_HERE, "..", "include", "ramdsir.h"))
    if not force and os.path.exists(LIB_PATH):
        if os.path.getmtime(LIB_PATH) >= max(os.path.getmtime(d) for d in deps):
            return LIB_PATH
    cmd = ["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
           "-t", "4", "-cudart", "shared", "-shared", "-Xcompiler", "-fPIC", "-Xlinker", "-rpath=/usr/local/cuda/lib64",
           "-o", LIB_PATH] + srcs
    res = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or res.returncode != 0:
        print(res.stdout + res.stderr)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed building libramdsir.so")
    return LIB_PATH


_lock = threading.Lock()
_lib = None


def _declare(lib):
    c = ctypes
    f32p, i64p, i32p, vp = c.POINTER(c.c_float), c.POINTER(c.c_int64), c.POINTER(c.c_int32), c.c_void_p
    lib.ramdsir_version.restype = c.c_int
    lib.ramdsir_last_error.restype = c.c_char_p
    lib.ramdsir_band_halfwidth.argtypes = [c.c_int, c.c_int, c.c_double]
    lib.ramdsir_band_halfwidth.restype = c.c_int
    lib.ramdsir_path.argtypes = [c.c_int] * 4
    lib.ramdsir_prepare.argtypes = [c.c_int] * 4
    lib.ramdsir_workspace_bytes.argtypes = [c.c_int] * 5
    lib.ramdsir_workspace_bytes.restype = c.c_size_t
    lib.ramdsir_workspace_bytes_ex.argtypes = [c.c_int] * 5 + [c.c_uint]
    lib.ramdsir_workspace_bytes_ex.restype = c.c_size_t
    lib.ramdsir_bank_build_f32.argtypes = [vp, vp] + [c.c_int] * 5 + [vp, c.c_size_t, vp]
    lib.ramdsir_bank_build_f32_ex.argtypes = [vp, vp] + [c.c_int] * 5 + [vp, c.c_size_t, vp, c.c_uint]
    lib.ramdsir_mix_f32.argtypes = ([vp] * 5 + [c.c_int] * 7 + [c.c_float] * 4 + [vp, c.c_size_t, vp])
    lib.ramdsir_mix_f32_ex.argtypes = ([vp] * 5 + [c.c_int] * 7 + [c.c_float] * 4 + [vp, c.c_size_t, vp, c.c_uint])
    lib.ramdsir_mix_dual_f32.argtypes = ([vp] * 6 + [c.c_int] * 7 + [c.c_float] * 4 + [vp, c.c_size_t, vp, c.c_uint])
    lib.ramdsir_u8_supported.argtypes = [c.c_int] * 4
    lib.ramdsir_bank_build_u8.argtypes = [vp, vp] + [c.c_int] * 5 + [vp]
    lib.ramdsir_mix_dual_u8.argtypes = ([vp] * 6 + [c.c_int] * 7 + [c.c_float] * 4 + [vp, c.c_uint])
    lib.ramdsir_amp_spectrum_workspace_bytes.argtypes = [c.c_int] * 4
    lib.ramdsir_amp_spectrum_workspace_bytes.restype = c.c_size_t
    lib.ramdsir_amp_spectrum_f32.argtypes = [vp, vp] + [c.c_int] * 4 + [vp, c.c_size_t, vp]
    lib.ramdsir_band_blend_f32.argtypes = [vp, vp, vp] + [c.c_int] * 6 + [vp]
    lib.ramdsir_pipe_create.argtypes = [c.POINTER(vp)] + [c.c_int] * 5
    lib.ramdsir_pipe_destroy.argtypes = [vp]
    lib.ramdsir_pipe_destroy.restype = None
    lib.ramdsir_mix_host_f32.argtypes = [vp] * 6 + [c.c_int] * 3 + [c.c_float] * 4
    lib.ramdsir_mix_host_u8.argtypes = [vp] * 6 + [c.c_int] * 3 + [c.c_float] * 4
    lib.ramdsir_launch_count.restype = c.c_uint64
    u32p = c.POINTER(c.c_uint32)
    lib.ramdsir_rng_seed_host.argtypes = [c.c_uint32, vp, vp]
    lib.ramdsir_rng_create.argtypes = [c.POINTER(vp), c.c_int, vp, c.c_int, vp, vp, vp]
    lib.ramdsir_rng_destroy.argtypes = [vp]
    lib.ramdsir_rng_destroy.restype = None
    lib.ramdsir_rng_draw.argtypes = [vp, c.c_int, vp, vp, c.c_int, c.c_int, vp]
    lib.ramdsir_rng_get_state.argtypes = [vp, c.c_int, vp, vp]
    lib.ramdsir_encode_fundus_mask_u8.argtypes = [vp, vp, c.c_int, c.c_int, c.c_int, vp]
    lib.ramdsir_resize_workspace_bytes.argtypes = [c.c_int] * 5
    lib.ramdsir_resize_workspace_bytes.restype = c.c_size_t
    lib.ramdsir_resize_u8.argtypes = [vp, c.c_int, c.c_int, c.c_int, vp, c.c_int, c.c_int, c.c_int, vp, c.c_size_t, vp]
    lib.ramdsir_crop_pad_chw_u8.argtypes = [vp] + [c.c_int] * 8 + [vp, vp]
    lib.ramdsir_resize_tables_host.argtypes = [c.c_int, c.c_int, c.c_int, vp, vp, c.POINTER(c.c_int)]
    return lib


def load():
    """Load (once) and return the ctypes handle; raises if the library has not been built."""
    global _lib
    with _lock:
        if _lib is None:
            if not os.path.exists(LIB_PATH):
                raise RuntimeError(
                    "libramdsir.so is missing (%s). Build it with `python -c 'import __graft_entry__ as g; "
                    "g.build()'` or `sh ramdsir_b200/csrc/build.sh`; there is no CPU fallback." % LIB_PATH)
            import torch  # noqa: F401  loads torch's libcudart.so.12 first so both share one CUDA runtime
            _lib = _declare(ctypes.CDLL(LIB_PATH))
        return _lib


def check(rc):
    if rc != OK:
        raise RamdsirError(rc, load().ramdsir_last_error().decode("utf-8", "replace"))


def launch_count() -> int:
    return int(load().ramdsir_launch_count())
