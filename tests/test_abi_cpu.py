"""CPU-side checks of the C-ABI boundary: the library builds/loads without a GPU and exports
every symbol include/ramdsir.h declares.  No compute calls here."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from ramdsir_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        _lib.build()
    return _lib.load()


def _declared_functions():
    text = open(os.path.join(ROOT, "include", "ramdsir.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(ramdsir_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_expected_entry_points():
    names = _declared_functions()
    for must in ("ramdsir_bank_build_f32", "ramdsir_mix_f32", "ramdsir_mix_host_f32", "ramdsir_band_halfwidth",
                 "ramdsir_workspace_bytes", "ramdsir_last_error", "ramdsir_version"):
        assert must in names


def test_library_exports_every_declared_symbol(lib):
    from ramdsir_b200 import _lib
    declared = _declared_functions()
    assert sorted(_lib.SYMBOLS) == declared          # binding list and header agree
    raw = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared:
        assert hasattr(raw, name), "libramdsir.so does not export %s" % name


def test_band_halfwidth_matches_reference_expression(lib):
    import numpy as np
    for h, w, L in ((256, 256, 0.1), (384, 384, 0.1), (1024, 1024, 0.1), (31, 45, 0.1), (33, 33, 0.2),
                    (800, 800, 0.1), (512, 512, 0.1), (64, 64, 0.01), (64, 64, 0.25)):
        want = int(np.floor(np.amin((h, w)) * L))    # fundus.py:26
        assert lib.ramdsir_band_halfwidth(h, w, L) == want


def test_path_selection_and_workspace_without_gpu(lib):
    from ramdsir_b200 import _lib
    assert lib.ramdsir_path(3, 256, 256, 25) == _lib.PATH_FUSED
    assert lib.ramdsir_path(1, 384, 384, 38) == _lib.PATH_FUSED
    assert lib.ramdsir_path(2, 31, 45, 3) == _lib.PATH_GENERIC
    assert lib.ramdsir_path(3, 1024, 1024, 102) == _lib.PATH_MULTIPASS
    assert lib.ramdsir_path(1, 512, 512, 51) == _lib.PATH_MULTIPASS
    assert lib.ramdsir_workspace_bytes(64, 3, 1024, 1024, 102) > 0
    assert lib.ramdsir_path(3, 64, 64, 40) == _lib.PATH_NONE          # band exceeds the spectrum
    assert lib.ramdsir_workspace_bytes(16, 3, 256, 256, 25) == 0
    assert lib.ramdsir_workspace_bytes(2, 2, 31, 45, 3) == 2 * 2 * (31 * 4 + 7 * 4) * 8
    assert lib.ramdsir_version() >= 100


def test_invalid_arguments_return_codes_not_crashes(lib):
    from ramdsir_b200 import _lib
    rc = lib.ramdsir_mix_f32(None, None, None, None, None, 4, 3, 256, 256, 25, 1, 10, 0.0, 1.0, 1.0, 0.0, None, 0, None)
    assert rc == _lib.ERR_INVALID
    assert b"null" in lib.ramdsir_last_error()
    rc = lib.ramdsir_mix_f32(8, 16, 16, 16, 32, 4, 3, 256, 256, 25, 1, 10, 0.0, 1.0, 1.0, 0.0, None, 0, None)
    assert rc == _lib.ERR_INVALID and b"aligned" in lib.ramdsir_last_error()      # float4 kernels need 16-byte alignment
    rc = lib.ramdsir_bank_build_f32(None, None, 4, 3, 64, 64, 40, None, 0, None)
    assert rc == _lib.ERR_UNSUPPORTED


def test_uint8_entry_points_validate_without_gpu(lib):
    """ramdsir_u8_supported is pure host logic; the uint8 entry points refuse bad arguments before touching CUDA."""
    from ramdsir_b200 import _lib
    assert lib.ramdsir_u8_supported(3, 256, 256, 25) == 1
    assert lib.ramdsir_u8_supported(1, 384, 384, 38) == 1
    assert lib.ramdsir_u8_supported(3, 256, 256, 20) == 0        # runtime-shape fused kernel: float32 only
    assert lib.ramdsir_u8_supported(3, 1024, 1024, 102) == 0     # multipass
    assert lib.ramdsir_u8_supported(2, 31, 45, 3) == 0           # generic
    args = (4, 3, 256, 256, 25, 1, 10, 0.0, 255.0, 1.0, 0.0, None, 0)
    assert lib.ramdsir_mix_dual_u8(None, None, None, None, None, None, *args) == _lib.ERR_INVALID
    assert b"null" in lib.ramdsir_last_error()
    assert lib.ramdsir_mix_dual_u8(6, 16, 16, 16, 32, None, *args) == _lib.ERR_INVALID      # src not 4-byte aligned
    assert b"aligned" in lib.ramdsir_last_error()
    assert lib.ramdsir_mix_dual_u8(8, 16, 16, 16, 32, 32, *args) == _lib.ERR_INVALID        # clean_out aliases out
    rc = lib.ramdsir_mix_dual_u8(8, 16, 16, 16, 32, None, 4, 3, 128, 128, 12, 1, 10, 0.0, 255.0, 1.0, 0.0, None, 0)
    assert rc == _lib.ERR_UNSUPPORTED and b"uint8" in lib.ramdsir_last_error()
    assert lib.ramdsir_bank_build_u8(8, 16, 2, 3, 128, 128, 12, None) == _lib.ERR_UNSUPPORTED
    assert lib.ramdsir_bank_build_u8(None, None, 0, 3, 256, 256, 25, None) == _lib.OK       # empty batch


def test_uint8_to_float_trick_is_exact():
    """quad_of_bytes (ram_fused.cu): float(0x4B000000 | byte) - 2^23 == byte for every byte value."""
    import numpy as np
    b = np.arange(256, dtype=np.uint32)
    f = (np.uint32(0x4B000000) | b).view(np.float32) - np.float32(8388608.0)
    assert f.dtype == np.float32 and np.array_equal(f, b.astype(np.float32))


def test_product_path_refuses_cpu_tensors():
    import torch
    import ramdsir_b200 as R
    with pytest.raises(TypeError):
        R.build_amp_bank(torch.zeros(1, 3, 64, 64))
    with pytest.raises(TypeError):
        R.ram_mix_batch(torch.zeros(1, 3, 64, 64), torch.zeros(1, 3, 13, 13), [0], [5])


def test_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "ramdsir_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in text and "from oracle" not in text, f
                assert "ram_oracle" not in text, f


def test_bound_launch_refuses_host_tensors():
    """MixCall is the GPU path only: host tensors are refused at binding (no CPU fallback anywhere in the package)."""
    import torch
    import ramdsir_b200 as R
    x = torch.zeros(1, 1, 64, 64)
    with pytest.raises(TypeError):
        R.MixCall(x, torch.zeros(1, 1, 13, 13), [0], [1], torch.empty_like(x))
