"""Host side of the GPU geometric transforms: the coefficient / index tables libramdsir builds for Pillow-exact
resizing (ramdsir_resize_tables_host), applied with numpy, reproduce PIL.Image.resize byte for byte.  No GPU needed."""
import ctypes

import numpy as np
import pytest
from PIL import Image


def _tables(n_in, n_out, flt):
    from ramdsir_b200 import _lib
    lib = _lib.load()
    ks = ctypes.c_int()
    _lib.check(lib.ramdsir_resize_tables_host(n_in, n_out, flt, None, None, ctypes.byref(ks)))
    if flt == _lib.FILTER_NEAREST:
        t = np.empty(n_out, np.int32)
        _lib.check(lib.ramdsir_resize_tables_host(n_in, n_out, flt, t.ctypes.data, None, ctypes.byref(ks)))
        return t
    bounds, kk = np.empty((n_out, 2), np.int32), np.empty((n_out, ks.value), np.int32)
    _lib.check(lib.ramdsir_resize_tables_host(n_in, n_out, flt, bounds.ctypes.data, kk.ctypes.data, ctypes.byref(ks)))
    return bounds, kk


def _apply_1d(a, bounds, kk, axis):
    """Pillow's 8-bit pass: 22-bit fixed-point accumulation from 2^21, shift, clip."""
    a = np.moveaxis(a.astype(np.int64), axis, 0)
    out = np.empty((bounds.shape[0],) + a.shape[1:], np.uint8)
    for i, (lo, n) in enumerate(bounds):
        ss = np.full(a.shape[1:], 1 << 21, np.int64)
        for j in range(n):
            ss += a[lo + j] * int(kk[i, j])
        out[i] = np.clip(ss >> 22, 0, 255)
    return np.moveaxis(out, 0, axis)


@pytest.mark.parametrize("shape", [(300, 280, 256, 256), (800, 801, 256, 256), (97, 131, 200, 150), (256, 256, 384, 300),
                                   (64, 200, 64, 77), (123, 45, 17, 45), (5, 7, 40, 33)])
def test_host_tables_reproduce_pillow(shape):
    from ramdsir_b200 import _lib
    hin, win, hout, wout = shape
    rs = np.random.RandomState(hin + win)
    img = rs.randint(0, 256, (hin, win, 3)).astype(np.uint8)
    want = np.array(Image.fromarray(img).resize((wout, hout), Image.BILINEAR))
    cur = img
    if win != wout:
        cur = _apply_1d(cur, *_tables(win, wout, _lib.FILTER_BILINEAR), axis=1)      # horizontal first, 8-bit intermediate
    if hin != hout:
        cur = _apply_1d(cur, *_tables(hin, hout, _lib.FILTER_BILINEAR), axis=0)
    np.testing.assert_array_equal(cur, want)
    mask = rs.randint(0, 256, (hin, win)).astype(np.uint8)
    want_m = np.array(Image.fromarray(mask).resize((wout, hout), Image.NEAREST))
    xt, yt = _tables(win, wout, _lib.FILTER_NEAREST), _tables(hin, hout, _lib.FILTER_NEAREST)
    assert xt.min() >= 0 and yt.min() >= 0
    np.testing.assert_array_equal(mask[yt][:, xt], want_m)
