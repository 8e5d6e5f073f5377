"""Pin the numpy oracle (oracle/ram_oracle.py) to outputs of the unmodified reference
(fixtures minted by oracle/gen_golden.py from /root/reference/code/dataset/{fundus,prostate}.py).
CPU only."""
import os

import numpy as np
import pytest

from oracle import ram_oracle as O


def _load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name))


def test_small_shapes_match_reference(golden_dir):
    g = _load(golden_dir, "ram_small.npz")
    for key in g["names"]:
        L, k = g[key + "_meta"]
        src, trg, ref = g[key + "_src"], g[key + "_trg"], g[key + "_out"]
        amp = O.amp_spectrum(trg)
        out = O.mix_sample(src.transpose(1, 2, 0), amp, L=float(L), ratio=int(k) / 10)
        assert out.shape == ref.shape
        assert O.rel_l2(out, ref) < 1e-13, key
        if key + "_amp" in g.files:
            np.testing.assert_allclose(amp, g[key + "_amp"], rtol=1e-13, atol=1e-9)


def test_closed_form_matches_reference(golden_dir):
    """y = x + Re(ifft2(dF)) on the band only -- the identity the CUDA kernels implement."""
    g = _load(golden_dir, "ram_small.npz")
    for key in g["names"]:
        L, k = g[key + "_meta"]
        src, trg, ref = g[key + "_src"], g[key + "_trg"], g[key + "_out"]
        c, h, w = src.shape
        b = O.band_halfwidth(h, w, float(L))
        patch = O.low_band_patch(O.amp_spectrum(trg), float(L))
        assert patch.shape == (c, 2 * b + 1, 2 * b + 1)
        out = O.mix_sample_closed_form(src, patch, b, int(k) / 10).transpose(1, 2, 0)
        assert O.rel_l2(out, ref) < 1e-12, key


def test_fundus256_f64_and_native(golden_dir):
    g = _load(golden_dir, "ram_fundus256.npz")
    L, k = g["meta"]
    src = g["src"].astype(np.float32)
    trg = g["trg"].astype(np.float32)
    out = O.mix_sample(src.transpose(1, 2, 0), O.amp_spectrum(trg), L=float(L), ratio=int(k) / 10)
    assert O.rel_l2(out, g["out_f64"]) < 1e-13
    nat = O.mix_sample(src.transpose(1, 2, 0), O.amp_spectrum(trg, "native"), L=float(L),
                       ratio=int(k) / 10, precision="native")
    assert nat.dtype == np.float32          # numpy >= 2 keeps complex64
    assert O.rel_l2(nat, g["out_native_f32"]) < 1e-6
    # f32 path vs f64 path of the reference itself: two orders inside the 1e-5 tolerance
    assert O.rel_l2(g["out_native_f32"], g["out_f64"]) < 1e-6


def test_prostate384(golden_dir):
    g = _load(golden_dir, "ram_prostate384.npz")
    for tag, tol in (("c1", 1e-13), ("c3", 2e-7)):   # c3 reference output stored as f32
        L, k = g[tag + "_meta"]
        src = g[tag + "_src"].astype(np.float32)
        trg = g[tag + "_trg"].astype(np.float32)
        out = O.mix_sample(src.transpose(1, 2, 0), O.amp_spectrum(trg), L=float(L), ratio=int(k) / 10)
        assert O.rel_l2(out, g[tag + "_out"]) < tol, tag


def test_known_answers(golden_dir):
    g = _load(golden_dir, "ram_kats.npz")
    src = g["ratio1_src"]
    # ratio == 1.0 -> identity (both in the reference output and in the oracle)
    assert O.max_abs(g["ratio1_out"], src.transpose(1, 2, 0)) < 1e-10
    out = O.mix_sample(src.transpose(1, 2, 0), O.amp_spectrum(g["ratio1_trg"]), ratio=1.0)
    assert O.max_abs(out, g["ratio1_out"]) < 1e-10
    # partner == source -> identity for any ratio
    assert O.max_abs(g["same_out"], g["same_src"].transpose(1, 2, 0)) < 1e-10
    # all-zero source: output is NOT zero (phase 0 bins carry (1-ratio)*A_t)
    k = int(g["zero_k"])
    out = O.mix_sample(np.zeros_like(src).transpose(1, 2, 0), O.amp_spectrum(g["zero_trg"]), ratio=k / 10)
    assert O.rel_l2(out, g["zero_out"]) < 1e-12
    if k != 10:
        assert np.abs(g["zero_out"]).max() > 1.0
    # b == 0: only DC changes, y = x + (1-ratio) * (|sum t| - |sum x|) * sign(sum x) / (H*W)
    k = int(g["dc_k"])
    s, t = g["dc_src"].astype(np.float64), g["dc_trg"].astype(np.float64)
    assert O.band_halfwidth(64, 64, 0.01) == 0
    shift = (1 - k / 10) * (np.abs(t.sum((1, 2))) - np.abs(s.sum((1, 2)))) * np.sign(s.sum((1, 2))) / (64 * 64)
    want = (s + shift[:, None, None]).transpose(1, 2, 0)
    assert O.max_abs(g["dc_out"], want) < 1e-9


def test_band_geometry_matches_reference_expressions():
    # floor(n * 0.1) in float64: 256->25, 384->38, 512->51, 800->80, 1024->102
    for n, b in ((256, 25), (384, 38), (512, 51), (800, 80), (1024, 102), (31, 3), (45, 4)):
        assert O.band_halfwidth(n, n, 0.1) == b
    assert O.band_window(256, 256, 0.1) == (103, 154, 103, 154)
    assert O.band_window(31, 45, 0.1) == (12, 19, 19, 26)


def test_patch_is_what_the_blend_reads():
    rs = np.random.RandomState(0)
    src = rs.rand(2, 31, 45)
    trg = rs.rand(2, 31, 45)
    a_s, a_t = O.amp_spectrum(src), O.amp_spectrum(trg)
    full = O.blend_low_band(a_s, a_t, 0.1, 0.3)
    # replacing everything outside the band of the partner by garbage changes nothing
    h1, h2, w1, w2 = O.band_window(31, 45, 0.1)
    junk = np.fft.fftshift(a_t, axes=(-2, -1)).copy()
    mask = np.ones_like(junk, bool)
    mask[:, h1:h2, w1:w2] = False
    junk[mask] = 1e9
    np.testing.assert_array_equal(full, O.blend_low_band(a_s, np.fft.ifftshift(junk, axes=(-2, -1)), 0.1, 0.3))


def test_epilogue_bitfaithful_to_reference_arithmetic():
    v = np.linspace(-20, 280, 1001)
    got = O.epilogue(v, *O.FUNDUS_EPILOGUE)
    want = np.clip(v, 0, 255).astype(np.float32)
    want /= 127.5
    want -= 1.0
    np.testing.assert_array_equal(got, want)
    np.testing.assert_array_equal(O.epilogue(v / 100, *O.PROSTATE_EPILOGUE),
                                  np.clip(v / 100, -1, 1).astype(np.float32))
