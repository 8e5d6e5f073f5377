#!/usr/bin/env python
"""Mint golden vectors from the UNMODIFIED reference (test infrastructure).

Run in the build container only (it needs /root/reference, which does not exist on the
GPU box):

    python oracle/gen_golden.py            # rewrites tests/golden/*.npz

Every fixture stores the exact inputs next to the reference's outputs, so the tests
never have to re-create inputs.  `ratio` is drawn by the reference itself through
`random.randint(1, 10)` (fundus.py:35); we seed Python's `random` before each call and
record which numerator came out (peeked by replaying the same seed), so the value the
reference used is known without patching any reference code.

Fixtures
  ram_small.npz      odd / non-square / tiny shapes, several L, all through fundus.py
  ram_fundus256.npz  one 3x256x256 uint8-valued sample (config 1/2 shape), f64 path
  ram_prostate384.npz one 1x384x384 and one 3x384x384 [-1,1]-valued sample via prostate.py
  ram_kats.npz       known-answer cases: ratio==1, partner==source, all-zero source, b==0
"""
import os
import random
import sys

import numpy as np

REF = "/root/reference/code"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")


def _load_reference():
    if not os.path.isdir(REF):
        raise SystemExit("reference not mounted at %s; golden vectors can only be minted "
                         "in the build container" % REF)
    sys.path.insert(0, REF)
    from dataset import fundus, prostate  # noqa: E402  (unmodified reference modules)
    return fundus, prostate


def _smooth(rs, c, h, w, lo, hi, integer):
    """Image-like field: white noise low-passed in the Fourier domain (so amplitude
    spectra are not flat and the band blend is actually exercised)."""
    z = rs.standard_normal((c, h, w))
    fy = np.fft.fftfreq(h)[:, None]
    fx = np.fft.fftfreq(w)[None, :]
    filt = 1.0 / (1.0 + (np.hypot(fy, fx) / 0.04) ** 2)
    y = np.real(np.fft.ifft2(np.fft.fft2(z) * filt))
    y = (y - y.min()) / (y.max() - y.min())
    y = lo + (hi - lo) * y
    if integer:
        y = np.round(y)
    return y


def _ref_mix(mod, src_chw, trg_chw, L, seed, f64=True):
    """Call the reference exactly as its dataset does (fundus.py:213-214):
    amp = extract_amp_spectrum(partner CHW); out = source_to_target_freq(src HWC, amp, L).
    Returns (out_hwc, ratio_k, amp)."""
    dt = np.float64 if f64 else np.float32
    src_hwc = np.ascontiguousarray(src_chw.transpose(1, 2, 0)).astype(dt)
    trg = trg_chw.astype(dt)
    random.seed(seed)
    k = random.randint(1, 10)          # peek: first draw after this seed
    random.seed(seed)                  # replay for the reference
    amp = mod.extract_amp_spectrum(trg)
    out = mod.source_to_target_freq(src_hwc, amp, L=L)
    return np.ascontiguousarray(out), k, amp


def gen_dataset_goldens(fundus, prostate):
    """Run the reference's OWN dataset classes (Fundus_Multi fundus.py:128-242, Prostate_Multi
    prostate.py:121-212) on the synthetic fixtures of tests/fixture_data.py and record, per sample, the
    returned tensors (sub-sampled to keep the fixture small) and the next draw of both RNG streams after
    the call (proves equal RNG consumption)."""
    import tempfile
    import torch
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
    import fixture_data as FX

    def after_state():
        return np.array([random.random(), np.random.rand()])

    with tempfile.TemporaryDirectory() as tmp:
        froot = FX.make_fundus(os.path.join(tmp, "fundus"))
        out = {}
        for tag, kw in (("in", dict(is_out_domain=False)), ("out", dict(is_out_domain=True))):
            ds = fundus.Fundus_Multi(domain_idx_list=[0, 1, 2], base_dir=froot, split="train", transform=FX.ResizeJitter(),
                                     test_domain_idx=3, **kw)
            for i in (0, 4, 9):
                np.random.seed(100 + i)
                random.seed(100 + i)
                torch.manual_seed(0)
                img, img_freq, mask = ds[i]
                key = "fundus_%s_%d" % (tag, i)
                out[key + "_img"] = img.numpy()[:, ::8, ::8]
                out[key + "_freq"] = img_freq.numpy()[:, ::4, ::4]
                out[key + "_mask"] = mask.numpy()[:, ::8, ::8]
                out[key + "_sums"] = np.array([img.double().sum().item(), img_freq.double().sum().item(), mask.double().sum().item()])
                out[key + "_after"] = after_state()
        proot = FX.make_prostate(os.path.join(tmp, "prostate"))
        real_listdir = os.listdir
        os.listdir = lambda p: sorted(real_listdir(p))      # prostate.py:134,183 depend on directory order
        try:
            ds = prostate.Prostate_Multi(domain_idx_list=[0, 1, 2, 3, 4], base_dir=proot, split="train", test_domain_idx=5,
                                         is_out_domain=True)
            for i in (0, 7, 13):
                np.random.seed(200 + i)
                random.seed(200 + i)
                img, img_freq, mask = ds[i]
                key = "prostate_%d" % i
                out[key + "_img"] = img.numpy()[:, ::8, ::8]
                out[key + "_freq"] = img_freq.numpy()[:, ::4, ::4]
                out[key + "_mask"] = mask.numpy()[::8, ::8]
                out[key + "_sums"] = np.array([img.double().sum().item(), img_freq.double().sum().item(), mask.double().sum().item()])
                out[key + "_after"] = after_state()
                out[key + "_dtypes"] = np.array([str(img.dtype), str(img_freq.dtype), str(mask.dtype)])
        finally:
            os.listdir = real_listdir
        np.savez_compressed(os.path.join(OUT, "ram_datasets.npz"), **out)


def main():
    fundus, prostate = _load_reference()
    os.makedirs(OUT, exist_ok=True)
    rs = np.random.RandomState(1337)      # reference default seed (train.py:60)

    # ---- small shapes -------------------------------------------------------------
    small = {}
    cases = [  # (name, C, H, W, L, value range, integer?)
        ("sq32", 3, 32, 32, 0.1, (0, 255), True),
        ("rect48x40", 1, 48, 40, 0.1, (-1, 1), False),
        ("odd31x45", 2, 31, 45, 0.1, (0, 255), True),
        ("odd33x33", 3, 33, 33, 0.2, (0, 255), True),
        ("sq64", 3, 64, 64, 0.1, (0, 255), True),
        ("sq64_L25", 1, 64, 64, 0.25, (-1, 1), False),
        ("rect96x64", 3, 96, 64, 0.1, (0, 255), True),
        ("sq128", 1, 128, 128, 0.05, (-1, 1), False),
    ]
    names = []
    for idx, (name, c, h, w, L, (lo, hi), integer) in enumerate(cases):
        for rep in range(2):
            key = "%s_%d" % (name, rep)
            src = _smooth(rs, c, h, w, lo, hi, integer).astype(np.float32)
            trg = _smooth(rs, c, h, w, lo, hi, integer).astype(np.float32)
            out, k, amp = _ref_mix(fundus, src, trg, L, seed=1000 + 10 * idx + rep)
            small[key + "_src"] = src
            small[key + "_trg"] = trg
            small[key + "_out"] = out
            if rep == 0 and name in ("sq32", "odd31x45", "rect48x40"):
                small[key + "_amp"] = amp       # extract_amp_spectrum output (a1), few cases
            small[key + "_meta"] = np.array([L, k], dtype=np.float64)
            names.append(key)
    small["names"] = np.array(names)
    np.savez_compressed(os.path.join(OUT, "ram_small.npz"), **small)

    # ---- fundus 3x256x256, uint8-valued (fundus.py:209-214) -----------------------
    src = _smooth(rs, 3, 256, 256, 0, 255, True).astype(np.uint8)
    trg = _smooth(rs, 3, 256, 256, 0, 255, True).astype(np.uint8)
    out64, k, _ = _ref_mix(fundus, src.astype(np.float32), trg.astype(np.float32), 0.1, seed=2024)
    out32, k32, _ = _ref_mix(fundus, src.astype(np.float32), trg.astype(np.float32), 0.1,
                             seed=2024, f64=False)
    assert k == k32
    np.savez_compressed(os.path.join(OUT, "ram_fundus256.npz"), src=src, trg=trg,
                        out_f64=out64, out_native_f32=out32.astype(np.float32),
                        meta=np.array([0.1, k]))

    # ---- prostate 1x384x384 and 3x384x384, [-1,1] (prostate.py:184-188) -----------
    pro = {}
    for tag, c, seed in (("c1", 1, 31), ("c3", 3, 32)):
        # inputs are float16-representable so the fixture stores them losslessly in 2 B/px
        src = _smooth(rs, c, 384, 384, -1, 1, False).astype(np.float16)
        trg = _smooth(rs, c, 384, 384, -1, 1, False).astype(np.float16)
        out, k, _ = _ref_mix(prostate, src.astype(np.float32), trg.astype(np.float32), 0.1, seed=seed)
        pro[tag + "_src"] = src
        pro[tag + "_trg"] = trg
        pro[tag + "_out"] = out if c == 1 else out.astype(np.float32)  # c3 kept f32 for size
        pro[tag + "_meta"] = np.array([0.1, k])
    np.savez_compressed(os.path.join(OUT, "ram_prostate384.npz"), **pro)

    # ---- known-answer cases ---------------------------------
    kat = {}
    src = _smooth(rs, 3, 64, 64, 0, 255, True).astype(np.float32)
    trg = _smooth(rs, 3, 64, 64, 0, 255, True).astype(np.float32)
    # ratio == 1.0 -> identity: find a seed whose first randint(1,10) is 10
    seed10 = next(s for s in range(10000) if (random.seed(s), random.randint(1, 10))[1] == 10)
    out, k, _ = _ref_mix(fundus, src, trg, 0.1, seed=seed10)
    assert k == 10
    kat["ratio1_src"], kat["ratio1_trg"], kat["ratio1_out"] = src, trg, out
    # partner == source -> identity for any ratio
    out, k, _ = _ref_mix(fundus, src, src, 0.1, seed=5)
    kat["same_src"], kat["same_out"], kat["same_k"] = src, out, np.array(k)
    # all-zero source: zero-amplitude bins take phase 0 (np.angle(0) == 0)
    zero = np.zeros_like(src)
    out, k, _ = _ref_mix(fundus, zero, trg, 0.1, seed=6)
    kat["zero_trg"], kat["zero_out"], kat["zero_k"] = trg, out, np.array(k)
    # b == 0: only DC changes (L so small that floor(min(h,w)*L) == 0)
    out, k, _ = _ref_mix(fundus, src, trg, 0.01, seed=7)
    kat["dc_src"], kat["dc_trg"], kat["dc_out"], kat["dc_k"] = src, trg, out, np.array(k)
    np.savez_compressed(os.path.join(OUT, "ram_kats.npz"), **kat)

    gen_dataset_goldens(fundus, prostate)

    for f in sorted(os.listdir(OUT)):
        print("%-24s %8.1f KiB" % (f, os.path.getsize(os.path.join(OUT, f)) / 1024))


if __name__ == "__main__":
    main()
