"""CPU checks of the dataset drop-ins' HOST logic against goldens recorded from the reference's own
dataset classes: file handling, transform hand-off, mask encoding, normalisation and -- most importantly
-- RNG consumption order (fundus.py:194-208,35; prostate.py:182-183).  `ram="deferred"` needs no GPU."""
import os
import random

import numpy as np
import pytest

import fixture_data as FX


@pytest.fixture(scope="module")
def golden(golden_dir):
    return np.load(os.path.join(golden_dir, "ram_datasets.npz"))


@pytest.mark.parametrize("tag,out_domain", [("in", False), ("out", True)])
def test_fundus_host_side_matches_reference(golden, tmp_path, tag, out_domain):
    from ramdsir_b200.datasets import Fundus_Multi
    root = FX.make_fundus(str(tmp_path))
    ds = Fundus_Multi(domain_idx_list=[0, 1, 2], base_dir=root, split="train", transform=FX.ResizeJitter(),
                      test_domain_idx=3, is_out_domain=out_domain, ram="deferred")
    for i in (0, 4, 9):
        np.random.seed(100 + i)
        random.seed(100 + i)
        img_raw, partner, ratio_k, mask = ds[i]
        after = np.array([random.random(), np.random.rand()])
        key = "fundus_%s_%d" % (tag, i)
        np.testing.assert_array_equal(after, golden[key + "_after"])       # identical RNG consumption
        assert img_raw.dtype == partner.dtype == __import__("torch").uint8    # 8-bit images travel as bytes
        img = img_raw.numpy().astype(np.float32)
        img /= 127.5
        img -= 1.0
        np.testing.assert_array_equal(img[:, ::8, ::8], golden[key + "_img"])
        np.testing.assert_array_equal(mask.numpy()[:, ::8, ::8], golden[key + "_mask"])
        assert partner.shape == (3, 256, 256) and 1 <= ratio_k <= 10


def test_prostate_host_side_matches_reference(golden, tmp_path, monkeypatch):
    from ramdsir_b200.datasets import Prostate_Multi
    root = FX.make_prostate(str(tmp_path))
    real = os.listdir
    monkeypatch.setattr(os, "listdir", lambda p: sorted(real(p)))
    ds = Prostate_Multi(domain_idx_list=[0, 1, 2, 3, 4], base_dir=root, split="train", test_domain_idx=5,
                        is_out_domain=True, ram="deferred")
    for i in (0, 7, 13):
        np.random.seed(200 + i)
        random.seed(200 + i)
        img, partner, ratio_k, mask = ds[i]
        after = np.array([random.random(), np.random.rand()])
        key = "prostate_%d" % i
        np.testing.assert_array_equal(after, golden[key + "_after"])
        np.testing.assert_array_equal(img.numpy()[:, ::8, ::8], golden[key + "_img"])
        np.testing.assert_array_equal(mask.numpy()[::8, ::8], golden[key + "_mask"])


def test_test_split_and_plain_branches(tmp_path):
    from ramdsir_b200.datasets import Fundus_Multi
    root = FX.make_fundus(str(tmp_path))
    # the reference reads <Domain>_test.list for split == "test" (fundus.py:147-150); reuse the train lists
    for d in FX.FUNDUS_DOMAINS:
        with open(os.path.join(root, d + "_train.list")) as f, open(os.path.join(root, d + "_test.list"), "w") as g:
            g.write(f.read())
    ds = Fundus_Multi(domain_idx_list=[3], base_dir=root, split="test", transform=FX.ResizeJitter(), test_domain_idx=3)
    img, mask, full_mask, rec = ds[0]
    assert full_mask.shape[0] == 2 and rec.startswith("Domain4/") and img.size == (256, 256)
    plain = Fundus_Multi(domain_idx_list=[0], base_dir=root, split="train", transform=FX.ResizeJitter(), test_domain_idx=3,
                         is_freq=False)
    assert len(plain[0]) == 2
