"""Dataset-level drop-ins (ramdsir_b200.datasets) against goldens recorded from the reference's OWN
Fundus_Multi / Prostate_Multi on the synthetic fixtures of tests/fixture_data.py
(oracle/gen_golden.py::gen_dataset_goldens)."""
import os
import random

import numpy as np
import pytest
import torch

import fixture_data as FX

pytestmark = pytest.mark.gpu
REL_L2_TOL, MAX_ABS_TOL = 1e-5, 1e-4


def _rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300)


@pytest.fixture(scope="module")
def golden(golden_dir):
    return np.load(os.path.join(golden_dir, "ram_datasets.npz"))


@pytest.fixture(scope="module")
def fundus_root(tmp_path_factory):
    return FX.make_fundus(str(tmp_path_factory.mktemp("fundus")))


@pytest.fixture(scope="module")
def prostate_root(tmp_path_factory):
    return FX.make_prostate(str(tmp_path_factory.mktemp("prostate")))


@pytest.mark.parametrize("tag,out_domain", [("in", False), ("out", True)])
def test_fundus_multi_matches_reference(golden, fundus_root, tag, out_domain):
    from ramdsir_b200.datasets import Fundus_Multi
    ds = Fundus_Multi(domain_idx_list=[0, 1, 2], base_dir=fundus_root, split="train", transform=FX.ResizeJitter(),
                      test_domain_idx=3, is_out_domain=out_domain)
    assert len(ds) == 10
    for i in (0, 4, 9):
        np.random.seed(100 + i)
        random.seed(100 + i)
        img, img_freq, mask = ds[i]
        after = np.array([random.random(), np.random.rand()])
        key = "fundus_%s_%d" % (tag, i)
        assert img.shape == (3, 256, 256) and img_freq.shape == (3, 256, 256) and mask.shape == (2, 256, 256)
        assert img.dtype == torch.float32 and img_freq.dtype == torch.float32 and mask.dtype == torch.float32
        # same RNG consumption (transform draws, 2 x np.random.choice, 1 x random.randint) -> same next draws
        np.testing.assert_array_equal(after, golden[key + "_after"])
        # clean image and mask are integer/byte work: bit exact
        np.testing.assert_array_equal(img.numpy()[:, ::8, ::8], golden[key + "_img"])
        np.testing.assert_array_equal(mask.numpy()[:, ::8, ::8], golden[key + "_mask"])
        assert img.double().sum().item() == golden[key + "_sums"][0]
        assert mask.double().sum().item() == golden[key + "_sums"][2]
        # RAM output: float tolerance on normalised intensities (partner + ratio selection are exact, else this fails)
        got, want = img_freq.numpy()[:, ::4, ::4], golden[key + "_freq"]
        assert _rel(got, want) <= REL_L2_TOL and np.abs(got - want).max() <= MAX_ABS_TOL
        assert abs(img_freq.double().sum().item() - golden[key + "_sums"][1]) <= 1e-4 * img_freq.numel() ** 0.5 + 1e-2


def test_prostate_multi_matches_reference(golden, prostate_root, monkeypatch):
    from ramdsir_b200.datasets import Prostate_Multi
    real = os.listdir
    monkeypatch.setattr(os, "listdir", lambda p: sorted(real(p)))   # same directory order as the golden run
    ds = Prostate_Multi(domain_idx_list=[0, 1, 2, 3, 4], base_dir=prostate_root, split="train", test_domain_idx=5,
                        is_out_domain=True)
    assert len(ds) == 15
    for i in (0, 7, 13):
        np.random.seed(200 + i)
        random.seed(200 + i)
        img, img_freq, mask = ds[i]
        after = np.array([random.random(), np.random.rand()])
        key = "prostate_%d" % i
        assert [str(img.dtype), str(img_freq.dtype), str(mask.dtype)] == list(golden[key + "_dtypes"])
        assert img.shape == (3, 384, 384) and img_freq.shape == (3, 384, 384) and mask.shape == (384, 384)
        np.testing.assert_array_equal(after, golden[key + "_after"])
        np.testing.assert_array_equal(img.numpy()[:, ::8, ::8], golden[key + "_img"])
        np.testing.assert_array_equal(mask.numpy()[::8, ::8], golden[key + "_mask"])
        got, want = img_freq.numpy()[:, ::4, ::4], golden[key + "_freq"]
        assert _rel(got, want) <= REL_L2_TOL and np.abs(got - want).max() <= MAX_ABS_TOL
        assert float(img_freq.min()) >= -1.0 and float(img_freq.max()) <= 1.0      # prostate.py:188


def test_deferred_mode_with_forked_workers_equals_sample_mode(fundus_root):
    """ram="deferred": host-only __getitem__ in forked DataLoader workers + one batched RAM launch in the main
    process (RamBatchHook) gives the same batch as the per-sample mode."""
    from torch.utils.data import DataLoader
    from ramdsir_b200.datasets import Fundus_Multi, RamBatchHook

    def seed_worker(worker_id):                                   # train.py:80-83
        s = torch.initial_seed() % 2 ** 32
        np.random.seed(s)
        random.seed(s)

    kw = dict(domain_idx_list=[0, 1, 2], base_dir=fundus_root, split="train", transform=FX.ResizeJitter(), test_domain_idx=3)
    hook = RamBatchHook("fundus")
    g = torch.Generator().manual_seed(7)
    loader = DataLoader(Fundus_Multi(ram="deferred", **kw), batch_size=5, shuffle=False, num_workers=2,
                        worker_init_fn=seed_worker, generator=g, multiprocessing_context="fork")
    batches = [hook(b) for b in loader]
    assert len(batches) == 2 and batches[0][1].shape == (5, 3, 256, 256) and batches[0][1].is_cuda
    # replay the same per-worker seeds through the per-sample dataset (num_workers=0 semantics)
    g = torch.Generator().manual_seed(7)
    base_seed = int(torch.empty((), dtype=torch.int64).random_(generator=g).item())
    ds = Fundus_Multi(ram="sample", **kw)
    for wid, batch in enumerate(batches):                          # worker w produced batch w (2 workers, 2 batches)
        s = (base_seed + wid) % 2 ** 32
        np.random.seed(s)
        random.seed(s)
        ref = [ds[i] for i in range(5 * wid, 5 * wid + 5)]
        img = torch.stack([r[0] for r in ref])
        freq = torch.stack([r[1] for r in ref])
        mask = torch.stack([r[2] for r in ref])
        assert torch.allclose(batch[0].cpu(), img, atol=1e-6)
        assert torch.equal(batch[1].cpu(), freq)                   # same kernels, same inputs: bit-identical
        assert torch.equal(batch[2].cpu(), mask)
