"""BASELINE config 5b: one `train.py --ram --rec` step with RAM on the GPU in the loader path
(tools/e2e_step.py restates /root/reference/code/train.py:223-296 around a small consumer network).

Both arms start from the same weights and see the same samples and RNG draws (forked DataLoader workers seeded by
seed_worker, train.py:80-83); the only difference is where img_freq comes from: the CUDA RAM kernel behind
RamBatchHook, or the numpy oracle per sample.  Loss and gradient norm must agree."""
import os
import sys

import numpy as np
import pytest
import torch

import fixture_data as FX
from oracle import ram_oracle as O

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))


def _oracle_ram(src_chw, partner_chw, ratio_k):
    return O.mix_batch(src_chw[None], partner_chw[None], [0], [ratio_k], 0.1, epi=O.FUNDUS_EPILOGUE)[0]


@pytest.mark.parametrize("workers", [0, 2], ids=["inline", "forked-workers"])
def test_train_step_loss_parity_gpu_ram_vs_cpu_ram(tmp_path_factory, workers):
    import e2e_step as E
    root = FX.make_fundus(str(tmp_path_factory.mktemp("fundus_e2e")), counts=E.E2E_FIXTURE)
    res = E.compare(root, FX.ResizeJitter(), _oracle_ram, steps=3, num_workers=workers)
    assert len(res["gpu"]) == 3 and len(res["cpu"]) == 3 and res["batch"] == 16          # B = 3 + 6 + 7 (train.py:35)
    for g in res["gpu"]:
        assert np.isfinite(g["loss"]) and g["grad_norm"] > 0
    # step 0 starts from identical weights: only the 1e-7-level difference of img_freq separates the arms
    assert res["diff"][0]["loss_rel"] <= 1e-5, res["diff"]
    assert res["diff"][0]["grad_norm_rel"] <= 1e-5, res["diff"]
    # later steps inherit the optimizer's amplification of that difference; they must still track closely
    for d in res["diff"][1:]:
        assert d["loss_rel"] <= 1e-3 and d["grad_norm_rel"] <= 1e-2, res["diff"]
    # the loss actually depends on img_freq (the RAM branch is consumed): its seg term differs from the clean one
    assert abs(res["gpu"][0]["seg_freq"] - res["gpu"][0]["seg_img"]) > 1e-6
