"""Dataset-level drop-ins: the RAM call sites of the reference's loaders, with RAM on the GPU.

    Fundus_Multi(domain_idx_list, base_dir, split, num, transform, is_freq, is_out_domain, test_domain_idx)
        reference: code/dataset/fundus.py:128-242   -> __getitem__ returns (img, img_freq, mask)
    Prostate_Multi(...same constructor...)
        reference: code/dataset/prostate.py:121-212 -> __getitem__ returns (img, img_freq, mask.long())

Same constructor arguments, same files read, same RNG consumption in the same order
(transform draws from `random`; np.random.choice(domain) ; np.random.choice(partner) ; random.randint(1, 10)
-- fundus.py:194-195,205,208,35 / prostate.py:182-183), same tuple layout, dtypes and value ranges.

Two modes (constructor keyword `ram=`):
  "sample"   (default) RAM runs per sample inside __getitem__ through the CUDA library -> a strict drop-in
             for `train.py --ram`; CUDA cannot be used in forked workers, so use num_workers=0.
  "deferred" __getitem__ only does the host work and returns (img_raw, partner_chw, ratio_k, mask);
             `RamBatchHook` then runs ONE `ram_mix_batch` per collated batch on the device, right where the
             reference calls `.cuda()` (train.py:244,411).  Works with forked DataLoader workers.
"""
from __future__ import annotations

import os
import random

import numpy as np
import torch
from PIL import Image
from torch.utils.data import Dataset

from .api import FUNDUS_EPILOGUE, PROSTATE_EPILOGUE, build_amp_bank, ram_mix_batch

__all__ = ["Fundus_Multi", "Prostate_Multi", "RamBatchHook", "encode_fundus_mask", "encode_fundus_mask_batch"]


def encode_fundus_mask(mask_img) -> torch.Tensor:
    """8-bit fundus mask -> 2-channel multilabel float tensor (fundus.py:227-239 + transform.py:10-14).

    Grey levels: > 200 background, 51..200 optic disc rim, <= 50 cup.  The reference maps them to classes
    0 / 1 / 2 and then to [0,0] / [0,1] / [1,1]; i.e. channel 0 = cup, channel 1 = disc (cup included).
    """
    m = np.array(mask_img).astype(np.uint8)
    out = np.zeros((2,) + m.shape, dtype=np.float32)
    out[0] = m <= 50
    out[1] = m <= 200
    return torch.from_numpy(out)


def encode_fundus_mask_batch(mask_u8: torch.Tensor, stream=None) -> torch.Tensor:
    """Device-side mask encoding (ramdsir_encode_fundus_mask_u8): uint8 CUDA [B, H, W] -> float32 [B, 2, H, W], the
    same two comparisons as `encode_fundus_mask` (fundus.py:227-239 + transform.py:10-14), bit-exact."""
    from . import _lib
    if not (isinstance(mask_u8, torch.Tensor) and mask_u8.is_cuda and mask_u8.dtype == torch.uint8 and mask_u8.dim() == 3):
        raise TypeError("mask_u8 must be a CUDA uint8 tensor [B, H, W]")
    mask_u8 = mask_u8.contiguous()
    bsz, h, w = mask_u8.shape
    out = torch.empty((bsz, 2, h, w), dtype=torch.float32, device=mask_u8.device)
    with torch.cuda.device(mask_u8.device):
        s = torch.cuda.current_stream() if stream is None else stream
        _lib.check(_lib.load().ramdsir_encode_fundus_mask_u8(mask_u8.data_ptr(), out.data_ptr(), bsz, h, w, int(s.cuda_stream)))
    return out


def _gpu_ram_one(src_chw_f32: np.ndarray, partner_chw_f32: np.ndarray, ratio_k: int, epilogue, L=0.1) -> torch.Tensor:
    dev = torch.device("cuda", torch.cuda.current_device())
    bank = build_amp_bank(torch.from_numpy(np.ascontiguousarray(partner_chw_f32))[None].to(dev), L=L)
    lo, hi, scale, shift = epilogue
    y = ram_mix_batch(torch.from_numpy(np.ascontiguousarray(src_chw_f32))[None].to(dev), bank, [0], [ratio_k], L=L,
                      clip=(lo, hi), scale=scale, shift=shift)
    return y[0].cpu()


class _MultiBase(Dataset):
    domain_name: list

    def _init_common(self, domain_idx_list, base_dir, split, num, transform, is_freq, is_out_domain, test_domain_idx, ram):
        if ram not in ("sample", "deferred"):
            raise ValueError("ram must be 'sample' or 'deferred'")
        self.transform = transform
        self.base_dir = base_dir
        self.num = num
        self.domain_idx_list = domain_idx_list
        self.split = split
        self.is_freq = is_freq
        self.is_out_domain = is_out_domain
        self.test_domain_idx = test_domain_idx
        self.ram = ram

    def __len__(self):
        return len(self.id_path)

    def _partner_domains(self, cur_domain_name):
        names = list(self.domain_name)
        names.remove(self.domain_name[self.test_domain_idx])      # fundus.py:161-162 / prostate.py:146-147
        if self.is_out_domain:
            names.remove(cur_domain_name)                         # fundus.py:202-203
        return names


class Fundus_Multi(_MultiBase):
    domain_name = ["Domain1", "Domain2", "Domain3", "Domain4"]

    def __init__(self, domain_idx_list=None, base_dir=None, split="train", num=None, transform=None, is_freq=True,
                 is_out_domain=False, test_domain_idx=None, ram="sample"):
        self._init_common(domain_idx_list, base_dir, split, num, transform, is_freq, is_out_domain, test_domain_idx, ram)
        lines = []
        for d in self.domain_idx_list:                            # fundus.py:141-151
            with open(os.path.join(self.base_dir + "/{}_{}.list".format(self.domain_name[d], split))) as f:
                lines += f.readlines()
        self.id_path = [ln.replace("\n", "") for ln in lines]
        if self.num is not None:
            self.id_path = self.id_path[: self.num]
        print("total {} samples".format(len(self.id_path)))

    def __getitem__(self, index):
        rec = self.id_path[index]
        img_rel, mask_rel = rec.split(" ")[0], rec.split(" ")[1]
        img = Image.open(os.path.join(self.base_dir, img_rel))
        cur_domain = img_rel.split("/")[0]
        mask = Image.open(os.path.join(self.base_dir, mask_rel)).convert("L")
        if self.split == "test":                                  # fundus.py:167-189 (no RAM on this branch)
            full_mask = encode_fundus_mask(mask)
            sample = {"img": img, "mask": mask}
            if self.transform:
                sample = self.transform(sample)
            return sample["img"], sample["mask"], full_mask, rec
        sample = {"img": img, "mask": mask}
        if self.transform:
            sample = self.transform(sample)                      # consumes `random` draws ahead of RAM's
        if not self.is_freq:
            return sample["img"], sample["mask"]
        img, mask = sample["img"], sample["mask"]
        # --- partner selection: the reference's two numpy draws, in order (fundus.py:205-208) ---
        other_domain = np.random.choice(self._partner_domains(cur_domain), 1)[0]
        with open(os.path.join(self.base_dir, other_domain, "train.list")) as f:
            other_lines = f.readlines()
        other_id = np.random.choice(other_lines).replace("\n", "").split(" ")[0]
        other = Image.open(os.path.join(self.base_dir, other_domain, other_id)).resize((256, 256), Image.BILINEAR)
        other_u8, src_u8 = np.array(other), np.array(img)
        other = other_u8.astype(np.float32).transpose(2, 0, 1)
        src = src_u8.astype(np.float32).transpose(2, 0, 1)
        ratio_k = random.randint(1, 10)                          # fundus.py:35 (drawn inside low_freq_mutate_np)
        mask_t = encode_fundus_mask(mask)
        if self.ram == "deferred":
            if src_u8.dtype == np.uint8 and other_u8.dtype == np.uint8:
                # 8-bit images stay 8-bit until the RAM kernel reads them: a quarter of the bytes through the
                # worker queue and over PCIe; the kernel's float(src) is the reference's astype(float32)
                return (torch.from_numpy(np.ascontiguousarray(src_u8.transpose(2, 0, 1))),
                        torch.from_numpy(np.ascontiguousarray(other_u8.transpose(2, 0, 1))), ratio_k, mask_t)
            return torch.from_numpy(src), torch.from_numpy(other), ratio_k, mask_t
        img_freq = _gpu_ram_one(src, other, ratio_k, FUNDUS_EPILOGUE)           # fundus.py:213-215,222-225
        img_t = torch.from_numpy(_normalise_u8(src))                            # fundus.py:217-220
        return img_t.float(), img_freq.float(), mask_t


def _normalise_u8(src_chw):
    """img /= 127.5; img -= 1.0 in float32, in that order (fundus.py:217-218)."""
    a = src_chw.copy()
    a /= 127.5
    a -= 1.0
    return a


class Prostate_Multi(_MultiBase):
    domain_name = ["Domain1", "Domain2", "Domain3", "Domain4", "Domain5", "Domain6"]

    def __init__(self, domain_idx_list=None, base_dir=None, split="train", num=None, transform=None, is_freq=True,
                 is_out_domain=False, test_domain_idx=None, ram="sample"):
        self._init_common(domain_idx_list, base_dir, split, num, transform, is_freq, is_out_domain, test_domain_idx, ram)
        self.id_path = []
        for d in self.domain_idx_list:                            # prostate.py:133-136
            files = os.listdir(os.path.join(self.base_dir, self.domain_name[d], "image"))
            self.id_path += [self.domain_name[d] + "/image/" + f for f in files]
        if self.num is not None:
            self.id_path = self.id_path[: self.num]
        print("total {} samples".format(len(self.id_path)))

    def __getitem__(self, index):
        rec = self.id_path[index]
        img = np.load(os.path.join(self.base_dir, rec))           # H x W x 3
        cur_domain = rec.split("/")[0]
        mask = np.load(os.path.join(self.base_dir, rec.replace("image", "mask")))
        if self.split == "test":                                  # prostate.py:152-169
            return torch.from_numpy(img.transpose(2, 0, 1)).float(), torch.from_numpy(mask).long(), rec.split("/")[-1]
        if not self.is_freq:
            return torch.from_numpy(img.transpose(2, 0, 1)).float(), torch.from_numpy(mask).long()
        other_domain = np.random.choice(self._partner_domains(cur_domain), 1)[0]            # prostate.py:182
        other_id = np.random.choice(os.listdir(os.path.join(self.base_dir, other_domain, "image")))   # prostate.py:183
        other = np.load(os.path.join(self.base_dir, other_domain, "image", other_id))
        src = np.ascontiguousarray(img.transpose(2, 0, 1), dtype=np.float32)
        other = np.ascontiguousarray(other.transpose(2, 0, 1), dtype=np.float32)
        ratio_k = random.randint(1, 10)                          # prostate.py:35 via low_freq_mutate_np
        mask_t = torch.from_numpy(mask).long()
        if self.ram == "deferred":
            return torch.from_numpy(src), torch.from_numpy(other), ratio_k, mask_t
        img_freq = _gpu_ram_one(src, other, ratio_k, PROSTATE_EPILOGUE)        # prostate.py:186-188
        return torch.from_numpy(src).float(), img_freq.float(), mask_t


class RamBatchHook:
    """Post-collate hook for `ram="deferred"` datasets: one bank build + one RAM launch per batch.

    batch = (img_raw [B,C,H,W], partner [B,C,H,W], ratio_k [B], mask) as produced by the default collate.
    Returns (img, img_freq, mask) on the device with the reference's normalisation
    (fundus: /127.5 - 1 on both; prostate: img untouched, img_freq clipped to [-1, 1]).
    """

    def __init__(self, dataset: str, device=None, L: float = 0.1):
        if dataset not in ("fundus", "prostate"):
            raise ValueError("dataset must be 'fundus' or 'prostate'")
        self.dataset = dataset
        self.L = L
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)

    def __call__(self, batch):
        img_raw, partner, ratio_k, mask = batch
        img_raw = img_raw.to(self.device, non_blocking=True).contiguous()
        partner = partner.to(self.device, non_blocking=True).contiguous()
        if img_raw.dtype != torch.uint8:
            img_raw = img_raw.float()
        if partner.dtype != torch.uint8:
            partner = partner.float()
        bank = build_amp_bank(partner, L=self.L)                 # row i is sample i's partner
        idx = torch.arange(img_raw.shape[0], device=self.device)
        lo, hi, scale, shift = FUNDUS_EPILOGUE if self.dataset == "fundus" else PROSTATE_EPILOGUE
        # fundus: the kernel also emits the normalised clean batch from the same read of img_raw (fused dual output)
        img = torch.empty(img_raw.shape, dtype=torch.float32, device=self.device) if self.dataset == "fundus" else None
        img_freq = ram_mix_batch(img_raw, bank, idx, torch.as_tensor(ratio_k, dtype=torch.int32), L=self.L,
                                 clip=(lo, hi), scale=scale, shift=shift, clean_out=img)
        if img is None:
            img = img_raw.float()                                 # prostate.py:190-193: the clean image is untouched
        return img, img_freq, mask.to(self.device, non_blocking=True)
