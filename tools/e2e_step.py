#!/usr/bin/env python
"""End-to-end `train.py --ram --rec` step with RAM on the GPU in the loader path (BASELINE config 5b).

The reference's training step (/root/reference/code/train.py:223-296) consumes, per source domain, a batch
`(img, img_freq, mask)` from `Fundus_Multi`, concatenates the domains (train.py:232-244), runs the encoder and the
segmentation decoder on `img` AND on `img_freq` (train.py:246-256), adds a consistency term (train.py:258-261) and
reconstructs `img` from the `img_freq` features through a domain-specific decoder (train.py:263-276), then
backward + Adam (train.py:285-287).  The networks are out of scope for this repo, so this
harness ships a SMALL consumer of the same shape -- an encoder that returns a feature list, a segmentation decoder,
a reconstruction decoder with domain-specific batch-norm -- and restates the step around it.  What is under test
is the DATA path: one loader per domain with forked workers and `seed_worker` (train.py:556-559) producing

    GPU arm:  Fundus_Multi(ram="deferred") -> RamBatchHook -> (img, img_freq, mask) on the device
    CPU arm:  the same samples and the same RNG draws, `img_freq` computed per sample by a caller-supplied CPU
              implementation of RAM (the tests and bench.py pass the numpy oracle; this file never imports it)

and the assertion that loss and gradient norm of the step agree between the two arms.

  python tools/e2e_step.py [--root DIR] [--steps 5]        # GPU arm only: step time with the GPU-RAM loader
"""
from __future__ import annotations

import argparse
import os
import random
import sys
import time

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F
from torch.utils.data import DataLoader

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

FUNDUS_BATCHES = [3, 6, 7]          # per-domain batch sizes for target domain 0 (train.py:35): 16 images per step
E2E_FIXTURE = {"Domain1": 6, "Domain2": 12, "Domain3": 14, "Domain4": 3}   # synthetic images per domain (2 steps per epoch)


def seed_worker(worker_id):
    """train.py:80-83."""
    s = torch.initial_seed() % 2 ** 32
    np.random.seed(s)
    random.seed(s)


# ------------------------------------------------------------------------------------------------------------------
# a small consumer with the reference's three roles (networks/unet.py:248-322 are out of scope)
class DomainBN(nn.Module):
    """Domain-specific batch-norm (the role of networks/dsbn.py:4-34): one BatchNorm2d per source domain."""

    def __init__(self, ch, n_domains):
        super().__init__()
        self.bns = nn.ModuleList([nn.BatchNorm2d(ch) for _ in range(n_domains)])

    def forward(self, x, domain):
        return self.bns[domain](x)


class Encoder(nn.Module):
    def __init__(self, cin=3, n=8):
        super().__init__()
        self.c1 = nn.Sequential(nn.Conv2d(cin, n, 3, padding=1), nn.BatchNorm2d(n), nn.ReLU(inplace=True))
        self.c2 = nn.Sequential(nn.Conv2d(n, 2 * n, 3, padding=1, stride=2), nn.BatchNorm2d(2 * n), nn.ReLU(inplace=True))

    def forward(self, x):
        f1 = self.c1(x)
        return [f1, self.c2(f1)]          # a feature list; the last entry feeds the reconstruction decoder


class SegDecoder(nn.Module):
    def __init__(self, n=8, classes=2):
        super().__init__()
        self.up = nn.ConvTranspose2d(2 * n, n, 2, stride=2)
        self.out = nn.Sequential(nn.Conv2d(2 * n, n, 3, padding=1), nn.ReLU(inplace=True), nn.Conv2d(n, classes, 1))

    def forward(self, feats):
        return self.out(torch.cat([self.up(feats[1]), feats[0]], 1))


class RecDecoder(nn.Module):
    def __init__(self, n=8, cout=3, n_domains=3):
        super().__init__()
        self.up = nn.ConvTranspose2d(2 * n, n, 2, stride=2)
        self.bn = DomainBN(n, n_domains)
        self.out = nn.Conv2d(n, cout, 3, padding=1)

    def forward(self, f, domain_label):
        return self.out(F.relu(self.bn(self.up(f), int(domain_label))))


def dice_loss(score, target):
    """utils/losses.py:8-16."""
    target = target.float()
    smooth = 1e-5
    inter = torch.sum(score * target)
    return 1 - (2 * inter + smooth) / (torch.sum(score * score) + torch.sum(target * target) + smooth)


class Nets:
    def __init__(self, n_domains, device, seed=1337, lr=2e-3):
        torch.manual_seed(seed)
        self.encoder, self.seg, self.rec = Encoder().to(device), SegDecoder().to(device), RecDecoder(n_domains=n_domains).to(device)
        # three parameter groups, encoder at half the learning rate (train.py:580-587, 291-293)
        self.opt = torch.optim.Adam([{"params": self.encoder.parameters(), "lr": lr / 2},
                                     {"params": self.seg.parameters(), "lr": lr},
                                     {"params": self.rec.parameters(), "lr": lr}], betas=(0.9, 0.99))

    def parameters(self):
        return list(self.encoder.parameters()) + list(self.seg.parameters()) + list(self.rec.parameters())


def train_step(nets: Nets, img, img_freq, mask, batch_sizes, lambda_rec=0.1):
    """One `--ram --rec --consistency` step (train.py:246-287).  Returns python floats."""
    bce = nn.BCELoss()
    mse = nn.MSELoss()
    feats = nets.encoder(img)
    p1 = torch.sigmoid(nets.seg(feats))
    loss_1 = bce(p1, mask) + dice_loss(p1, mask)
    feats_f = nets.encoder(img_freq)                       # second encoder pass on the RAM batch (train.py:252-256)
    p2 = torch.sigmoid(nets.seg(feats_f))
    loss_2 = bce(p2, mask) + dice_loss(p2, mask)
    loss_cons = mse(p2, p1)
    loss = 0.0
    left = 0
    rec_total = 0.0
    for d, n in enumerate(batch_sizes):                    # per-domain reconstruction of the CLEAN image (train.py:263-276)
        rec = torch.tanh(nets.rec(feats_f[-1][left:left + n], d))
        l_rec = mse(rec, img[left:left + n])
        loss = loss + lambda_rec * l_rec
        rec_total += float(l_rec)
        left += n
    loss = loss + loss_1 + loss_2 + 0.5 * loss_cons
    nets.opt.zero_grad()
    loss.backward()
    gn = torch.sqrt(sum((p.grad.double() ** 2).sum() for p in nets.parameters() if p.grad is not None))
    nets.opt.step()
    return {"loss": float(loss), "grad_norm": float(gn), "seg_img": float(loss_1), "seg_freq": float(loss_2),
            "consistency": float(loss_cons), "rec": rec_total}


# ------------------------------------------------------------------------------------------------------------------
# loaders: one per source domain, zipped (train.py:556-566, 223)
def make_loaders(root, batch_sizes, transform, seed, num_workers):
    from ramdsir_b200.datasets import Fundus_Multi
    loaders = []
    for d, bs in enumerate(batch_sizes):
        ds = Fundus_Multi(domain_idx_list=[d], base_dir=root, split="train", transform=transform, test_domain_idx=3,
                          is_out_domain=False, ram="deferred")
        g = torch.Generator().manual_seed(seed + d)
        loaders.append(DataLoader(ds, batch_size=bs, shuffle=True, drop_last=True, num_workers=num_workers,
                                  worker_init_fn=seed_worker, generator=g, pin_memory=True,
                                  persistent_workers=False))
    return loaders


def reseed_main(seed):
    """num_workers=0 runs __getitem__ in this process: seed the streams the loaders' draws come from."""
    np.random.seed(seed)
    random.seed(seed)
    torch.manual_seed(seed)


def gpu_batches(root, batch_sizes, transform, seed, num_workers, device):
    """GPU arm: deferred datasets + RamBatchHook per domain batch, concatenated like train.py:232-244."""
    from ramdsir_b200.datasets import RamBatchHook
    hook = RamBatchHook("fundus", device=device)
    reseed_main(seed)
    loaders = make_loaders(root, batch_sizes, transform, seed, num_workers)
    while True:                                            # epochs (train.py:222-223)
        for parts in zip(*loaders):
            outs = [hook(p) for p in parts]
            yield tuple(torch.cat([o[k] for o in outs], 0) for k in range(3))


def cpu_batches(root, batch_sizes, transform, seed, num_workers, device, cpu_ram_fn):
    """CPU arm: the same samples and draws; img_freq per sample from `cpu_ram_fn(src_chw_f32, partner_chw_f32, ratio_k)`
    (returns the normalised float32 CHW image), img normalised on the host as fundus.py:217-220, then `.cuda()`."""
    reseed_main(seed)
    loaders = make_loaders(root, batch_sizes, transform, seed, num_workers)
    while True:
        for parts in zip(*loaders):
            imgs, freqs, masks = [], [], []
            for src, partner, ratio_k, mask in parts:
                for i in range(src.shape[0]):
                    s = src[i].numpy().astype(np.float32)
                    p = partner[i].numpy().astype(np.float32)
                    freqs.append(torch.from_numpy(np.ascontiguousarray(cpu_ram_fn(s, p, int(ratio_k[i])), dtype=np.float32)))
                    a = s.copy()
                    a /= 127.5
                    a -= 1.0
                    imgs.append(torch.from_numpy(a))
                masks.append(mask)
            yield torch.stack(imgs).to(device), torch.stack(freqs).to(device), torch.cat(masks, 0).to(device)


def run_arm(batches, n_domains, batch_sizes, device, steps, seed=1337):
    torch.backends.cudnn.deterministic = True
    torch.backends.cudnn.benchmark = False
    nets = Nets(n_domains, device, seed=seed)
    recs, t_data, t_step = [], 0.0, 0.0
    it = iter(batches)
    for _ in range(steps):
        t0 = time.perf_counter()
        try:
            img, img_freq, mask = next(it)
        except StopIteration:
            break
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        recs.append(train_step(nets, img, img_freq, mask, batch_sizes))
        torch.cuda.synchronize()
        t2 = time.perf_counter()
        t_data += t1 - t0
        t_step += t2 - t1
    n = max(len(recs), 1)
    return recs, {"steps": len(recs), "data_ms_per_step": 1e3 * t_data / n, "model_ms_per_step": 1e3 * t_step / n}


def compare(root, transform, cpu_ram_fn, steps=2, seed=1337, num_workers=0, batch_sizes=FUNDUS_BATCHES, device=None):
    """Run both arms from identical initial weights on identical draws; return per-step records and timings."""
    device = torch.device("cuda", torch.cuda.current_device()) if device is None else device
    g_recs, g_t = run_arm(gpu_batches(root, batch_sizes, transform, seed, num_workers, device), len(batch_sizes),
                          batch_sizes, device, steps, seed)
    c_recs, c_t = run_arm(cpu_batches(root, batch_sizes, transform, seed, num_workers, device, cpu_ram_fn), len(batch_sizes),
                          batch_sizes, device, steps, seed)
    rel = lambda a, b: abs(a - b) / max(abs(b), 1e-30)
    diffs = [{"loss_rel": rel(g["loss"], c["loss"]), "grad_norm_rel": rel(g["grad_norm"], c["grad_norm"])}
             for g, c in zip(g_recs, c_recs)]
    return {"gpu": g_recs, "cpu": c_recs, "diff": diffs, "gpu_timing": g_t, "cpu_timing": c_t,
            "batch": int(sum(batch_sizes))}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--root", default=None, help="fundus fixture root (default: a fresh synthetic fixture in a temp dir)")
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--workers", type=int, default=0)
    a = ap.parse_args()
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import fixture_data as FX
    root = a.root
    if root is None:
        import tempfile
        root = FX.make_fundus(tempfile.mkdtemp(prefix="ramdsir_fundus_"), counts=E2E_FIXTURE)
    dev = torch.device("cuda", 0)
    recs, t = run_arm(gpu_batches(root, FUNDUS_BATCHES, FX.ResizeJitter(), 1337, a.workers, dev), 3, FUNDUS_BATCHES, dev, a.steps)
    print({"last": recs[-1] if recs else None, "timing": t})


if __name__ == "__main__":
    main()
