"""Multi-GPU plumbing for the RAM path: one process per GPU, batch sharded, bank replicated once.

The path shards naturally: every image is independent once the partner's
low-band amplitude patch is addressable.  So
  * the amplitude bank is BUILT sharded -- rank r builds rows [r*per, (r+1)*per) straight into its
    slice of the gather buffer (no staging copy) -- and replicated with ONE all-gather
    (`torch.distributed.all_gather_into_tensor`, NCCL over NVLink on the GPU box);
  * afterwards every rank mixes its contiguous batch shard with no further communication;
    `partner_idx` are GLOBAL bank rows, `ratio_k` is sharded alongside the batch.
There is no reference counterpart (the reference never initialises torch.distributed, train.py:205-208
uses nn.DataParallel for the model only).
"""
from __future__ import annotations

import torch
import torch.distributed as dist

from .api import band_halfwidth, build_amp_bank

__all__ = ["shard_range", "padded_rows", "build_amp_bank_sharded", "shard_batch"]


def padded_rows(n_total: int, world: int) -> int:
    """Rows per rank; the gather needs equal counts, so the bank is padded to world * per rows."""
    return (int(n_total) + world - 1) // world


def shard_range(n_total: int, rank: int, world: int):
    """Contiguous [lo, hi) of bank rows (or batch images) owned by `rank`."""
    per = padded_rows(n_total, world)
    lo = min(rank * per, n_total)
    return lo, min(lo + per, n_total)


def shard_batch(batch: int, rank: int, world: int):
    return shard_range(batch, rank, world)


def build_amp_bank_sharded(imgs_local: torch.Tensor, n_total: int, L: float = 0.1, group=None, _build_fn=None):
    """imgs_local: this rank's source images [hi - lo, C, H, W] for (lo, hi) = shard_range(n_total, rank, world).

    Returns the full replicated bank [n_total, C, 2b+1, 2b+1] on every rank (a view of the padded
    gather buffer).  `_build_fn(imgs, L, out)` exists so the gloo/CPU tests can exercise the plumbing
    without CUDA; the product default is the CUDA bank builder and nothing else.
    """
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    build = build_amp_bank if _build_fn is None else _build_fn
    lo, hi = shard_range(n_total, rank, world)
    n_local, c, h, w = imgs_local.shape
    if n_local != hi - lo:
        raise ValueError("rank %d owns bank rows [%d, %d) but got %d images" % (rank, lo, hi, n_local))
    b = band_halfwidth(h, w, L)
    nb = 2 * b + 1
    per = padded_rows(n_total, world)
    full = torch.zeros((world * per, c, nb, nb), dtype=torch.float32, device=imgs_local.device)
    mine = full[rank * per: rank * per + n_local]
    if n_local:
        build(imgs_local, L, mine)                    # written in place: zero-copy into the gather buffer
    if world > 1:
        shard = full[rank * per:(rank + 1) * per]
        if full.device.type == "cpu":
            shard = shard.clone()                     # gloo does not take an input aliasing the output
        dist.all_gather_into_tensor(full, shard, group=group)   # NCCL: in place, input = output + rank * count
    return full[:n_total]
