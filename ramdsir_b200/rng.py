"""Host-side draws that reproduce the reference data loader's RNG streams exactly.

The reference decides, per training sample,
  1. the partner DOMAIN:  np.random.choice(domain_list, 1)[0]          fundus.py:205 / prostate.py:182
  2. the partner IMAGE:   np.random.choice(lines_or_files)             fundus.py:208 / prostate.py:183
  3. the blend ratio:     random.randint(1, 10) / 10                   fundus.py:35
with the NUMPY global stream for (1, 2) and PYTHON's global `random` for (3), both reseeded per
DataLoader worker by `seed_worker` (train.py:80-83).  These helpers make the same calls in the same
order and return integers, so the GPU path receives bit-identical selections.
"""
from __future__ import annotations

import random

import numpy as np

__all__ = ["seed_worker_streams", "draw_partner", "draw_ratio_k", "BankIndex"]


def seed_worker_streams(worker_seed: int) -> None:
    """What train.py:80-83 does inside each worker (worker_seed = torch.initial_seed() % 2**32)."""
    np.random.seed(worker_seed)
    random.seed(worker_seed)


def draw_ratio_k() -> int:
    """fundus.py:35: ratio = random.randint(1, 10) / 10 -> we keep the integer numerator."""
    return random.randint(1, 10)


def draw_partner(domain_names, n_per_domain):
    """Two np.random.choice draws, exactly as fundus.py:205,208 / prostate.py:182-183.

    domain_names: candidate partner domains (already minus the test domain, minus the sample's own
    domain when is_out_domain).  n_per_domain: mapping name -> number of candidate images.
    Returns (domain_name, index_within_domain).
    """
    dom = np.random.choice(domain_names, 1)[0]
    # np.random.choice(seq) == seq[np.random.randint(0, len(seq))] on the same stream
    idx = int(np.random.choice(n_per_domain[str(dom)]))
    return str(dom), idx


class BankIndex:
    """Maps (domain, index within the domain) to a row of a concatenated amplitude bank."""

    def __init__(self, domain_sizes):
        self.offsets = {}
        off = 0
        for name, n in domain_sizes:
            self.offsets[name] = off
            off += int(n)
        self.sizes = {name: int(n) for name, n in domain_sizes}
        self.total = off

    def row(self, domain, idx):
        if not 0 <= idx < self.sizes[domain]:
            raise IndexError("partner index %d outside domain %s of %d images" % (idx, domain, self.sizes[domain]))
        return self.offsets[domain] + idx
