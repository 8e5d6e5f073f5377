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

__all__ = ["seed_worker_streams", "draw_partner", "draw_ratio_k", "BankIndex", "DeviceSelector", "seed_states_host"]


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


class DeviceSelector:
    """Partner / ratio selection ON THE DEVICE with the reference's MT19937 streams (ramdsir_rng_* in include/ramdsir.h).

    One stream per DataLoader worker, seeded like `seed_worker` seeds numpy and `random` (train.py:80-83).
    `candidates[s]` lists the partner domains stream s may draw (domain INDICES into `domain_sizes`; the caller has
    removed the test domain and, for is_out_domain, the stream's own domain -- fundus.py:161-162,202-203).
    `draw(n)` returns (partner_idx [S, n] int64, ratio_k [S, n] int32) CUDA tensors: element [s, i] is what the i-th
    next `draw_partner` / `draw_ratio_k` pair would return on the host for that worker, bank row = domain offset + index.
    """

    MAX_CAND = 8

    def __init__(self, worker_seeds, domain_sizes, candidates, device=None):
        import ctypes

        import torch

        from . import _lib
        self._torch, self._lib, self._libmod = torch, _lib.load(), _lib
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        seeds = np.ascontiguousarray(np.asarray(worker_seeds, dtype=np.int64) % (2 ** 32), dtype=np.uint32)
        sizes = np.ascontiguousarray(domain_sizes, dtype=np.int32)
        self.n_streams = int(seeds.size)
        if len(candidates) != self.n_streams:
            raise ValueError("one candidate list per stream")
        ncand = np.array([len(c) for c in candidates], np.int32)
        cand = np.full((self.n_streams, self.MAX_CAND), -1, np.int32)
        for s, c in enumerate(candidates):
            if not 1 <= len(c) <= self.MAX_CAND:
                raise ValueError("a stream needs 1..%d candidate domains" % self.MAX_CAND)
            cand[s, :len(c)] = c
        self._h = ctypes.c_void_p()
        with torch.cuda.device(self.device):
            _lib.check(self._lib.ramdsir_rng_create(ctypes.byref(self._h), self.n_streams, seeds.ctypes.data, int(sizes.size),
                                                    sizes.ctypes.data, ncand.ctypes.data, cand.ctypes.data))

    def draw(self, n_per_stream: int, ratio_lo: int = 1, ratio_hi: int = 10, stream=None):
        torch = self._torch
        pidx = torch.empty((self.n_streams, n_per_stream), dtype=torch.int64, device=self.device)
        rk = torch.empty((self.n_streams, n_per_stream), dtype=torch.int32, device=self.device)
        with torch.cuda.device(self.device):
            s = torch.cuda.current_stream() if stream is None else stream
            self._libmod.check(self._lib.ramdsir_rng_draw(self._h, int(n_per_stream), pidx.data_ptr(), rk.data_ptr(),
                                                         int(ratio_lo), int(ratio_hi), int(s.cuda_stream)))
        return pidx, rk

    def get_state(self, stream_index: int):
        """(numpy_state, python_state) as uint32[625] each: 624 raw MT19937 words + position, numpy's get_state() layout."""
        a, b = np.empty(625, np.uint32), np.empty(625, np.uint32)
        self._libmod.check(self._lib.ramdsir_rng_get_state(self._h, int(stream_index), a.ctypes.data, b.ctypes.data))
        return a, b

    def close(self):
        if getattr(self, "_h", None) is not None and self._h:
            self._lib.ramdsir_rng_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def seed_states_host(worker_seed: int):
    """(numpy_state, python_state) uint32[625] right after np.random.seed(s) / random.seed(s) (ramdsir_rng_seed_host)."""
    from . import _lib
    a, b = np.empty(625, np.uint32), np.empty(625, np.uint32)
    _lib.check(_lib.load().ramdsir_rng_seed_host(int(worker_seed) % (2 ** 32), a.ctypes.data, b.ctypes.data))
    return a, b
