"""world_size-2 gloo tests (CPU) for the multi-GPU plumbing: shard ranges, sharded bank build written
in place into the gather buffer, one all-gather, batch sharding with GLOBAL partner indices.
The CUDA bank builder is replaced by an injected stand-in (the numpy oracle) -- this checks the host
logic only; the real kernels are covered by the -m gpu tests."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _oracle_build(imgs, L, out):
    from oracle import ram_oracle as O
    out.copy_(torch.from_numpy(O.build_bank(imgs.numpy(), L).astype(np.float32)))
    return out


def _worker(rank, world, port, n_total, tmp):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from ramdsir_b200 import sharded
    from oracle import ram_oracle as O
    rs = np.random.RandomState(0)
    imgs = rs.rand(n_total, 2, 32, 48).astype(np.float32)          # every rank can regenerate the full set
    lo, hi = sharded.shard_range(n_total, rank, world)
    bank = sharded.build_amp_bank_sharded(torch.from_numpy(imgs[lo:hi]), n_total, 0.1, _build_fn=_oracle_build)
    want = O.build_bank(imgs, 0.1).astype(np.float32)
    ok = bank.shape == want.shape and np.array_equal(bank.numpy(), want)
    # batch sharding: global partner rows index the replicated bank
    B = 11
    pidx = rs.randint(0, n_total, B)
    blo, bhi = sharded.shard_batch(B, rank, world)
    mine = bank[torch.from_numpy(pidx[blo:bhi])]
    ok = ok and np.array_equal(mine.numpy(), want[pidx[blo:bhi]])
    np.save(os.path.join(tmp, "ok%d.npy" % rank), np.array([ok, lo, hi, blo, bhi]))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n_total", [8, 7, 1])
def test_sharded_bank_allgather_world2(tmp_path, n_total):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), n_total, str(tmp_path)), nprocs=world, join=True)
    r = [np.load(tmp_path / ("ok%d.npy" % i)) for i in range(world)]
    assert all(x[0] == 1 for x in r)
    # contiguous, disjoint, covering shards
    assert r[0][1] == 0 and r[0][2] == r[1][1] and r[1][2] == n_total
    assert r[0][3] == 0 and r[0][4] == r[1][3] and r[1][4] == 11


def test_shard_ranges_cover_everything():
    from ramdsir_b200 import sharded
    for n in (0, 1, 7, 8, 768, 789, 8192):
        for world in (1, 2, 4, 8):
            spans = [sharded.shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert max(hi - lo for lo, hi in spans) == sharded.padded_rows(n, world) or n == 0


def test_rng_draws_match_reference_calls():
    """np.random.choice(list, 1)[0] / np.random.choice(n) / random.randint(1, 10) in the reference's order
    (fundus.py:205,208,35) consume the global streams exactly like the helper does."""
    import random
    from ramdsir_b200 import rng
    names = ["Domain1", "Domain2", "Domain4"]
    sizes = {"Domain1": 50, "Domain2": 99, "Domain4": 320}
    rng.seed_worker_streams(1337)
    got = [(rng.draw_partner(names, sizes), rng.draw_ratio_k()) for _ in range(20)]
    np.random.seed(1337)
    random.seed(1337)
    want = []
    for _ in range(20):
        dom = np.random.choice(names, 1)[0]                        # fundus.py:205
        lines = ["x%d" % i for i in range(sizes[str(dom)])]
        pick = np.random.choice(lines)                             # fundus.py:208
        want.append(((str(dom), lines.index(pick)), random.randint(1, 10)))   # fundus.py:35
    assert got == want
