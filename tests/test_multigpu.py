"""Multi-GPU (NCCL) check: bank built sharded + all-gathered is bit-identical to the single-GPU bank,
and the batch-sharded output is bit-identical to the unsharded one.  Needs >= 2 GPUs (skipped on the
1-GPU box; run with `gpurun --gpus 2 -- python -m pytest tests/test_multigpu.py -m gpu`)."""
import os
import socket
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir):
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    import ramdsir_b200 as R
    from ramdsir_b200 import sharded
    g = torch.Generator(device="cuda").manual_seed(5)
    n_bank, B = 37, 10                                        # ragged on purpose (37 = 19 + 18)
    imgs = torch.randint(0, 256, (n_bank, 3, 256, 256), device="cuda", generator=g).float()   # same on every rank
    lo, hi = sharded.shard_range(n_bank, rank, world)
    bank = sharded.build_amp_bank_sharded(imgs[lo:hi].contiguous(), n_bank, 0.1)
    ref = R.build_amp_bank(imgs)
    ok_bank = torch.equal(bank, ref)
    src = torch.randint(0, 256, (B, 3, 256, 256), device="cuda", generator=g).float()
    pidx = (torch.arange(B, device="cuda") * 3) % n_bank
    rk = torch.arange(1, B + 1, dtype=torch.int32, device="cuda")
    full = R.ram_mix_batch(src, ref, pidx, rk)
    blo, bhi = sharded.shard_batch(B, rank, world)
    mine = R.ram_mix_batch(src[blo:bhi].contiguous(), bank, pidx[blo:bhi], rk[blo:bhi])
    ok_out = torch.equal(mine, full[blo:bhi])
    with open(os.path.join(out_dir, "r%d" % rank), "w") as f:
        f.write("%d %d" % (ok_bank, ok_out))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs >= 2 GPUs")
def test_sharded_bank_and_batch_nccl(tmp_path):
    import torch.multiprocessing as mp
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    for r in range(world):
        assert open(tmp_path / ("r%d" % r)).read() == "1 1"
