"""Device-side partner / ratio selection (ramdsir_rng_*) against the host streams it must reproduce bit for bit:
np.random.choice(domain_list, 1)[0], np.random.choice(partner_list), random.randint(1, 10) per sample
(fundus.py:205,208,35; prostate.py:182-183) after seed_worker (train.py:80-83).  Also the device mask encoder."""
import random

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _host_stream(seed, names, sizes, index, n):
    from ramdsir_b200 import rng
    rng.seed_worker_streams(seed)
    pidx, rk = np.empty(n, np.int64), np.empty(n, np.int32)
    for i in range(n):
        dom, j = rng.draw_partner(names, sizes)
        pidx[i] = index.row(dom, j)
        rk[i] = rng.draw_ratio_k()
    return pidx, rk, np.random.get_state(), random.getstate()


def test_device_streams_equal_host_streams_one_million_draws():
    from ramdsir_b200 import rng
    # real fundus list sizes (dataset/fundus/Domain{1..4}_train.list: 50, 99, 320, 320) plus corner cases:
    # a power-of-two domain (no rejection), a single-image domain (numpy consumes nothing), a 1-candidate stream
    domains = [("Domain1", 50), ("Domain2", 99), ("Domain3", 320), ("Domain4", 320), ("D5", 64), ("D6", 1)]
    names = [d for d, _ in domains]
    sizes = {d: n for d, n in domains}
    index = rng.BankIndex(domains)
    cands = [[0, 1, 2], [1, 2, 3], [0, 1, 2, 3, 4, 5], [4, 5], [5], [2], [0, 3], [1, 2, 4, 5]]
    seeds = [1337, 1338, 7, 2 ** 32 - 1, 0, 123456789, 42, 99]
    n = 125000                                               # x 8 streams = 10^6 samples = 3 * 10^6 bounded draws
    sel = rng.DeviceSelector(seeds, [s for _, s in domains], cands)
    got_p, got_r = sel.draw(n)
    # the streams persist: a second call continues where the first stopped
    got_p2, got_r2 = sel.draw(1000)
    torch.cuda.synchronize()
    for s, (seed, c) in enumerate(zip(seeds, cands)):
        want_p, want_r, np_state, py_state = _host_stream(seed, [names[k] for k in c], sizes, index, n + 1000)
        np.testing.assert_array_equal(got_p[s].cpu().numpy(), want_p[:n])
        np.testing.assert_array_equal(got_r[s].cpu().numpy(), want_r[:n])
        np.testing.assert_array_equal(got_p2[s].cpu().numpy(), want_p[n:])
        np.testing.assert_array_equal(got_r2[s].cpu().numpy(), want_r[n:])
        # and the generators themselves end in the same state (same number of 32-bit outputs consumed)
        dn, dp = sel.get_state(s)
        np.testing.assert_array_equal(dn[:624], np_state[1])
        assert int(dn[624]) == np_state[2]
        np.testing.assert_array_equal(dp[:624], np.array(py_state[1][:624], dtype=np.uint64).astype(np.uint32))
        assert int(dp[624]) == py_state[1][624]
    assert int(got_r.min()) >= 1 and int(got_r.max()) <= 10
    sel.close()


def test_device_selection_feeds_the_mix_kernel():
    """The selector's outputs are the kernel's inputs: same result as the host-drawn selection."""
    import ramdsir_b200 as R
    from ramdsir_b200 import rng
    g = torch.Generator(device="cuda").manual_seed(3)
    per, C, H, W, B = 6, 3, 256, 256, 8
    par = torch.randint(0, 256, (3 * per, C, H, W), device="cuda", generator=g).float()
    src = torch.randint(0, 256, (B, C, H, W), device="cuda", generator=g).float()
    bank = R.build_amp_bank(par)
    sel = rng.DeviceSelector([1337], [per, per, per], [[0, 1, 2]])
    pidx, rk = sel.draw(B)
    got = R.ram_mix_batch(src, bank, pidx[0], rk[0], clip=(0.0, 255.0), scale=1 / 127.5, shift=-1.0)
    names = ["a", "b", "c"]
    hp, hr, _, _ = _host_stream(1337, names, {n: per for n in names}, rng.BankIndex([(n, per) for n in names]), B)
    want = R.ram_mix_batch(src, bank, hp, hr, clip=(0.0, 255.0), scale=1 / 127.5, shift=-1.0)
    assert torch.equal(got, want)


def test_device_mask_encoding_is_bit_exact():
    from PIL import Image

    from ramdsir_b200.datasets import encode_fundus_mask, encode_fundus_mask_batch
    rs = np.random.RandomState(0)
    m = rs.randint(0, 256, (5, 97, 131)).astype(np.uint8)
    m[0, :3, :6] = [[0, 50, 51, 200, 201, 255]] * 3                       # the thresholds themselves
    got = encode_fundus_mask_batch(torch.from_numpy(m).cuda())
    for i in range(5):
        want = encode_fundus_mask(Image.fromarray(m[i]))
        assert torch.equal(got[i].cpu(), want)
    assert got.dtype == torch.float32 and tuple(got.shape) == (5, 2, 97, 131)
