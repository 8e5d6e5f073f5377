"""Property-style GPU parity: random shapes incl. odd / non-square, C in {1,2,3},
every ratio, uint8-valued and [-1,1]-valued inputs, random band fractions -- CUDA path vs the numpy float64
oracle.  Shapes that are multiples of (16, 64) exercise the runtime-shape fused kernel, the rest the generic one."""
import math

import numpy as np
import pytest
import torch
from hypothesis import HealthCheck, given, settings
from hypothesis import strategies as st

from oracle import ram_oracle as O

pytestmark = pytest.mark.gpu


def _field(rs, shape, lo, hi, integer):
    n, c, h, w = shape
    z = rs.standard_normal(shape)
    fy = np.fft.fftfreq(h)[:, None]
    fx = np.fft.fftfreq(w)[None, :]
    y = np.real(np.fft.ifft2(np.fft.fft2(z) / (1.0 + (np.hypot(fy, fx) / 0.08) ** 2)))
    y = (y - y.min()) / (y.max() - y.min() + 1e-12)
    y = lo + (hi - lo) * y
    return (np.round(y) if integer else y).astype(np.float32)


shape_st = st.one_of(
    st.tuples(st.integers(8, 97), st.integers(8, 97)),                                   # anything -> generic
    st.tuples(st.sampled_from([32, 48, 64, 96, 128]), st.sampled_from([64, 128, 192])),  # fused runtime-shape
)


@settings(max_examples=25, deadline=None, suppress_health_check=[HealthCheck.too_slow, HealthCheck.function_scoped_fixture])
@given(hw=shape_st, c=st.integers(1, 3), L=st.sampled_from([0.03, 0.05, 0.1, 0.15, 0.2]), integer=st.booleans(),
       seed=st.integers(0, 2 ** 16))
def test_random_shapes_match_oracle(hw, c, L, integer, seed):
    import ramdsir_b200 as R
    h, w = hw
    b = O.band_halfwidth(h, w, L)
    if 2 * b + 1 > min(h, w):
        return
    rs = np.random.RandomState(seed)
    lo, hi = (0, 255) if integer else (-1, 1)
    src = _field(rs, (3, c, h, w), lo, hi, integer)
    par = _field(rs, (2, c, h, w), lo, hi, integer)
    pidx = rs.randint(0, 2, 3)
    rk = rs.randint(1, 11, 3).astype(np.int32)
    epi = O.FUNDUS_EPILOGUE if integer else O.PROSTATE_EPILOGUE
    bank = R.build_amp_bank(torch.from_numpy(par).cuda(), L=L)
    out = R.ram_mix_batch(torch.from_numpy(src).cuda(), bank, pidx, rk, L=L, clip=epi[:2], scale=epi[2], shift=epi[3])
    want = O.mix_batch(src, par, pidx, rk, L, epi=epi)
    got = out.cpu().numpy()
    assert got.shape == want.shape
    assert O.rel_l2(got, want) <= 1e-5 and O.max_abs(got, want) <= 1e-4, (hw, c, L, integer, seed)
    assert O.rel_l2(bank.cpu().numpy(), O.build_bank(par, L)) <= 2e-6
