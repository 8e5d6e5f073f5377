"""Host side of the device selector: the MT19937 states that seed the device streams are exactly
what np.random.seed(s) / random.seed(s) leave behind (train.py:80-83).  No GPU needed."""
import random

import numpy as np
import pytest


@pytest.mark.parametrize("seed", [0, 1, 1337, 2 ** 31 + 5, 2 ** 32 - 1])
def test_seeded_states_equal_numpy_and_cpython(seed):
    from ramdsir_b200 import rng
    np_state, py_state = rng.seed_states_host(seed)
    want_np = np.random.RandomState(seed).get_state()
    assert want_np[0] == "MT19937"
    np.testing.assert_array_equal(np_state[:624], want_np[1])
    assert int(np_state[624]) == want_np[2] == 624
    ver, internal, _ = random.Random(seed).getstate()
    assert ver == 3
    np.testing.assert_array_equal(py_state[:624], np.array(internal[:624], dtype=np.uint64).astype(np.uint32))
    assert int(py_state[624]) == internal[624] == 624
