"""``replay_buffer.python_randoms``: the PER sampler's uniforms (reference replay_buffer.py:142-147, one ``random.random()`` per
stratum) drawn by one C call from the interpreter's MT19937 state -- same values, same generator state afterwards.  Host-only."""
import random

import numpy as np
import pytest


@pytest.mark.parametrize('seed', [0, 1, 20231019])
@pytest.mark.parametrize('consumed', [0, 5, 623, 624, 625, 1000])
@pytest.mark.parametrize('n', [1, 1023, 1024, 1025, 4096, 5000])
def test_python_randoms_continue_the_interpreters_stream(seed, consumed, n):
    from cacto_b200.replay_buffer import python_randoms
    random.seed(seed)
    for _ in range(consumed):
        random.random()
    want = np.array([random.random() for _ in range(n)])
    after = random.getstate()
    random.seed(seed)
    for _ in range(consumed):
        random.random()
    got = python_randoms(n)
    assert np.array_equal(want, got)
    assert random.getstate() == after
    assert random.random() == (random.setstate(after) or random.random())


def test_python_randoms_keep_a_pending_gaussian():
    from cacto_b200.replay_buffer import python_randoms
    random.seed(3)
    random.gauss(0, 1)                       # leaves gauss_next set: part of the state that must survive
    st = random.getstate()
    a = [random.random() for _ in range(2048)] + [random.gauss(0, 1)]
    random.setstate(st)
    b = list(python_randoms(2048)) + [random.gauss(0, 1)]
    assert a == b


@pytest.mark.parametrize('max_idx', [1, 3000, 65536, 65537, 2 ** 31 + 5])
@pytest.mark.parametrize('B,K', [(64, 37), (1, 5), (4096, 3)])
def test_one_randint_call_equals_consecutive_sample_draws(max_idx, B, K):
    """ReplayBuffer.draw_indices relies on it: K calls of np.random.randint(0, n, B) (reference replay_buffer.py:45) and one call of
    size (K, B) give the same numbers and leave the same generator state."""
    np.random.seed(5)
    a = np.stack([np.random.randint(0, max_idx, size=B) for _ in range(K)])
    sa = np.random.get_state()
    np.random.seed(5)
    b = np.random.randint(0, max_idx, size=(K, B))
    sb = np.random.get_state()
    assert np.array_equal(a, b)
    assert all(np.array_equal(x, y) if isinstance(x, np.ndarray) else x == y for x, y in zip(sa, sb))
