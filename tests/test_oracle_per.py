"""oracle.per pinned bit-exactly against the reference's segment_tree.py / replay_buffer.py."""
import numpy as np
import pytest

from conftest import golden
from oracle import per
from types import SimpleNamespace


def test_segment_tree_matches_reference():
    g = golden('per_segment_tree.npz')
    cap = int(g['cap'])
    s, m = per.SumSegmentTree(cap), per.MinSegmentTree(cap)
    for i, v in zip(g['w_idx'], g['w_val']):
        s[int(i)] = float(v)
        m[int(i)] = float(v)
    np.testing.assert_array_equal(np.array(s.val), g['sum_tree'])
    np.testing.assert_array_equal(np.array(m.val), g['min_tree'])
    for (a, b), rs, rm in zip(g['ranges'], g['range_sum'], g['range_min']):
        b = None if b == -999 else int(b)
        assert s.sum(int(a), b) == rs and m.min(int(a), b) == rm
    assert [s.find_prefixsum_idx(float(q)) for q in g['queries']] == list(g['found'])


def _eps(g, tag):
    lens = g[f'{tag}_lens']
    out = []
    for k in range(6):
        a = g[f'{tag}_{k}']
        out.append(tuple(np.split(a, np.cumsum(lens)[:-1], axis=0)))
    return out


def test_uniform_buffer_matches_reference():
    g = golden('per_uniform.npz')
    conf = SimpleNamespace(REPLAY_SIZE=int(g['R']), BATCH_SIZE=int(g['B']), nb_state=int(g['ns']))
    rb = per.ReplayBuffer(conf)
    for r in range(4):
        rb.add(*_eps(g, f'add{r}'))
        np.testing.assert_array_equal(rb.storage_mat, g[f'storage{r}'])
        assert rb.next_idx == int(g[f'next_idx{r}'])
        out = rb.sample(g[f'idx{r}'])
        for k in range(7):
            np.testing.assert_array_equal(out[k], g[f'sample{r}_{k}'])
            assert out[k].dtype == g[f'sample{r}_{k}'].dtype


@pytest.mark.parametrize('tag', ['small', 'medium'])
def test_prioritized_buffer_matches_reference(tag):
    g = golden(f'per_{tag}.npz')
    conf = SimpleNamespace(REPLAY_SIZE=int(g['R']), BATCH_SIZE=int(g['B']), nb_state=int(g['ns']), prioritized_replay_alpha=0.6,
                           prioritized_replay_beta=0.6, prioritized_replay_eps=1e-2, fresh_factor=0.95)
    pb = per.PrioritizedReplayBuffer(conf)
    for r in range(int(g['rounds'])):
        pb.add(*_eps(g, f'add{r}'))
        for it in range(2):
            out = pb.sample(g[f'u{r}_{it}'])
            np.testing.assert_array_equal(out[7], g[f'idx{r}_{it}'])
            np.testing.assert_array_equal(out[6], g[f'w{r}_{it}'])
            for k in range(6):
                np.testing.assert_array_equal(out[k], g[f'sample{r}_{it}_{k}'])
            pb.update_priorities(out[7], g[f'rtg{r}_{it}'], g[f'V{r}_{it}'])
            assert pb._max_priority == float(g[f'maxp{r}_{it}'])
            np.testing.assert_array_equal(pb.exp_counter, g[f'expc{r}_{it}'])
            if f'sum{r}_{it}' in g:
                np.testing.assert_array_equal(np.array(pb._it_sum.val), g[f'sum{r}_{it}'])
                np.testing.assert_array_equal(np.array(pb._it_min.val), g[f'min{r}_{it}'])
    np.testing.assert_array_equal(np.array(pb._it_sum.val), g['sum_final'])
    np.testing.assert_array_equal(np.array(pb._it_min.val), g['min_final'])
