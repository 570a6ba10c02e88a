"""CUDA segment trees / replay buffers: bit-exact against the reference-generated goldens and the oracle."""
import random
from types import SimpleNamespace

import numpy as np
import pytest
import torch

from conftest import golden
from oracle import per as oper

pytestmark = pytest.mark.gpu


def _eps(g, tag):
    lens = g[f'{tag}_lens']
    return [tuple(np.split(g[f'{tag}_{k}'], np.cumsum(lens)[:-1], axis=0)) for k in range(6)]


def test_segment_tree_scalar_api_matches_reference():
    from cacto_b200.segment_tree import MinSegmentTree, SumSegmentTree
    g = golden('per_segment_tree.npz')
    cap = int(g['cap'])
    s, m = SumSegmentTree(cap), MinSegmentTree(cap)
    for i, v in zip(g['w_idx'][:40], g['w_val'][:40]):          # scalar writes
        s[int(i)] = float(v)
        m[int(i)] = float(v)
    s.set_batch(g['w_idx'][40:], g['w_val'][40:])                 # batched writes with duplicates
    m.set_batch(g['w_idx'][40:], g['w_val'][40:])
    np.testing.assert_array_equal(s._value.cpu().numpy(), g['sum_tree'])
    np.testing.assert_array_equal(m._value.cpu().numpy(), g['min_tree'])
    for (a, b), rs, rm in zip(g['ranges'], g['range_sum'], g['range_min']):
        b = None if b == -999 else int(b)
        assert s.sum(int(a), b) == rs and m.min(int(a), b) == rm
    assert s.find_prefixsum_idx_batch(g['queries']).cpu().tolist() == list(g['found'])
    assert s.find_prefixsum_idx(float(g['queries'][0])) == int(g['found'][0])
    assert s[int(g['w_idx'][-1])] == float(g['w_val'][-1])


def test_uniform_buffer_matches_reference():
    from cacto_b200.replay_buffer import ReplayBuffer
    g = golden('per_uniform.npz')
    conf = SimpleNamespace(REPLAY_SIZE=int(g['R']), BATCH_SIZE=int(g['B']), nb_state=int(g['ns']))
    rb = ReplayBuffer(conf)
    for r in range(4):
        rb.add(*_eps(g, f'add{r}'))
        np.testing.assert_array_equal(rb.storage_mat.cpu().numpy(), g[f'storage{r}'])
        assert rb.next_idx == int(g[f'next_idx{r}'])
        np.random.seed(100 + r)
        out = rb.sample()                                          # host np.random stream, as the reference
        for k in range(7):
            ref = g[f'sample{r}_{k}']
            got = out[k].cpu().numpy()
            np.testing.assert_array_equal(got, ref)
            assert got.dtype == ref.dtype
        assert out[7] is None


@pytest.mark.parametrize('tag', ['small', 'medium'])
def test_prioritized_buffer_matches_reference(tag):
    from cacto_b200.replay_buffer import PrioritizedReplayBuffer
    g = golden(f'per_{tag}.npz')
    conf = SimpleNamespace(REPLAY_SIZE=int(g['R']), BATCH_SIZE=int(g['B']), nb_state=int(g['ns']), prioritized_replay_alpha=0.6,
                           prioritized_replay_beta=0.6, prioritized_replay_eps=1e-2, fresh_factor=0.95)
    pb = PrioritizedReplayBuffer(conf)
    for r in range(int(g['rounds'])):
        pb.add(*_eps(g, f'add{r}'))
        for it in range(2):
            random.seed(1000 * r + it)                             # the reference's random.random() stream
            out = pb.sample()
            np.testing.assert_array_equal(out[7], g[f'idx{r}_{it}'])
            np.testing.assert_array_equal(out[6].cpu().numpy(), g[f'w{r}_{it}'])
            for k in range(6):
                np.testing.assert_array_equal(out[k].cpu().numpy(), g[f'sample{r}_{it}_{k}'])
            pb.update_priorities(out[7], torch.tensor(g[f'rtg{r}_{it}'], device='cuda'), torch.tensor(g[f'V{r}_{it}'], device='cuda'))
            assert pb._max_priority == float(g[f'maxp{r}_{it}'])
            np.testing.assert_array_equal(pb.exp_counter, g[f'expc{r}_{it}'])
            if f'sum{r}_{it}' in g:
                np.testing.assert_array_equal(pb._it_sum._value.cpu().numpy(), g[f'sum{r}_{it}'])
                np.testing.assert_array_equal(pb._it_min._value.cpu().numpy(), g[f'min{r}_{it}'])
    np.testing.assert_array_equal(pb._it_sum._value.cpu().numpy(), g['sum_final'])
    np.testing.assert_array_equal(pb._it_min._value.cpu().numpy(), g['min_final'])


def test_full_size_per_round_matches_oracle():
    """BASELINE config 2 sizes: capacity 2^16 full, batch 4096, alpha = beta = 0.6; three rounds."""
    from cacto_b200.replay_buffer import PrioritizedReplayBuffer
    ns = 5
    conf = SimpleNamespace(REPLAY_SIZE=2 ** 16, BATCH_SIZE=4096, nb_state=ns, prioritized_replay_alpha=0.6,
                           prioritized_replay_beta=0.6, prioritized_replay_eps=1e-4, fresh_factor=1)
    rng = np.random.default_rng(0)
    rows = rng.normal(size=(2 ** 16 + 1000, 3 * ns + 3))
    pb, ob = PrioritizedReplayBuffer(conf), oper.PrioritizedReplayBuffer(conf)
    cols = (rows[:, :ns], rows[:, ns], rows[:, ns + 1:2 * ns + 1], rows[:, 2 * ns + 1:3 * ns + 1], rows[:, 3 * ns + 1], rows[:, 3 * ns + 2])
    pb.add(*[(c,) for c in cols])
    ob.add(*[(c,) for c in cols])
    assert pb.full == 1 and pb.next_idx == ob.next_idx == 1000
    for r in range(3):
        u = rng.uniform(size=4096)
        got, ref = pb.sample(u), ob.sample(u)
        np.testing.assert_array_equal(got[7], ref[7])
        np.testing.assert_array_equal(got[6].cpu().numpy(), ref[6])
        np.testing.assert_array_equal(got[0].cpu().numpy(), ref[0])
        np.testing.assert_array_equal(got[5].cpu().numpy(), ref[5])
        rtg = rng.normal(size=(4096, 1)).astype(np.float32)
        V = rng.normal(size=(4096, 1)).astype(np.float32)
        pb.update_priorities(got[7], torch.tensor(rtg, device='cuda'), torch.tensor(V, device='cuda'))
        ob.update_priorities(ref[7], rtg, V)
        np.testing.assert_array_equal(pb._it_sum._value.cpu().numpy(), np.array(ob._it_sum.val))
        np.testing.assert_array_equal(pb._it_min._value.cpu().numpy(), np.array(ob._it_min.val))
        assert pb._max_priority == ob._max_priority
    # checksum-of-sums property: the root equals the left-to-right pairwise tree sum of the leaves
    leaves = pb._it_sum._value[2 ** 16:].cpu().numpy()
    lvl = leaves
    while len(lvl) > 1:
        lvl = lvl[0::2] + lvl[1::2]
    assert lvl[0] == pb._it_sum.sum()


def test_relo_priorities_match_oracle():
    """RB_type = 'ReLO' (replay_buffer.py:193-196): priorities from MSE(rtg, V) - MSE(rtg, V_target), clipped at 0; trees bit-exact."""
    from cacto_b200.replay_buffer import PrioritizedReplayBuffer
    ns = 3
    conf = SimpleNamespace(REPLAY_SIZE=256, BATCH_SIZE=32, nb_state=ns, prioritized_replay_alpha=0.6,
                           prioritized_replay_beta=0.6, prioritized_replay_eps=1e-4, fresh_factor=0.95)
    rng = np.random.default_rng(3)
    rows = rng.normal(size=(200, 3 * ns + 3))
    pb, ob = PrioritizedReplayBuffer(conf), oper.PrioritizedReplayBuffer(conf)
    pb.RB_type = ob.RB_type = 'ReLO'
    cols = (rows[:, :ns], rows[:, ns], rows[:, ns + 1:2 * ns + 1], rows[:, 2 * ns + 1:3 * ns + 1], rows[:, 3 * ns + 1], rows[:, 3 * ns + 2])
    pb.add(*[(c,) for c in cols])
    ob.add(*[(c,) for c in cols])
    for r in range(3):
        u = rng.uniform(size=32)
        got, ref = pb.sample(u), ob.sample(u)
        np.testing.assert_array_equal(got[7], ref[7])
        rtg = rng.normal(size=(32, 1)).astype(np.float32)
        V = rng.normal(size=(32, 1)).astype(np.float32)
        Vt = rng.normal(size=(32, 1)).astype(np.float32)
        pb.update_priorities(got[7], torch.tensor(rtg, device='cuda'), torch.tensor(V, device='cuda'), torch.tensor(Vt, device='cuda'))
        ob.update_priorities(ref[7], rtg, V, Vt)
        np.testing.assert_array_equal(pb._it_sum._value.cpu().numpy(), np.array(ob._it_sum.val))
        np.testing.assert_array_equal(pb._it_min._value.cpu().numpy(), np.array(ob._it_min.val))
        assert pb._max_priority == ob._max_priority
