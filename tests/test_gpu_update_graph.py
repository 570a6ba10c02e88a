"""CUDA-graph replay of the update (RL.UpdateGraph) against the eager launches, and learn_and_update end to end
on the GPU replay buffers (RL.py:120-143)."""
import random

import numpy as np
import pytest
import torch

from cacto_b200.conf import get_conf

pytestmark = pytest.mark.gpu


def build(system='manipulator', seed=0, **over):
    from cacto_b200 import environment as genv
    from cacto_b200.NeuralNetwork import NN
    from cacto_b200.RL import RL_AC
    conf = get_conf(system, **over)
    env = genv.make_env(conf)
    rl = RL_AC(env, NN(env, conf, 1e-2, seed=seed), conf, 0)
    rl.setup_model()
    return conf, rl


def fill(buffer, conf, n_rows, seed=0):
    rng = np.random.default_rng(seed)
    ns = conf.nb_state
    lo, hi = np.asarray(conf.x_init_min, float), np.asarray(conf.x_init_max, float)
    s = rng.uniform(lo, hi, (n_rows, ns))
    sn = rng.uniform(lo, hi, (n_rows, ns))
    dv = rng.normal(size=(n_rows, ns))
    dv[:, -1] = 0
    buffer.add((s,), (rng.uniform(-5, 0, n_rows),), (sn,), (dv,), ((rng.uniform(size=n_rows) < 0.5).astype(float),),
               ((rng.uniform(size=n_rows) < 0.05).astype(float),))


def weights_of(rl):
    return [w.copy() for n in (rl.critic_model, rl.target_critic, rl.actor_model) for w in n.get_weights()]


@pytest.mark.parametrize('lr_schedule', [0, 1])
def test_graph_replay_equals_eager_updates(lr_schedule):
    from cacto_b200.replay_buffer import ReplayBuffer
    conf, rl_e = build(LR_SCHEDULE=lr_schedule)
    _, rl_g = build(LR_SCHEDULE=lr_schedule)
    buf = ReplayBuffer(conf)
    fill(buf, conf, 5000)
    before = weights_of(rl_g)
    g = rl_g.make_update_graph()
    for a, b in zip(before, weights_of(rl_g)):                 # capture must leave the training state untouched
        np.testing.assert_array_equal(a, b)
    assert rl_g.critic_optimizer.iterations == 0 and int(rl_g.critic_optimizer._dev['step'][0]) == 0
    for it in range(4):
        idx = np.random.default_rng(it).integers(0, 5000, conf.BATCH_SIZE)
        batch = buf.sample(idx)
        rtg_e, V_e, Vt_e = rl_e.update(batch[0], batch[2], batch[1], batch[3], batch[4], batch[5], batch[6], fuse_target=True)
        buf.sample(idx, out=g.io)
        rtg_g, V_g, Vt_g = g.replay()
        torch.testing.assert_close(rtg_g, rtg_e, rtol=1e-5, atol=1e-6)
        torch.testing.assert_close(V_g, V_e, rtol=1e-5, atol=1e-6)
    assert rl_g.critic_optimizer.iterations == rl_e.critic_optimizer.iterations == 4
    assert int(rl_g.actor_optimizer._dev['step'][0]) == 4
    for a, b in zip(weights_of(rl_e), weights_of(rl_g)):
        assert np.abs(a - b).max() <= 2e-5 * max(np.abs(a).max(), 1e-3)


@pytest.mark.parametrize('alpha', [0, 0.6])
def test_learn_and_update_runs_with_both_buffers(alpha):
    from cacto_b200.replay_buffer import PrioritizedReplayBuffer, ReplayBuffer
    conf, rl = build(prioritized_replay_alpha=alpha, UPDATE_LOOPS=np.array([12]), save_interval=10 ** 9)
    buf = PrioritizedReplayBuffer(conf) if alpha else ReplayBuffer(conf)
    fill(buf, conf, 3000)
    random.seed(0)
    np.random.seed(0)
    w0 = weights_of(rl)
    rl.use_update_graph = False                    # eager launches first ...
    cnt = rl.learn_and_update(0, buf, 0)
    assert cnt == 12 and rl.actor_optimizer.iterations == 12 and getattr(rl, 'update_graph', None) is None
    w1 = weights_of(rl)
    assert all(np.isfinite(w).all() for w in w1)
    assert any(np.abs(a - b).max() > 0 for a, b in zip(w0, w1))
    rl.use_update_graph = True                     # ... then the default: the graph is built on first use
    cnt = rl.learn_and_update(cnt, buf, 0)
    assert rl.update_graph.B == conf.BATCH_SIZE
    assert cnt == 24 and rl.critic_optimizer.iterations == 24
    assert all(np.isfinite(w).all() for w in weights_of(rl))
    if alpha:
        assert buf._max_priority >= 1.0 and buf.exp_counter.sum() > 0


@pytest.mark.parametrize('system,lr_schedule,mc', [('manipulator', 1, 0), ('car', 0, 0), ('ur5', 0, 0), ('manipulator', 0, 1)])
def test_pipelined_updates_equal_sequential_updates(system, lr_schedule, mc):
    """RL.PipelinedUpdateGraph: the actor step of update i runs beside the critic gradient of update i + 1 (they are independent,
    NeuralNetwork.py:150-178 / RL.py:104-109); after flush() the weights are those of the sequential graph, and every replay
    returns the critic outputs of its own batch."""
    from cacto_b200.replay_buffer import ReplayBuffer
    conf, rl_s = build(system, LR_SCHEDULE=lr_schedule, MC=mc)         # MC = 1: no target network in the loss, no Polyak step
    _, rl_p = build(system, LR_SCHEDULE=lr_schedule, MC=mc)
    buf = ReplayBuffer(conf)
    fill(buf, conf, 5000)
    before = weights_of(rl_p)
    gs, gp = rl_s.make_update_graph(), rl_p.make_pipelined_update_graph()
    for a, b in zip(before, weights_of(rl_p)):                 # capture must leave the training state untouched
        np.testing.assert_array_equal(a, b)
    assert rl_p.critic_optimizer.iterations == 0 and int(rl_p.actor_optimizer._dev['step'][0]) == 0
    n = 7
    for it in range(n):
        idx = np.random.default_rng(it).integers(0, 5000, conf.BATCH_SIZE)
        buf.sample(idx, out=gs.io)
        rtg_s, V_s, _ = gs.replay()
        buf.sample(idx, out=gp.io)
        rtg_p, V_p, _ = gp.replay()
        torch.testing.assert_close(rtg_p, rtg_s, rtol=2e-5, atol=2e-6)
        torch.testing.assert_close(V_p, V_s, rtol=2e-5, atol=2e-6)
        if it == 3:                                            # a flush in the middle (checkpoint): the pipeline restarts
            gp.flush()
            for a, b in zip(weights_of(rl_s), weights_of(rl_p)):
                assert np.abs(a - b).max() <= 2e-5 * max(np.abs(a).max(), 1e-3)
    assert rl_p.actor_optimizer.iterations == n - 1            # the last actor step is outstanding
    gp.flush()
    assert rl_p.critic_optimizer.iterations == rl_p.actor_optimizer.iterations == n
    assert int(rl_p.actor_optimizer._dev['step'][0]) == n and int(rl_p.critic_optimizer._dev['step'][0]) == n
    for a, b in zip(weights_of(rl_s), weights_of(rl_p)):
        assert np.abs(a - b).max() <= 2e-5 * max(np.abs(a).max(), 1e-3)


def test_pipelined_update_graph_refuses_what_it_cannot_overlap():
    conf, rl = build(BATCH_SIZE=4096)
    with pytest.raises(ValueError):
        rl.make_pipelined_update_graph()                       # tcgen05 engine: one workspace shared by the two steps


@pytest.mark.parametrize('alpha', [0, 0.6])
def test_learn_and_update_pipelined_by_default_matches_the_sequential_graph(alpha):
    """learn_and_update on one GPU replays the pipelined graph (default); same index draws -> same weights as with
    use_pipelined_updates = False, PER priorities included (they come from the critic outputs of each batch), checkpoints flushed."""
    from cacto_b200.RL import PipelinedUpdateGraph, UpdateGraph
    from cacto_b200.replay_buffer import PrioritizedReplayBuffer, ReplayBuffer
    out = {}
    for pipelined in (True, False):
        conf, rl = build(prioritized_replay_alpha=alpha, UPDATE_LOOPS=np.array([9]), save_interval=10 ** 9)
        rl.use_pipelined_updates = pipelined
        buf = PrioritizedReplayBuffer(conf) if alpha else ReplayBuffer(conf)
        fill(buf, conf, 3000)
        random.seed(0)
        np.random.seed(0)
        cnt = rl.learn_and_update(0, buf, 0)
        assert cnt == 9 and rl.actor_optimizer.iterations == rl.critic_optimizer.iterations == 9
        assert isinstance(rl.update_graph, PipelinedUpdateGraph if pipelined else UpdateGraph) and getattr(rl.update_graph, 'pending', None) is None
        out[pipelined] = (weights_of(rl), buf._it_sum._value.cpu().numpy().copy() if alpha else None)
    for a, b in zip(out[True][0], out[False][0]):
        assert np.abs(a - b).max() <= 1e-4 * max(np.abs(a).max(), 1e-3)
    if alpha:
        np.testing.assert_allclose(out[True][1], out[False][1], rtol=1e-4, atol=1e-9)


def test_draw_indices_rows_are_the_draws_of_consecutive_samples():
    from cacto_b200.replay_buffer import PrioritizedReplayBuffer, ReplayBuffer
    conf, rl = build()
    buf = ReplayBuffer(conf)
    fill(buf, conf, 3000)
    np.random.seed(3)
    rows = buf.draw_indices(5)
    state = np.random.get_state()
    np.random.seed(3)
    for k in range(5):
        want = np.random.randint(0, buf._max_idx(), size=conf.BATCH_SIZE)
        assert np.array_equal(rows[k].cpu().numpy(), want)
        a, b = buf.sample(rows[k]), buf.sample(want)
        for x, y in zip(a[:7], b[:7]):
            assert torch.equal(x, y)
    assert all(np.array_equal(x, y) if isinstance(x, np.ndarray) else x == y for x, y in zip(state, np.random.get_state()))
    assert PrioritizedReplayBuffer(conf).draw_indices(5) is None
    assert list(PrioritizedReplayBuffer(conf).index_stream(3)) == [None, None, None]
    np.random.seed(3)                                          # the chunked stream: same rows, whatever the chunk size
    got = [r.cpu().numpy() for r in buf.index_stream(5, chunk_indices=2 * conf.BATCH_SIZE)]
    assert len(got) == 5 and all(np.array_equal(g_, rows[k].cpu().numpy()) for k, g_ in enumerate(got))
    assert all(np.array_equal(x, y) if isinstance(x, np.ndarray) else x == y for x, y in zip(state, np.random.get_state()))
    # graph inputs: the ones of the uniform buffer are written once, and again after anything else wrote into the tensor
    g = rl.make_update_graph()
    buf.sample(rows[0], out=g.io)
    assert bool((g.io['weights'] == 1).all())
    g.io['weights'].mul_(0.5)
    buf.sample(rows[1], out=g.io)
    assert bool((g.io['weights'] == 1).all())


def test_learn_and_update_graph_path_draws_the_same_indices_as_the_eager_loop():
    """learn_and_update draws the indices of its whole loop in one call when it replays a graph: same np.random stream as the
    per-update draws of the eager loop (reference replay_buffer.py:45), hence the same training."""
    from cacto_b200.replay_buffer import ReplayBuffer
    out = {}
    for graph in (True, False):
        conf, rl = build(UPDATE_LOOPS=np.array([6]), save_interval=10 ** 9)
        rl.use_update_graph = graph
        buf = ReplayBuffer(conf)
        fill(buf, conf, 3000)
        np.random.seed(11)
        assert rl.learn_and_update(0, buf, 0) == 6
        out[graph] = (weights_of(rl), np.random.get_state()[1].copy())
    assert np.array_equal(out[True][1], out[False][1])
    for a, b in zip(out[True][0], out[False][0]):
        assert np.abs(a - b).max() <= 1e-4 * max(np.abs(a).max(), 1e-3)
