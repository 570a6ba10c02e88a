"""CUDA reward-to-go windows: bit-exact against RL_AC.RL_Solve goldens and the oracle."""
from types import SimpleNamespace

import numpy as np
import pytest

from conftest import golden
from oracle import rtg as ortg

pytestmark = pytest.mark.gpu


def test_rtg_matches_reference_goldens():
    from cacto_b200.rtg import rtg_batch
    g = golden('rtg_cases.npz')
    for k in range(int(g['ncases'])):
        T, n, MC, ns = [int(x) for x in g[f'c{k}_meta']]
        conf = SimpleNamespace(nb_state=ns, MC=MC, nsteps_TD_N=n)
        out = rtg_batch(conf, [g[f'c{k}_states']], [g[f'c{k}_cost']])
        np.testing.assert_array_equal(out['partial'].cpu().numpy(), g[f'c{k}_partial'])
        np.testing.assert_array_equal(out['total'].cpu().numpy(), g[f'c{k}_total'])
        np.testing.assert_array_equal(out['state_next'].cpu().numpy(), g[f'c{k}_snext'])
        np.testing.assert_array_equal(out['done'].cpu().numpy(), g[f'c{k}_done'])
        np.testing.assert_array_equal(out['term'].cpu().numpy(), g[f'c{k}_term'])
        np.testing.assert_array_equal(out['rwrd'].cpu().numpy(), g[f'c{k}_rwrd'])
        assert float(out['ep_return'][0]) == float(g[f'c{k}_ret'])


def test_ragged_batch_matches_oracle():
    """EP_UPDATE = 200 trajectories of ragged length (manipulator sizes: T <= 100, n = 50)."""
    from cacto_b200.rtg import rtg_batch
    rng = np.random.default_rng(0)
    conf = SimpleNamespace(nb_state=7, MC=0, nsteps_TD_N=50)
    lens = rng.integers(1, 101, 200)
    lens[0], lens[1] = 100, 1
    states = [rng.normal(size=(T + 1, 7)) for T in lens]
    costs = [rng.uniform(0, 2, T + 1) * 10.0 ** rng.integers(-5, 1, T + 1) for T in lens]
    out = rtg_batch(conf, states, costs)
    off = out['offsets']
    for e in range(200):
        _, partial, total, snext, done, rwrd, term, ret = ortg.rl_solve(conf, states[e], costs[e])
        sl = slice(off[e], off[e + 1])
        np.testing.assert_array_equal(out['partial'][sl].cpu().numpy(), partial)
        np.testing.assert_array_equal(out['total'][sl].cpu().numpy(), total)
        np.testing.assert_array_equal(out['state_next'][sl].cpu().numpy(), snext)
        np.testing.assert_array_equal(out['done'][sl].cpu().numpy(), done)
        np.testing.assert_array_equal(out['term'][sl].cpu().numpy(), term)
        assert float(out['ep_return'][e]) == ret


def test_rl_solve_env_rl_resimulates_the_controls():
    """env_RL = 1 (RL.py:159-166): states and rewards come from Env.step on the TO controls, then the same windows."""
    import copy
    from cacto_b200 import environment as genv
    from cacto_b200.NeuralNetwork import NN
    from cacto_b200.RL import RL_AC
    from cacto_b200.conf import get_conf
    from oracle import systems as osys
    conf = copy.deepcopy(get_conf('car'))
    conf.env_RL = 1
    env, oenv = genv.make_env(conf), osys.make_env(conf)
    rl = RL_AC(env, NN(env, conf, 1e-2, seed=0), conf, 0)
    rl.setup_model()
    rng = np.random.default_rng(4)
    x0 = rng.uniform(conf.x_init_min, conf.x_init_max)
    x0[-1] = (conf.NSTEPS - 12) * conf.dt
    assert rl.create_TO_init(0, x0)[-1] == 1 and rl.NSTEPS_SH == 12
    U = rng.uniform(conf.u_min, conf.u_max, (12, conf.nb_action))
    got = rl.RL_Solve(U, None, None)
    st = np.zeros((13, conf.nb_state)); st[0] = x0
    rw = np.zeros(13)
    for k in range(12):
        st[k + 1], rw[k] = oenv.step(conf.cost_weights_running, st[k], U[k])
    rw[-1] = oenv.reward(conf.cost_weights_terminal, st[-1])
    np.testing.assert_allclose(got[0], st, rtol=1e-9, atol=1e-12)
    np.testing.assert_allclose(got[5], rw, rtol=1e-9, atol=1e-12)
    _, partial, total, snext, done, rwrd, term, ret = ortg.rl_solve(conf, got[0], -got[5])
    np.testing.assert_array_equal(got[1], partial)
    np.testing.assert_array_equal(got[2], total)
    np.testing.assert_array_equal(got[3], snext)
    np.testing.assert_array_equal(got[4], done)


@pytest.mark.parametrize('mc', [0, 1])
def test_long_trajectories_and_mc(mc):
    """Car-sized trajectories (501 knots, n = 125: the shared-memory staging limit of 512 knots), one beyond it (rewards read from
    global memory), windows longer than the trajectory, and MC mode (window = whole trajectory, nothing copied into s_next)."""
    from cacto_b200.rtg import rtg_batch
    rng = np.random.default_rng(3)
    conf = SimpleNamespace(nb_state=6, MC=mc, nsteps_TD_N=125)
    lens = [500, 511, 512, 700, 3, 125, 126]
    states = [rng.normal(size=(T + 1, 6)) for T in lens]
    costs = [rng.uniform(0, 2, T + 1) * 10.0 ** rng.integers(-5, 1, T + 1) for T in lens]
    out = rtg_batch(conf, states, costs)
    off = out['offsets']
    for e in range(len(lens)):
        _, partial, total, snext, done, rwrd, term, ret = ortg.rl_solve(conf, states[e], costs[e])
        sl = slice(off[e], off[e + 1])
        np.testing.assert_array_equal(out['partial'][sl].cpu().numpy(), partial)
        np.testing.assert_array_equal(out['total'][sl].cpu().numpy(), total)
        np.testing.assert_array_equal(out['state_next'][sl].cpu().numpy(), snext)
        np.testing.assert_array_equal(out['done'][sl].cpu().numpy(), done)
        assert float(out['ep_return'][e]) == ret
