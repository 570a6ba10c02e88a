"""oracle.systems pinned against the reference's environment.py run in the build container
(tests/golden/env_<system>.npz, made by tests/golden/make_golden.py)."""
import numpy as np
import pytest

from cacto_b200.conf import SYSTEM_IDS, get_conf
from conftest import golden
from oracle import systems

ANALYTIC = ('single_integrator', 'car', 'car_park')


@pytest.mark.parametrize('system', SYSTEM_IDS)
def test_rewards_match_reference(system):
    g = golden(f'env_{system}.npz')
    env = systems.make_env(get_conf(system))
    S, A, W = g['states'], g['actions'], g['weights']
    if 'ee_injected' in g:      # FK was injected into the reference from the oracle (Pinocchio absent)
        np.testing.assert_array_equal(np.array([env.get_end_effector_position(s) for s in S]), g['ee_injected'])
    r_sa = np.array([env.reward(w, s, a) for w, s, a in zip(W, S, A)])
    r_s = np.array([env.reward(w, s) for w, s in zip(W, S)])
    np.testing.assert_allclose(r_sa, g['reward_sa'], rtol=1e-13, atol=0)
    np.testing.assert_allclose(r_s, g['reward_s'], rtol=1e-13, atol=0)
    rb = env.reward_batch(W, S.astype(np.float32), A.astype(np.float32))
    np.testing.assert_allclose(rb, g['reward_batch'], rtol=2e-6, atol=0)


@pytest.mark.parametrize('system', ANALYTIC)
def test_analytic_dynamics_match_reference(system):
    g = golden(f'env_{system}.npz')
    env = systems.make_env(get_conf(system))
    S, A = g['states'], g['actions']
    np.testing.assert_array_equal(np.array([env.simulate(s, a) for s, a in zip(S, A)]), g['simulate'])
    np.testing.assert_array_equal(np.array([env.derivative(s, a) for s, a in zip(S, A)]), g['derivative'])
    fx, fu = zip(*[env.augmented_derivative(s, a) for s, a in zip(S, A)])
    np.testing.assert_allclose(np.array(fx), g['Fx'], rtol=1e-15, atol=0)
    np.testing.assert_array_equal(np.array(fu), g['Fu'])
    np.testing.assert_array_equal(np.array([env.get_end_effector_position(s) for s in S]), g['ee'])
    s32, a32 = S.astype(np.float32), A.astype(np.float32)
    np.testing.assert_array_equal(env.simulate_batch(s32, a32), g['simulate_batch'])
    np.testing.assert_array_equal(env.derivative_batch(s32, a32), g['derivative_batch'])


def test_reward_batch_da_matches_finite_difference():
    conf = get_conf('manipulator')
    env = systems.make_env(conf)
    rng = np.random.default_rng(0)
    a = rng.uniform(-150, 150, (5, 3))
    w = np.tile(conf.cost_weights_running, (5, 1))
    g = env.reward_batch_da(w, a)
    s = rng.uniform(-1, 1, (5, 7))
    for i in range(5):
        for j in range(3):
            e = np.zeros(3)
            e[j] = 1e-4
            fd = (env.reward(w[i], s[i], a[i] + e) - env.reward(w[i], s[i], a[i] - e)) / 2e-4
            assert abs(fd - g[i, j]) <= 1e-6 * max(1e-8, abs(fd))
