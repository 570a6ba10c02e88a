"""Oracle of TO_Casadi.backward_pass (oracle/backward.py): the hyper-dual reward equals the pinned oracle reward,
its derivatives equal finite differences, and the recursion is self-consistent (CasADi absent: parity unpinned)."""
import numpy as np
import pytest

from cacto_b200.conf import SYSTEM_IDS, get_conf
from oracle import backward as obw
from oracle import systems as osys


def sample(conf, rng):
    x = rng.uniform(np.asarray(conf.x_init_min[:-1], float), np.asarray(conf.x_init_max[:-1], float))
    u = rng.uniform(np.asarray(conf.u_min, float), np.asarray(conf.u_max, float)) * 0.5
    return x, u


@pytest.mark.parametrize('system', SYSTEM_IDS)
def test_generic_reward_equals_env_reward(system):
    conf = get_conf(system)
    env = osys.make_env(conf)
    rng = np.random.default_rng(0)
    for w in (conf.cost_weights_running, conf.cost_weights_terminal):
        for _ in range(4):
            x, u = sample(conf, rng)
            r = obw.reward_generic(env, w, list(x), None)
            assert r == pytest.approx(env.reward(w, np.append(x, 0.0)), rel=1e-12, abs=1e-12)
            if system != 'ur5':          # UR5.reward uses the plain u.u control cost (quirk Q8); the TO cost is bounded
                r = obw.reward_generic(env, w, list(x), list(u))
                assert r == pytest.approx(env.reward(w, np.append(x, 0.0), u), rel=1e-12, abs=1e-12)


@pytest.mark.parametrize('system', SYSTEM_IDS)
def test_hyperdual_derivatives_match_finite_differences(system):
    conf = get_conf(system)
    env = osys.make_env(conf)
    rng = np.random.default_rng(1)
    x, u = sample(conf, rng)
    w = conf.cost_weights_running
    g, H = obw.reward_x_derivatives(env, w, x)
    f = lambda z: obw.reward_generic(env, w, list(z))
    h = 1e-5
    n = len(x)
    g_fd = np.array([(f(x + h * np.eye(n)[i]) - f(x - h * np.eye(n)[i])) / (2 * h) for i in range(n)])
    sc = max(1.0, np.abs(g).max())
    assert np.abs(g - g_fd).max() <= 1e-6 * sc
    H_fd = np.zeros((n, n))
    hh = 1e-4
    for i in range(n):
        for j in range(n):
            ei, ej = hh * np.eye(n)[i], hh * np.eye(n)[j]
            H_fd[i, j] = (f(x + ei + ej) - f(x + ei - ej) - f(x - ei + ej) + f(x - ei - ej)) / (4 * hh * hh)
    assert np.abs(H - H_fd).max() <= 1e-4 * max(1.0, np.abs(H).max())
    assert np.allclose(H, H.T)
    gu, Hu = obw.reward_u_derivatives(env, w, u)
    fu = lambda a: obw.reward_generic(env, w, list(x), list(a))
    m = len(u)
    gu_fd = np.array([(fu(u + h * np.eye(m)[i]) - fu(u - h * np.eye(m)[i])) / (2 * h) for i in range(m)])
    assert np.abs(gu - gu_fd).max() <= 1e-6 * max(1.0, np.abs(gu).max())


def test_backward_pass_equals_gradient_of_the_lq_value_for_the_double_integrator():
    """Linear dynamics: V_x[0] of the recursion must equal d/dx0 of the quadratic model's optimal reward-to-go; checked on
    a 3-knot problem by brute-force maximisation over the controls of the second-order model around the trajectory."""
    conf = get_conf('double_integrator')
    env = osys.make_env(conf)
    rng = np.random.default_rng(2)
    T = 6
    X = np.zeros((T, 4)); U = rng.uniform(-1, 1, (T - 1, 2))
    X[0] = sample(conf, rng)[0]
    for t in range(T - 1):
        X[t + 1] = env.simulate(np.append(X[t], 0.0), U[t])[:-1]
    Vx = obw.backward_pass(env, T, X, U)
    assert Vx.shape == (T, 5) and np.all(Vx[:, -1] == 0)
    g_T, _ = obw.reward_x_derivatives(env, conf.cost_weights_terminal, X[-1])
    np.testing.assert_allclose(Vx[-1, :-1], g_T)
    # one step back by hand
    A, B = env.augmented_derivative(np.append(X[-2], 0.0), U[-1])
    l_x, l_xx = obw.reward_x_derivatives(env, conf.cost_weights_running, X[-2])
    l_u, l_uu = obw.reward_u_derivatives(env, conf.cost_weights_running, U[-1])
    _, V_xx = obw.reward_x_derivatives(env, conf.cost_weights_terminal, X[-1])
    Quu = l_uu + B.T @ V_xx @ B + 1e-9 * np.eye(2)
    ref = (l_x + A.T @ g_T) - (A.T @ V_xx @ B) @ np.linalg.solve(Quu, l_u + B.T @ g_T)
    np.testing.assert_allclose(Vx[-2, :-1], ref, rtol=1e-9, atol=1e-12)


@pytest.mark.parametrize('system', ['single_integrator', 'car', 'car_park'])
def test_backward_pass_matches_the_reference_run_on_the_casadi_stub(system):
    """tests/golden/bp_cases.npz: the reference's own TO_Casadi.backward_pass + *_CAMS cost models + Env.augmented_derivative,
    executed unmodified on tests/golden/_casadi_stub.py (exact hyper-dual derivatives instead of CasADi's symbolic ones)."""
    from conftest import golden
    g = golden('bp_cases.npz')
    conf = get_conf(system)
    env = osys.make_env(conf)
    for k in range(3):
        X, U, ref = g[f'{system}_{k}_X'], g[f'{system}_{k}_U'], g[f'{system}_{k}_Vx']
        got = obw.backward_pass(env, len(X), X, U)
        np.testing.assert_allclose(got, ref, rtol=1e-9, atol=1e-11)
        # the reference's cost model is the negative reward (running weights, bounded control cost)
        for t in range(len(X)):
            u = U[min(t, len(X) - 2)] if len(U) else np.zeros(conf.nb_action)
            assert -g[f'{system}_{k}_cost'][t] == pytest.approx(obw.reward_generic(env, conf.cost_weights_running, list(X[t]), list(u)), rel=1e-12, abs=1e-12)
