"""K6 -- TO_Casadi.backward_pass on the GPU (cacto_backward_pass) against the oracle's restatement of TO.py:119-202."""
import numpy as np
import pytest
import torch

from cacto_b200.conf import SYSTEM_IDS, get_conf
from oracle import backward as obw
from oracle import systems as osys

pytestmark = pytest.mark.gpu


def trajectory(conf, oenv, rng, T):
    x = rng.uniform(np.asarray(conf.x_init_min[:-1], float), np.asarray(conf.x_init_max[:-1], float))
    X, U = [x], []
    for _ in range(T):
        u = rng.uniform(np.asarray(conf.u_min, float), np.asarray(conf.u_max, float)) * 0.3
        U.append(u)
        X.append(oenv.simulate(np.append(X[-1], 0.0), u)[:-1])
    return np.array(X), np.array(U).reshape(-1, conf.nb_action)


@pytest.mark.parametrize('system', SYSTEM_IDS)
def test_backward_pass_matches_oracle(system):
    from cacto_b200 import environment as genv
    from cacto_b200.TO import TO_Casadi
    conf = get_conf(system)
    env, oenv = genv.make_env(conf), osys.make_env(conf)
    to = TO_Casadi(env, conf, None, w_S=1e-2)
    rng = np.random.default_rng(5)
    lens = [7, 1, 12] if system != 'ur5' else [5, 3]                # ragged batch, incl. a single-knot trajectory
    trajs = [trajectory(conf, oenv, rng, T - 1) for T in lens]
    Vx, off = to.backward_pass_batch([t[0] for t in trajs], [t[1] for t in trajs])
    Vx = Vx.cpu().numpy()
    assert Vx.shape == (sum(lens), conf.nb_state) and np.all(Vx[:, -1] == 0)
    for e, (X, U) in enumerate(trajs):
        ref = obw.backward_pass(oenv, len(X), X, U)
        got = Vx[off[e]:off[e + 1]]
        sc = np.abs(ref).max(axis=0) + 1e-9
        assert (np.abs(got - ref) / sc).max() < 1e-6, (system, e)
    # the reference-signature call
    X, U = trajs[0]
    one = to.backward_pass(len(X), np.hstack([X, np.zeros((len(X), 1))]), U)
    np.testing.assert_allclose(one, Vx[off[0]:off[1]], rtol=0, atol=0)


def test_backward_pass_pinv_handles_zero_control_weight():
    """w_u = 0 at a rest state of the single integrator makes Q_uu = B' V_xx B + mu I only: pinv path, still matches."""
    from cacto_b200 import environment as genv
    from cacto_b200.TO import TO_Casadi
    import copy
    conf = copy.deepcopy(get_conf('single_integrator'))
    conf.cost_weights_running = np.array(conf.cost_weights_running, dtype=float)
    conf.cost_weights_running[6] = 0.0
    env, oenv = genv.make_env(conf), osys.make_env(conf)
    rng = np.random.default_rng(9)
    X, U = trajectory(conf, oenv, rng, 9)
    got = TO_Casadi(env, conf).backward_pass(len(X), X, U)
    ref = obw.backward_pass(oenv, len(X), X, U)
    sc = np.abs(ref).max(axis=0) + 1e-9
    assert (np.abs(got - ref) / sc).max() < 1e-6


@pytest.mark.parametrize('system', ['single_integrator', 'car', 'car_park'])
def test_backward_pass_matches_reference_goldens(system):
    """The reference's own backward_pass (tests/golden/bp_cases.npz, see make_golden.backward_pass_goldens)."""
    from conftest import golden
    from cacto_b200 import environment as genv
    from cacto_b200.TO import TO_Casadi
    g = golden('bp_cases.npz')
    conf = get_conf(system)
    to = TO_Casadi(genv.make_env(conf), conf, None, w_S=1e-2)
    Xs = [g[f'{system}_{k}_X'] for k in range(3)]
    Us = [g[f'{system}_{k}_U'] for k in range(3)]
    Vx, off = to.backward_pass_batch(Xs, Us)
    Vx = Vx.cpu().numpy()
    for k in range(3):
        ref = g[f'{system}_{k}_Vx']
        sc = np.abs(ref).max(axis=0) + 1e-9
        assert (np.abs(Vx[off[k]:off[k + 1]] - ref) / sc).max() < 1e-6
