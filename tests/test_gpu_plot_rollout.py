"""PLOT.rollout (plot_utils.py:245-279): evaluation rollouts with rewards -- the second caller of the fused rollout kernel."""
import numpy as np
import pytest
import torch

from cacto_b200.conf import get_conf
from oracle import nn as onn
from oracle import systems as osys

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize('system', ['manipulator', 'car'])
def test_plot_rollout_matches_the_reference_loop(system, capsys):
    from cacto_b200 import environment as genv
    from cacto_b200.NeuralNetwork import NN
    from cacto_b200.RL import RL_AC
    from cacto_b200.plot_utils import PLOT
    conf = get_conf(system)
    env = genv.make_env(conf)
    nn = NN(env, conf, 1e-2, seed=3)
    rl = RL_AC(env, nn, conf, 0)
    rl.setup_model()
    rng = np.random.default_rng(4)
    init = rng.uniform(np.asarray(conf.x_init_min, float), np.asarray(conf.x_init_max, float), (4, conf.nb_state))
    init[:, -1] = 0.0
    plot = PLOT(7, env, nn, conf)
    returns = plot.rollout(123, rl.actor_model, init)
    assert 'N try = 7: Simulation Return @ N updates = 123 ==> ' in capsys.readouterr().out
    # the reference's loop (plot_utils.py:252-275) on the oracle
    oenv = osys.make_env(conf)
    ap = onn.to_torch(rl.actor_model.get_weights())
    T = conf.NSTEPS
    assert list(returns.keys()) == [(init[k][0], init[k][1]) for k in range(4)] and len(plot.p_ee_all_sim) == 4
    for k in range(4):
        x = init[k].copy()
        p_ee = np.zeros((T + 1, 3))
        p_ee[0] = oenv.get_end_effector_position(x)
        total = 0
        for i in range(T):
            with torch.no_grad():
                u = onn.actor_forward(ap, torch.tensor(x[None], dtype=torch.float32), conf).numpy()[0].astype(np.float64)
            x, r = oenv.step(conf.cost_weights_running, x, u)
            p_ee[i + 1] = oenv.get_end_effector_position(x)
            p_ee[i + 1, -1] = x[2]
            total += r
        np.testing.assert_allclose(plot.rollout_states[k][-1], x, rtol=2e-4, atol=2e-5)
        np.testing.assert_allclose(plot.p_ee_all_sim[k], p_ee, rtol=2e-4, atol=2e-5)
        assert abs(returns[init[k][0], init[k][1]] - total) <= 2e-4 * abs(total) + 1e-6
