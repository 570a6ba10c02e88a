"""Consumers of the PINNING KIT (tests/golden/make_golden_ext.py): golden vectors produced by the UNMODIFIED reference with its
real third-party stack (pinocchio, tensorflow 2.11, tf_siren) -- the arithmetic that cannot be pinned in the build container.
The files are picked up when present (tests/golden/ext_env_<system>.npz, ext_nn_<system>.npz); until someone with the wheels runs
the kit these tests skip and SURVEY.md rows A3 / A6 / A7 / A11 / N2-N8 stay "parity unpinned".
CPU half: the oracle against the vectors.  GPU half (-m gpu): the CUDA kernels against the vectors at BASELINE.md's gates."""
import os

import numpy as np
import pytest

from cacto_b200.conf import get_conf
from conftest import GOLDEN

SYSTEMS = ('single_integrator', 'double_integrator', 'car', 'manipulator', 'ur5')


def _load(kind, system):
    path = os.path.join(GOLDEN, f'ext_{kind}_{system}.npz')
    if not os.path.exists(path):
        pytest.skip(f'{os.path.basename(path)} not generated yet (run tests/golden/make_golden_ext.py where tensorflow / pinocchio exist)')
    return np.load(path, allow_pickle=False)


def _rel(a, b):
    a, b = np.asarray(a, dtype=float), np.asarray(b, dtype=float)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


@pytest.mark.parametrize('system', ('double_integrator', 'manipulator', 'ur5'))
def test_oracle_dynamics_match_pinocchio_goldens(system):
    g = _load('env', system)
    from oracle import systems as osys
    conf = get_conf(system)
    env = osys.make_env(conf)
    for i in range(len(g['state'])):
        s, a = g['state'][i], g['action'][i]
        assert _rel(env.simulate(s, a), g['next'][i]) < 1e-9
        assert _rel(env.derivative(s, a), g['Fu_norm'][i]) < 1e-9
        Fx, Fu = env.augmented_derivative(s, a)
        assert _rel(Fx, g['Fx'][i]) < 1e-8 and _rel(Fu, g['Fu'][i]) < 1e-9
        assert _rel(env.get_end_effector_position(s), g['ee'][i]) < 1e-10
        assert abs(env.reward(conf.cost_weights_running, s, a) - g['reward_run'][i]) <= 1e-9 * max(1.0, abs(g['reward_run'][i]))
        assert abs(env.reward(conf.cost_weights_terminal, s) - g['reward_ter'][i]) <= 1e-9 * max(1.0, abs(g['reward_ter'][i]))


def _case(g, case):
    P = case + '/'
    get = lambda name, n: [g[f'{P}{name}{i}'] for i in range(n)]
    batch = tuple(g[P + k] for k in ('s', 'pr', 'sn', 'dv', 'd', 'term', 'w'))
    return P, get, batch


@pytest.mark.parametrize('system', SYSTEMS)
@pytest.mark.parametrize('case,w_S,mc', [('sobolev', 1e-2, 0), ('value_only', 0.0, 0), ('mc', 1e-2, 1)])
def test_oracle_update_matches_tensorflow_goldens(system, case, w_S, mc):
    g = _load('nn', system)
    import torch
    from oracle import nn as onn, systems as osys
    conf = get_conf(system, MC=mc)
    P, get, (s, pr, sn, dv, d, term, w) = _case(g, case)
    actor, critic, target = get('actor_w', 6), get('critic_w', 10), get('target_w', 10)
    st = torch.tensor(s)
    assert _rel(onn.actor_forward(onn.to_torch(actor), st, conf).numpy(), g[P + 'actor_out']) < 2e-5
    assert _rel(onn.critic_forward(onn.to_torch(critic), st, conf).numpy(), g[P + 'critic_out']) < 2e-5
    cg, rtg, V, Vt, _ = onn.critic_grad(critic, target, conf, w_S, s, sn, pr, dv, d, w)
    assert _rel(rtg, g[P + 'rtg']) < 2e-5 and _rel(V, g[P + 'V']) < 2e-5 and _rel(Vt, g[P + 'Vt']) < 2e-5
    for i, x in enumerate(cg):
        assert _rel(x, g[f'{P}critic_grad{i}']) < 1e-4
    env = osys.make_env(conf)
    ag = onn.actor_grad(actor, critic, conf, env, s, term)[0]
    for i, x in enumerate(ag):
        assert _rel(x, g[f'{P}actor_grad{i}']) < 1e-4
    oc, oa = onn.Adam(critic, conf.CRITIC_LEARNING_RATE), onn.Adam(actor, conf.ACTOR_LEARNING_RATE)
    for k in (1, 2):
        onn.update(critic, target, actor, oc, oa, conf, w_S, env, (s, pr, sn, dv, d, term, w))
        for name, arrs in (('critic', critic), ('actor', actor), ('target', target)):
            for i, x in enumerate(arrs):
                assert _rel(x, g[f'{P}step{k}_{name}_w{i}']) < 1e-4, (k, name, i)


@pytest.mark.gpu
@pytest.mark.parametrize('system', ('double_integrator', 'manipulator', 'ur5'))
def test_kernels_match_pinocchio_goldens(system):
    g = _load('env', system)
    import torch
    from cacto_b200 import environment as genv
    conf = get_conf(system)
    env = genv.make_env(conf)
    s, a = torch.tensor(g['state'], device='cuda'), torch.tensor(g['action'], device='cuda')
    assert _rel(env.simulate_batch(s, a).cpu().numpy(), g['next']) < 1e-6
    assert _rel(env.derivative_batch(s, a).cpu().numpy(), g['Fu_norm']) < 1e-6
    Fx, Fu = env.augmented_derivative_batch(s, a)
    assert _rel(Fx.cpu().numpy(), g['Fx']) < 1e-6 and _rel(Fu.cpu().numpy(), g['Fu']) < 1e-6
    s32, a32 = s.float(), a.float()
    assert _rel(env.simulate_batch(s32, a32).cpu().numpy(), g['next_batch_f32']) < 1e-5
    assert _rel(env.derivative_batch(s32, a32).cpu().numpy(), g['Fu_batch_f32']) < 1e-5


@pytest.mark.gpu
@pytest.mark.parametrize('system', SYSTEMS)
@pytest.mark.parametrize('case,w_S,mc', [('sobolev', 1e-2, 0), ('value_only', 0.0, 0), ('mc', 1e-2, 1)])
def test_kernels_match_tensorflow_goldens(system, case, w_S, mc):
    g = _load('nn', system)
    import torch
    from cacto_b200 import environment as genv
    from cacto_b200.NeuralNetwork import NN
    from cacto_b200.RL import RL_AC
    conf = get_conf(system, MC=mc)
    env = genv.make_env(conf)
    nn = NN(env, conf, w_S, seed=0)
    rl = RL_AC(env, nn, conf, 0)
    rl.setup_model()
    P, get, (s, pr, sn, dv, d, term, w) = _case(g, case)
    rl.actor_model.set_weights(get('actor_w', 6)); rl.critic_model.set_weights(get('critic_w', 10)); rl.target_critic.set_weights(get('target_w', 10))
    assert _rel(nn.eval(rl.actor_model, s).cpu().numpy(), g[P + 'actor_out']) < 2e-5
    assert _rel(nn.eval(rl.critic_model, s).cpu().numpy(), g[P + 'critic_out']) < 2e-5
    cg, rtg, V, Vt = nn.compute_critic_grad(rl.critic_model, rl.target_critic, s, sn, pr, dv, d, w)
    assert _rel(rtg.cpu().numpy(), g[P + 'rtg']) < 2e-5 and _rel(V.cpu().numpy(), g[P + 'V']) < 2e-5
    for i, x in enumerate(cg):
        assert _rel(x.cpu().numpy(), g[f'{P}critic_grad{i}']) < 1e-4
    ag = nn.compute_actor_grad(rl.actor_model, rl.critic_model, s, term, None)
    for i, x in enumerate(ag):
        assert _rel(x.cpu().numpy(), g[f'{P}actor_grad{i}']) < 1e-4
    for k in (1, 2):
        rl.update(s, sn, pr, dv, d, term, w, fuse_target=True)
        for name, net in (('critic', rl.critic_model), ('actor', rl.actor_model), ('target', rl.target_critic)):
            for i, x in enumerate(net.get_weights()):
                assert _rel(x, g[f'{P}step{k}_{name}_w{i}']) < 1e-4, (k, name, i)
