"""CUDA actor/critic kernels (forward, Sobolev critic gradient, actor gradient, Adam, Polyak) against the
oracle (torch-CPU autograd restatement of NeuralNetwork.py / RL.py).  BASELINE.md gate: losses and updated
weights after one step within 1e-4 relative."""
import numpy as np
import pytest
import torch

from cacto_b200.conf import get_conf
from conftest import golden
from oracle import nn as onn
from oracle import systems as osys

pytestmark = pytest.mark.gpu


def rel(a, b):
    a = a.detach().cpu().double().numpy() if isinstance(a, torch.Tensor) else np.asarray(a, dtype=float)
    b = np.asarray(b, dtype=float)
    assert a.shape == b.shape, (a.shape, b.shape)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


def make(system, B, seed=0, w_S=1e-2, **over):
    from cacto_b200 import environment as genv
    from cacto_b200.NeuralNetwork import NN
    from cacto_b200.RL import RL_AC
    conf = get_conf(system, **over)
    env = genv.make_env(conf)
    nn = NN(env, conf, w_S, seed=seed)
    rl = RL_AC(env, nn, conf, 0)
    rl.setup_model()
    rng = np.random.default_rng(seed + 1)
    ns = conf.nb_state
    lo, hi = np.asarray(conf.x_init_min, float), np.asarray(conf.x_init_max, float)
    s = rng.uniform(lo, hi, (B, ns)).astype(np.float32)
    sn = rng.uniform(lo, hi, (B, ns)).astype(np.float32)
    if system == 'car_park':
        for a in (s, sn):
            a[:, 3] = rng.uniform(-3, 3, B)
            a[:, 4] = rng.uniform(-0.5, 0.5, B)
    pr = rng.uniform(-5, 0, (B, 1)).astype(np.float32)
    dv = rng.normal(size=(B, ns)).astype(np.float32)
    dv[:, -1] = 0
    dv[0, 0] = 0.0                      # exercises the slog gradient gate at exactly 0 (quirk Q10)
    d = (rng.uniform(size=(B, 1)) < 0.5).astype(np.float32)
    term = (rng.uniform(size=(B, 1)) < 0.2).astype(np.float64)
    w = rng.uniform(0.2, 1.5, (B, 1)).astype(np.float32)
    return conf, env, nn, rl, (s, pr, sn, dv, d, term, w)


SYSTEMS_B = [('single_integrator', 128), ('double_integrator', 37), ('car', 64), ('car_park', 64), ('manipulator', 64), ('ur5', 19),
             ('manipulator', 2500)]


@pytest.mark.parametrize('system,B', SYSTEMS_B)
def test_forward_matches_oracle(system, B):
    conf, env, nn, rl, batch = make(system, B)
    s = batch[0]
    ap, cp = onn.to_torch(rl.actor_model.get_weights()), onn.to_torch(rl.critic_model.get_weights())
    st = torch.tensor(s, requires_grad=True)
    a_ref = onn.actor_forward(ap, st, conf)
    v_ref = onn.critic_forward(cp, st, conf)
    g_ref, = torch.autograd.grad(v_ref.sum(), st)
    assert rel(nn.eval(rl.actor_model, s), a_ref.detach().numpy()) < 2e-5
    assert rel(nn.eval(rl.critic_model, s), v_ref.detach().numpy()) < 2e-5
    V, dV = nn.eval_with_gradient(rl.critic_model, s)
    assert rel(V, v_ref.detach().numpy()) < 2e-5
    assert rel(dV, g_ref.numpy()) < 5e-5


@pytest.mark.parametrize('system,B', SYSTEMS_B)
@pytest.mark.parametrize('w_S', [1e-2, 0.0])
def test_critic_and_actor_gradients_match_oracle(system, B, w_S):
    conf, env, nn, rl, batch = make(system, B, w_S=w_S)
    s, pr, sn, dv, d, term, w = batch
    critic, target, actor = rl.critic_model.get_weights(), rl.target_critic.get_weights(), rl.actor_model.get_weights()
    target = [t + 0.01 * np.random.default_rng(5).normal(size=t.shape).astype(np.float32) for t in target]
    rl.target_critic.set_weights(target)
    cg, rtg, V, Vt, loss = onn.critic_grad(critic, target, conf, w_S, s, sn, pr, dv, d, w)
    g, g_rtg, g_V, g_Vt = nn.compute_critic_grad(rl.critic_model, rl.target_critic, s, sn, pr, dv, d, w)
    assert rel(g_rtg, rtg) < 2e-5 and rel(g_V, V) < 2e-5 and rel(g_Vt, Vt) < 2e-5
    assert abs(float(nn.last_critic_loss) - loss) <= 1e-4 * abs(loss)
    for gv, rv in zip(g, cg):
        assert rel(gv, rv) < 1e-4
    oenv = osys.make_env(conf)
    ag, actions, s_next, dQ = onn.actor_grad(actor, critic, conf, oenv, s, term)
    ga, act = nn.compute_actor_grad(rl.actor_model, rl.critic_model, s, term, None, return_actions=True)
    assert rel(act, actions) < 2e-5
    for gv, rv in zip(ga, ag):
        assert rel(gv, rv) < 1e-4


@pytest.mark.parametrize('system,B', [('manipulator', 64), ('double_integrator', 128), ('ur5', 16), ('car', 2400)])
def test_one_update_step_matches_oracle(system, B):
    """RL_AC.update + update_target: weights of critic, actor and target after one (and two) steps."""
    conf, env, nn, rl, batch = make(system, B)
    s, pr, sn, dv, d, term, w = batch
    critic, target, actor = rl.critic_model.get_weights(), rl.target_critic.get_weights(), rl.actor_model.get_weights()
    before = [x.copy() for x in critic + actor]
    oc, oa = onn.Adam(critic, conf.CRITIC_LEARNING_RATE), onn.Adam(actor, conf.ACTOR_LEARNING_RATE)
    oenv = osys.make_env(conf)
    for step in range(2):
        out = onn.update(critic, target, actor, oc, oa, conf, 1e-2, oenv, (s, pr, sn, dv, d, term, w))
        rtg, V, Vt = rl.update(s, sn, pr, dv, d, term, w)
        rl.update_target(rl.target_critic.variables, rl.critic_model.variables)
        assert rel(rtg, out['rtg']) < 2e-5 and rel(V, out['V']) < 2e-5
        mine = rl.critic_model.get_weights() + rl.actor_model.get_weights()
        for k, (m, r, b0) in enumerate(zip(mine, critic + actor, before)):
            assert rel(m, r) < 1e-4, (step, k)
            # the step itself (|dw| ~ lr): compare against the oracle's step with a tolerance relative to lr
            lr = conf.CRITIC_LEARNING_RATE if k < 10 else conf.ACTOR_LEARNING_RATE
            assert np.abs((m - b0) - (r - b0)).max() < 0.02 * lr * (step + 1), (step, k)
        for m, r in zip(rl.target_critic.get_weights(), target):
            assert rel(m, r) < 1e-5


# ---- BASELINE configs 2, 3, 5 at their full update batches (PER batch 4096, critic batch 16 384, UR5 shard) against the oracle.
# The oracle's per-sample dynamics loops are fanned over worker processes (oracle.systems.PooledEnv: same arithmetic).
LARGE = [('manipulator', 4096), ('car', 16384), ('ur5', 4096), ('double_integrator', 4096)]


def _close_in_fp32_class(got, ref32, ref64, gate=1e-4):
    """|got - fp64 oracle| within the gate plus what the oracle's own fp32 evaluation loses on this input: at these batch sizes
    a gradient is a sum of thousands of per-sample terms of both signs through LeakyReLU kinks, and e.g. the double integrator's
    actor gradient differs by 1.4e-3 (dW2) between the fp32 and the fp64 evaluation of the SAME oracle (a hidden unit whose
    pre-activation changes sign under a 1e-6 perturbation flips its slope).  An implementation is accepted when it is as close
    to the fp64 value as the reference's own precision class is; where the oracle is well conditioned (cond ~ 1e-7: every other
    system here) this IS the 1e-4 gate."""
    for g, r32, r64 in zip(got, ref32, ref64):
        cond = rel(torch.as_tensor(r32), r64)
        assert rel(g, r64) <= gate + 1.5 * cond, (rel(g, r64), cond)
        assert rel(g, r32) <= gate + 2.5 * cond, (rel(g, r32), cond)


@pytest.mark.parametrize('system,B', LARGE)
def test_large_batch_gradients_match_oracle(system, B):
    conf, env, nn, rl, batch = make(system, B)
    s, pr, sn, dv, d, term, w = batch
    critic, target, actor = rl.critic_model.get_weights(), rl.target_critic.get_weights(), rl.actor_model.get_weights()
    target = [t + 0.01 * np.random.default_rng(5).normal(size=t.shape).astype(np.float32) for t in target]
    rl.target_critic.set_weights(target)
    cg, rtg, V, Vt, loss = onn.critic_grad(critic, target, conf, 1e-2, s, sn, pr, dv, d, w)
    cg64 = onn.critic_grad(critic, target, conf, 1e-2, s, sn, pr, dv, d, w, dtype=torch.float64)[0]
    g, g_rtg, g_V, g_Vt = nn.compute_critic_grad(rl.critic_model, rl.target_critic, s, sn, pr, dv, d, w)
    assert rel(g_rtg, rtg) < 2e-5 and rel(g_V, V) < 2e-5 and rel(g_Vt, Vt) < 2e-5
    assert abs(float(nn.last_critic_loss) - loss) <= 1e-4 * abs(loss)
    _close_in_fp32_class(g, cg, cg64)
    oenv = osys.PooledEnv(osys.make_env(conf))
    ag, actions, s_next, dQ = onn.actor_grad(actor, critic, conf, oenv, s, term)
    ag64 = onn.actor_grad(actor, critic, conf, oenv, s, term, dtype=torch.float64)[0]
    ga, act = nn.compute_actor_grad(rl.actor_model, rl.critic_model, s, term, None, return_actions=True)
    assert rel(act, actions) < 2e-5
    _close_in_fp32_class(ga, ag, ag64)


@pytest.mark.parametrize('system,B', [('manipulator', 4096), ('car', 16384)])
def test_large_batch_update_step_matches_oracle(system, B):
    conf, env, nn, rl, batch = make(system, B)
    s, pr, sn, dv, d, term, w = batch
    critic, target, actor = rl.critic_model.get_weights(), rl.target_critic.get_weights(), rl.actor_model.get_weights()
    oc, oa = onn.Adam(critic, conf.CRITIC_LEARNING_RATE), onn.Adam(actor, conf.ACTOR_LEARNING_RATE)
    oenv = osys.PooledEnv(osys.make_env(conf))
    out = onn.update(critic, target, actor, oc, oa, conf, 1e-2, oenv, (s, pr, sn, dv, d, term, w))
    rtg, V, Vt = rl.update(s, sn, pr, dv, d, term, w, fuse_target=True)
    assert rel(rtg, out['rtg']) < 2e-5 and rel(V, out['V']) < 2e-5
    for m, r in zip(rl.critic_model.get_weights() + rl.actor_model.get_weights(), critic + actor):
        assert rel(m, r) < 1e-4
    for m, r in zip(rl.target_critic.get_weights(), target):
        assert rel(m, r) < 1e-5


# ---- conf.MC = 1 (NeuralNetwork.py:154-155: the target is the Monte-Carlo partial reward-to-go itself, no target critic, and
# RL.py:134 skips update_target)
@pytest.mark.parametrize('system,B', [('manipulator', 64), ('single_integrator', 128), ('car', 2400)])
@pytest.mark.parametrize('w_S', [1e-2, 0.0])
def test_mc_critic_gradient_matches_oracle(system, B, w_S):
    conf, env, nn, rl, batch = make(system, B, w_S=w_S, MC=1)
    assert conf.MC == 1
    s, pr, sn, dv, d, term, w = batch
    critic, target = rl.critic_model.get_weights(), rl.target_critic.get_weights()
    cg, rtg, V, Vt, loss = onn.critic_grad(critic, target, conf, w_S, s, sn, pr, dv, d, w)
    np.testing.assert_array_equal(rtg, pr)
    g, g_rtg, g_V, g_Vt = nn.compute_critic_grad(rl.critic_model, rl.target_critic, s, sn, pr, dv, d, w)
    assert (g_rtg.cpu().numpy() == pr).all()                # rtg IS the partial reward-to-go
    assert rel(g_V, V) < 2e-5 and rel(g_Vt, Vt) < 2e-5
    assert abs(float(nn.last_critic_loss) - loss) <= 1e-4 * abs(loss)
    for gv, rv in zip(g, cg):
        assert rel(gv, rv) < 1e-4


def test_mc_update_leaves_target_untouched():
    conf, env, nn, rl, batch = make('manipulator', 64, MC=1)
    s, pr, sn, dv, d, term, w = batch
    critic, target, actor = rl.critic_model.get_weights(), rl.target_critic.get_weights(), rl.actor_model.get_weights()
    t0 = [t.copy() for t in target]
    oc, oa = onn.Adam(critic, conf.CRITIC_LEARNING_RATE), onn.Adam(actor, conf.ACTOR_LEARNING_RATE)
    onn.update(critic, target, actor, oc, oa, conf, 1e-2, osys.make_env(conf), (s, pr, sn, dv, d, term, w))
    rl.update(s, sn, pr, dv, d, term, w, fuse_target=True)      # learn_and_update's call: the fused Polyak step must be skipped
    for m, r in zip(rl.critic_model.get_weights() + rl.actor_model.get_weights(), critic + actor):
        assert rel(m, r) < 1e-4
    for m, r in zip(rl.target_critic.get_weights(), t0):
        np.testing.assert_array_equal(m, r)
    g = rl.make_update_graph(64)                                # and in the captured update
    for k, v in zip(('state', 'state_next', 'partial_rtg', 'dVdx', 'done', 'term', 'weights'), (s, sn, pr, dv, d, term, w)):
        g.io[k].copy_(torch.as_tensor(v))
    g.replay()
    for m, r in zip(rl.target_critic.get_weights(), t0):
        np.testing.assert_array_equal(m, r)


def test_fused_target_update_equals_separate():
    conf, env, nn, rl, batch = make('manipulator', 64)
    s, pr, sn, dv, d, term, w = batch
    w0 = [rl.critic_model.get_weights(), rl.target_critic.get_weights(), rl.actor_model.get_weights()]
    rl.update(s, sn, pr, dv, d, term, w, fuse_target=True)
    fused = rl.target_critic.get_weights()
    conf2, env2, nn2, rl2, _ = make('manipulator', 64)
    rl2.critic_model.set_weights(w0[0]); rl2.target_critic.set_weights(w0[1]); rl2.actor_model.set_weights(w0[2])
    rl2.update(s, sn, pr, dv, d, term, w)
    rl2.update_target(rl2.target_critic.variables, rl2.critic_model.variables)
    for a, b in zip(fused, rl2.target_critic.get_weights()):
        assert rel(a, b) < 1e-6


def test_lr_schedule_and_adam_iterations():
    from cacto_b200.optim import PiecewiseConstantDecay
    conf = get_conf('manipulator')
    sch = PiecewiseConstantDecay(conf.boundaries_schedule_LR_C, conf.values_schedule_LR_C)
    for step in (0, 1, 204800, 204801, 307200, 307201, 600000):
        assert sch(step) == onn.piecewise_lr(step, conf.boundaries_schedule_LR_C, conf.values_schedule_LR_C)
    assert sch(204800) == conf.values_schedule_LR_C[0] and sch(204801) == conf.values_schedule_LR_C[1]


def test_reference_h5_weights_forward():
    """BASELINE config 1: the reference's archived Keras weights (Results Single Integrator/.../N_try_0/*_0.h5)."""
    g = golden('h5_si_try0.npz')
    conf, env, nn, rl, batch = make('single_integrator', 128)
    rl.actor_model.set_weights([g[f'actor_{i}'] for i in range(6)])
    rl.critic_model.set_weights([g[f'critic_{i}'] for i in range(10)])
    s = batch[0]
    st = torch.tensor(s)
    a_ref = onn.actor_forward(onn.to_torch([g[f'actor_{i}'] for i in range(6)]), st, conf).numpy()
    v_ref = onn.critic_forward(onn.to_torch([g[f'critic_{i}'] for i in range(10)]), st, conf).numpy()
    assert rel(nn.eval(rl.actor_model, s), a_ref) < 2e-5
    assert rel(nn.eval(rl.critic_model, s), v_ref) < 2e-5


def test_save_weights_h5_and_recover_training(tmp_path):
    """RL_save_weights writes the reference's archive layout (RL.py:191-195) and setup_model(recover_training) reads it back
    (RL.py:91-97): <path>/N_try_<n>/{actor,critic,target_critic}_<step>.h5."""
    conf, env, nn, rl, batch = make('manipulator', 64)
    s, pr, sn, dv, d, term, w = batch
    rl.update(s, sn, pr, dv, d, term, w, fuse_target=True)             # target != critic
    conf.NNs_path = str(tmp_path)
    rl.N_try = 3
    rl.RL_save_weights(500)
    import os
    assert sorted(os.listdir(tmp_path / 'N_try_3')) == ['actor_500.h5', 'critic_500.h5', 'target_critic_500.h5']
    conf2, env2, nn2, rl2, _ = make('manipulator', 64, seed=9)
    rl2.setup_model(recover_training=(str(tmp_path), 3, 500))
    for a, b in ((rl.actor_model, rl2.actor_model), (rl.critic_model, rl2.critic_model), (rl.target_critic, rl2.target_critic)):
        assert torch.equal(a.params, b.params)
    for b in (rl2.actor_model, rl2.critic_model, rl2.target_critic):          # set_weights refreshed the transposed copies
        ref = torch.empty_like(b.params_T)
        torch.ops.cacto.transpose_params(b.params, ref, b.is_critic, b.ns, b.na)
        assert torch.equal(b.params_T, ref)


def test_recover_training_from_a_reference_results_directory():
    """The reference's own archive (Results Single Integrator/Results set test/NNs/N_try_0/*_0.h5) through setup_model."""
    import os
    root = '/root/reference/Results Single Integrator/Results set test/NNs'
    if not os.path.isdir(root):
        pytest.skip('reference checkout not present')
    g = golden('h5_si_try0.npz')
    conf, env, nn, rl, batch = make('single_integrator', 16)
    rl.setup_model(recover_training=(root, 0, 0))
    for i, a in enumerate(rl.actor_model.get_weights()):
        np.testing.assert_array_equal(a, g[f'actor_{i}'])
    for i, a in enumerate(rl.target_critic.get_weights()):
        np.testing.assert_array_equal(a, g[f'target_critic_{i}'])
