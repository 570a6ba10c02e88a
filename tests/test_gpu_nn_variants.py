"""The critic variants of NeuralNetwork.py besides 'sine' ('elu' :65-78, 'sine-elu' :80-93, 'relu' :110-128) on the generic
one-CTA-per-sample kernels (csrc/mlp_generic.cu) against the oracle's torch-autograd restatement; and the generic kernels
against the fused tiled kernels on the default critic (two independent CUDA implementations of the same step)."""
import numpy as np
import pytest
import torch

from oracle import nn as onn
from oracle import systems as osys
from test_gpu_nn import make, rel

pytestmark = pytest.mark.gpu

VARIANTS = ['elu', 'sine-elu', 'relu']


@pytest.mark.parametrize('critic_type', VARIANTS)
@pytest.mark.parametrize('system,B', [('manipulator', 64), ('car', 33), ('ur5', 9)])
def test_variant_forward_and_gradients_match_oracle(critic_type, system, B):
    conf, env, nn, rl, batch = make(system, B, critic_type=critic_type)
    assert rl.critic_model.kind == 'critic_generic'
    hidden, acts, _ = onn.critic_spec(conf)
    assert rl.critic_model.dims == [conf.nb_state] + list(hidden) + [1] and rl.critic_model.acts == list(acts) + ['linear']
    s, pr, sn, dv, d, term, w = batch
    critic, target, actor = rl.critic_model.get_weights(), rl.target_critic.get_weights(), rl.actor_model.get_weights()
    rng = np.random.default_rng(5)
    critic = [c + (0.05 * rng.normal(size=c.shape)).astype(np.float32) for c in critic]      # non-zero biases
    target = [t + (0.01 * rng.normal(size=t.shape)).astype(np.float32) for t in critic]
    rl.critic_model.set_weights(critic)
    rl.target_critic.set_weights(target)
    # forward and input gradient
    st = torch.tensor(s, requires_grad=True)
    v_ref = onn.critic_forward(onn.to_torch(critic), st, conf)
    g_ref, = torch.autograd.grad(v_ref.sum(), st)
    assert rel(nn.eval(rl.critic_model, s), v_ref.detach().numpy()) < 2e-5
    V, dV = nn.eval_with_gradient(rl.critic_model, s)
    assert rel(V, v_ref.detach().numpy()) < 2e-5 and rel(dV, g_ref.numpy()) < 5e-5
    for w_S in (1e-2, 0.0):
        nn.w_S = w_S
        cg, rtg, Vr, Vt, loss = onn.critic_grad(critic, target, conf, w_S, s, sn, pr, dv, d, w)
        g, g_rtg, g_V, g_Vt = nn.compute_critic_grad(rl.critic_model, rl.target_critic, s, sn, pr, dv, d, w)
        assert rel(g_rtg, rtg) < 2e-5 and rel(g_V, Vr) < 2e-5 and rel(g_Vt, Vt) < 2e-5
        assert abs(float(nn.last_critic_loss) - loss) <= 1e-4 * abs(loss)
        for gv, rv in zip(g, cg):
            assert rel(gv, rv) < 1e-4
    oenv = osys.make_env(conf)
    ag, actions, s_next, dQ = onn.actor_grad(actor, critic, conf, oenv, s, term)
    ga, act = nn.compute_actor_grad(rl.actor_model, rl.critic_model, s, term, None, return_actions=True)
    assert rel(act, actions) < 2e-5
    for gv, rv in zip(ga, ag):
        assert rel(gv, rv) < 1e-4


@pytest.mark.parametrize('critic_type', VARIANTS)
def test_variant_update_step_matches_oracle(critic_type):
    """RL_AC.update + fused Polyak for a variant critic: weights after two steps; then the same update replayed as a CUDA graph."""
    conf, env, nn, rl, batch = make('manipulator', 64, critic_type=critic_type)
    s, pr, sn, dv, d, term, w = batch
    critic, target, actor = rl.critic_model.get_weights(), rl.target_critic.get_weights(), rl.actor_model.get_weights()
    oc, oa = onn.Adam(critic, conf.CRITIC_LEARNING_RATE), onn.Adam(actor, conf.ACTOR_LEARNING_RATE)
    oenv = osys.make_env(conf)
    for step in range(2):
        out = onn.update(critic, target, actor, oc, oa, conf, 1e-2, oenv, (s, pr, sn, dv, d, term, w))
        rtg, V, Vt = rl.update(s, sn, pr, dv, d, term, w, fuse_target=True)
        assert rel(rtg, out['rtg']) < 2e-5 and rel(V, out['V']) < 2e-5
        for m, r in zip(rl.critic_model.get_weights() + rl.actor_model.get_weights(), critic + actor):
            assert rel(m, r) < 1e-4
        for m, r in zip(rl.target_critic.get_weights(), target):
            # a zero-initialised bias of the target is tau x the critic's: it inherits the critic's relative error (an Adam step with
            # |g| ~ eps is the sensitive case: the relu variant has such a bias, and fp32 atomics reorder from run to run)
            assert rel(m, r) < 1e-4
    ug = rl.make_update_graph(64)
    for k_, t_ in zip(('state', 'state_next', 'partial_rtg', 'dVdx', 'done', 'term', 'weights'), (s, sn, pr, dv, d, term, w)):
        ug.io[k_].copy_(torch.as_tensor(t_))
    out = onn.update(critic, target, actor, oc, oa, conf, 1e-2, oenv, (s, pr, sn, dv, d, term, w))
    ug.replay()
    torch.cuda.synchronize()
    for m, r in zip(rl.critic_model.get_weights() + rl.actor_model.get_weights(), critic + actor):
        assert rel(m, r) < 1e-4


def test_generic_kernels_agree_with_the_fused_kernels_on_the_sine_critic():
    from cacto_b200.NeuralNetwork import Network
    conf, env, nn, rl, batch = make('manipulator', 200)
    s, pr, sn, dv, d, term, w = batch
    fused = rl.critic_model
    gen = Network('critic_generic', conf.nb_state, conf.nb_action, fused.dims, ['sin'] * 4 + ['linear'])
    gen_t = Network('critic_generic', conf.nb_state, conf.nb_action, fused.dims, ['sin'] * 4 + ['linear'])
    gen.set_weights(fused.get_weights())
    tw = [t + 0.01 for t in rl.target_critic.get_weights()]
    rl.target_critic.set_weights(tw)
    gen_t.set_weights(tw)
    a = nn.compute_critic_grad(fused, rl.target_critic, s, sn, pr, dv, d, w)
    ga = [g.clone() for g in a[0]]
    la = float(nn.last_critic_loss)
    b = nn.compute_critic_grad(gen, gen_t, s, sn, pr, dv, d, w)
    assert abs(float(nn.last_critic_loss) - la) <= 1e-5 * abs(la)
    for x, y in zip(ga, b[0]):
        assert rel(y, x.cpu().numpy()) < 2e-5
    for x, y in zip(a[1:], b[1:]):
        assert rel(y, x.cpu().numpy()) < 1e-5
    fa = [g.clone() for g in nn.compute_actor_grad(rl.actor_model, fused, s, term, None)]
    fb = nn.compute_actor_grad(rl.actor_model, gen, s, term, None)
    for x, y in zip(fa, fb):
        assert rel(y, x.cpu().numpy()) < 2e-5
