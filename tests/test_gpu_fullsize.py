"""BASELINE.json full-size configurations through size-independent properties (the oracle is too slow at these sizes):
 * config 3 (car, critic/actor batch 16384): the gradient of the full batch equals the sum of the gradients of its
   shards when every shard uses the global 1/B -- the rule the NCCL data-parallel path relies on -- and the two tile
   sizes of the kernels (S = 8 for small launches, S = 16 for large ones) agree;
 * config 4 (manipulator, 131072 rollouts x 100 steps per GPU): both rollout engines agree on sampled rollouts, every
   rollout of every CTA is written, and spot rollouts match the oracle."""
import numpy as np
import pytest
import torch

from cacto_b200.conf import get_conf
from oracle import nn as onn
from oracle import rtg as ortg
from oracle import systems as osys

pytestmark = pytest.mark.gpu


def build(system):
    from cacto_b200 import environment as genv
    from cacto_b200.NeuralNetwork import NN
    from cacto_b200.RL import RL_AC
    conf = get_conf(system)
    env = genv.make_env(conf)
    nn = NN(env, conf, 1e-2, seed=0)
    rl = RL_AC(env, nn, conf, 0)
    rl.setup_model()
    return conf, nn, rl


def test_config3_batch_16384_gradient_is_sum_of_shard_gradients():
    conf, nn, rl = build('car')
    B, ns = 16384, conf.nb_state
    g = torch.Generator(device='cpu').manual_seed(0)
    lo, hi = torch.as_tensor(conf.x_init_min), torch.as_tensor(conf.x_init_max)
    s = (lo + (hi - lo) * torch.rand((B, ns), generator=g, dtype=torch.float64)).float().cuda()
    sn = (lo + (hi - lo) * torch.rand((B, ns), generator=g, dtype=torch.float64)).float().cuda()
    pr = (-5 * torch.rand((B, 1), generator=g)).cuda()
    dv = torch.randn((B, ns), generator=g).cuda()
    dv[:, -1] = 0
    d = (torch.rand((B, 1), generator=g) < 0.5).float().cuda()
    term = (torch.rand((B, 1), generator=g) < 0.01).double().cuda()
    w = (0.5 + torch.rand((B, 1), generator=g)).cuda()
    rl.target_critic.set_weights([t + 0.01 for t in rl.target_critic.get_weights()])

    gc, rtg, V, Vt = nn.compute_critic_grad(rl.critic_model, rl.target_critic, s, sn, pr, dv, d, w)
    full_c = rl.critic_model.grad.clone()
    loss_full = float(nn.last_critic_loss)
    ga = nn.compute_actor_grad(rl.actor_model, rl.critic_model, s, term, None)
    full_a = rl.actor_model.grad.clone()

    acc_c, acc_a, loss_sum = torch.zeros_like(full_c), torch.zeros_like(full_a), 0.0
    shard = 1024                                                  # 16 shards of 1024 rows -> the S = 8 tile path
    for lo_ in range(0, B, shard):
        sl = slice(lo_, lo_ + shard)
        _, rtg_s, V_s, _ = nn.compute_critic_grad(rl.critic_model, rl.target_critic, s[sl], sn[sl], pr[sl], dv[sl], d[sl], w[sl], global_batch=B)
        acc_c += rl.critic_model.grad
        loss_sum += float(nn.last_critic_loss)
        # (the full batch runs on the tensor-core engine, the 1024-row shards on the fp32-FMA engine: fp32-class agreement)
        torch.testing.assert_close(V_s, V[sl], rtol=2e-5, atol=5e-6)
        torch.testing.assert_close(rtg_s, rtg[sl], rtol=2e-5, atol=5e-6)
        nn.compute_actor_grad(rl.actor_model, rl.critic_model, s[sl], term[sl], None, global_batch=B)
        acc_a += rl.actor_model.grad
    assert float((acc_c - full_c).abs().max()) <= 1e-4 * float(full_c.abs().max())
    assert float((acc_a - full_a).abs().max()) <= 1e-4 * float(full_a.abs().max())
    assert abs(loss_sum - loss_full) <= 1e-4 * abs(loss_full)
    assert torch.isfinite(full_c).all() and torch.isfinite(full_a).all()
    # spot-check 64 rows of the big batch against the oracle's forward values
    cp = onn.to_torch(rl.critic_model.get_weights())
    v_ref = onn.critic_forward(cp, s[:64].cpu(), conf).numpy()
    np.testing.assert_allclose(V[:64].cpu().numpy(), v_ref, rtol=2e-5, atol=2e-6)


def test_config4_131072_rollouts_full_horizon():
    conf, nn, rl = build('manipulator')
    B, T, ns = 131072, conf.NSTEPS, conf.nb_state
    rng = np.random.default_rng(0)
    X0 = rng.uniform(conf.x_init_min, conf.x_init_max, (B, ns))
    X0[:, -1] = 0.0
    X = torch.tensor(X0, device='cuda')
    hz = torch.full((B,), T, dtype=torch.int32, device='cuda')
    tc = rl.rollout_batch(X, 1, horizon=hz, engine='tc')
    assert bool(tc['success'].all())
    assert not bool(torch.isnan(tc['states']).any()) and not bool(torch.isnan(tc['controls']).any())      # every rollout of every CTA written
    np.testing.assert_array_equal(tc['states'][0].T.cpu().numpy(), X0)                                      # knot 0 = initial conditions, bit-exact
    torch.testing.assert_close(tc['states'][:, -1, :], (torch.arange(T + 1, device='cuda', dtype=torch.float64) * conf.dt)[:, None].expand(-1, B),
                               rtol=0, atol=1e-12)                                                          # time column
    pick = torch.tensor([0, 1, 127, 128, 255, 256, 65535, 65536, B - 257, B - 1], device='cuda')
    fma = rl.rollout_batch(X[pick], 1, horizon=hz[pick], engine='fma')
    assert float((tc['states'][:, :, pick] - fma['states']).abs().max()) < 1e-5
    assert float((tc['controls'][:, :, pick] - fma['controls']).abs().max()) < 1e-5
    tf32 = rl.rollout_batch(X, 1, horizon=hz, engine='tf32')                                                # the two tensor-core engines, all rollouts
    assert float((tc['states'] - tf32['states']).abs().max()) < 1e-5
    assert float((tc['controls'] - tf32['controls']).abs().max()) < 1e-5
    del tf32
    # two oracle rollouts (B=1 torch actor forward + NumPy RNEA per step)
    oenv = osys.make_env(conf)
    ap = onn.to_torch(rl.actor_model.get_weights())

    def actor_eval(x):
        with torch.no_grad():
            return onn.actor_forward(ap, torch.tensor(x, dtype=torch.float32), conf).numpy()[0]
    for b in (128, B - 1):
        _, st, ct, Tb, ok = ortg.create_to_init(conf, oenv, actor_eval, 1, X0[b])
        got = tc['states'][:, :, b].cpu().numpy()
        assert ok and Tb == T and np.abs(got - st).max() <= 1e-4 * max(1.0, np.abs(st).max())
    # a second launch reproduces the first bit for bit (no atomics / races on the rollout path)
    tc2 = rl.rollout_batch(X, 1, horizon=hz, engine='tc')
    assert bool((tc2['states'] == tc['states']).all()) and bool((tc2['controls'] == tc['controls']).all())


@pytest.mark.parametrize('system,B', [('double_integrator', 65536), ('car', 262144), ('car_park', 262144)])
def test_config2_config3_full_size_rollouts(system, B):
    """BASELINE configs 2 and 3: 64 k double-integrator rollouts x 200 steps, 256 k car (500 steps) / car_park rollouts in one
    launch of the persistent tensor-core kernel; sampled rollouts agree with the fp32-FMA engine and with the oracle, every
    knot of every rollout inside its horizon is written, nothing past it."""
    conf, nn, rl = build(system)
    T, ns = conf.NSTEPS, conf.nb_state
    rng = np.random.default_rng(1)
    X0 = rng.uniform(conf.x_init_min, conf.x_init_max, (B, ns))
    X0[:, -1] = conf.dt * np.round(X0[:, -1] / conf.dt)
    X0[:4, -1] = 0.0
    tc = rl.rollout_batch(X0, 1, engine='tc')
    hz = tc['horizon']
    assert bool(tc['success'].all())
    knots = torch.arange(T + 1, device='cuda')[:, None]
    inside = knots <= hz[None, :]
    written = ~torch.isnan(tc['states'][:, 0, :])
    assert bool((written == inside).all())                                     # exactly the knots 0..NSTEPS_SH of every rollout
    pick = torch.tensor([0, 1, 127, 128, 255, 256, B // 2, B - 129, B - 1], device='cuda')
    fma = rl.rollout_batch(X0[pick.cpu().numpy()], 1, engine='fma')
    m = ~torch.isnan(fma['states'])
    sc = float(fma['states'][m].abs().max())
    assert float((tc['states'][:, :, pick][m] - fma['states'][m]).abs().max()) < 1e-5 * max(1.0, sc)
    oenv = osys.make_env(conf)
    ap = onn.to_torch(rl.actor_model.get_weights())

    def actor_eval(x):
        with torch.no_grad():
            return onn.actor_forward(ap, torch.tensor(x, dtype=torch.float32), conf).numpy()[0]
    b = int(pick[6])
    _, st, ct, Tb, ok = ortg.create_to_init(conf, oenv, actor_eval, 1, X0[b])
    got = tc['states'][:Tb + 1, :, b].cpu().numpy()
    assert ok and Tb == int(hz[b]) and np.abs(got - st).max() <= 1e-4 * max(1.0, np.abs(st).max())
