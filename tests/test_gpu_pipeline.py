"""One CACTO training loop end to end (main.py:216-240), GPU modules against the oracle modules, stage by stage and as a whole:
create_TO_init rollouts (K1) -> [the TO solve is out of scope: its solution is taken to be the warm start itself] -> step costs
(Env.reward) -> TO backward pass (K6, dVdx) -> RL_Solve windows (K5) -> buffer.add -> learn_and_update (K4 sampling, K3 update as
a replayed CUDA graph, Polyak) -- two episodes, so that the second one rolls out the UPDATED actor.  What is compared: the replay
storage after every add, the priorities trees (PER), and every network after every episode."""
import random

import numpy as np
import pytest
import torch

from cacto_b200.conf import get_conf
from oracle import backward as obw
from oracle import nn as onn
from oracle import per as oper
from oracle import rtg as ortg
from oracle import systems as osys

pytestmark = pytest.mark.gpu


def rel(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


def draw_ics(conf, B, seed):
    rng = np.random.default_rng(seed)
    x = rng.uniform(np.asarray(conf.x_init_min, float), np.asarray(conf.x_init_max, float), (B, conf.nb_state))
    x[:, -1] = conf.dt * np.round(x[:, -1] / conf.dt)
    x[0, -1] = 0.0
    return x


@pytest.mark.parametrize('system,alpha', [('single_integrator', 0.0), ('single_integrator', 0.6), ('car', 0.0), ('manipulator', 0.0)])
def test_two_episodes_of_the_training_loop_match_the_oracle(system, alpha):
    from cacto_b200 import environment as genv
    from cacto_b200.NeuralNetwork import NN
    from cacto_b200.RL import RL_AC
    from cacto_b200.TO import TO_Casadi
    from cacto_b200.replay_buffer import PrioritizedReplayBuffer, ReplayBuffer
    over = dict(BATCH_SIZE=16, REPLAY_SIZE=2 ** 11, prioritized_replay_alpha=alpha)
    if system != 'single_integrator':
        over['NSTEPS'] = 60 if system == 'car' else 30
    conf = get_conf(system, **over)
    conf.UPDATE_LOOPS = np.array([3, 3])
    conf.save_interval = 10 ** 9
    w_S, n_ics = 1e-2, 4
    env = genv.make_env(conf)
    nn = NN(env, conf, w_S, seed=0)
    rl = RL_AC(env, nn, conf, 0)
    rl.setup_model()
    trop = TO_Casadi(env, conf, None, w_S)
    buf = ReplayBuffer(conf) if alpha == 0 else PrioritizedReplayBuffer(conf)

    oenv = osys.make_env(conf)
    critic, target, actor = rl.critic_model.get_weights(), rl.target_critic.get_weights(), rl.actor_model.get_weights()
    oc, oa = onn.Adam(critic, conf.values_schedule_LR_C[0]), onn.Adam(actor, conf.values_schedule_LR_A[0])
    obuf = oper.ReplayBuffer(conf) if alpha == 0 else oper.PrioritizedReplayBuffer(conf)

    counter, n_rows = 0, 0
    for ep in range(2):
        X0 = draw_ics(conf, n_ics, 10 + ep)
        # ------------------------------------------------------------------ GPU modules
        out = rl.rollout_batch(X0, ep, with_reward=True)
        assert out['success'].cpu().numpy().all()
        hz = out['horizon'].cpu().numpy()
        S = out['states'].permute(2, 0, 1).cpu().numpy()
        U = out['controls'].permute(2, 0, 1).cpu().numpy()
        R = out['rewards'].permute(1, 0).cpu().numpy()
        TO_states = [S[b, :hz[b] + 1] for b in range(n_ics)]
        TO_controls = [U[b, :hz[b]] for b in range(n_ics)]
        TO_cost = [-R[b, :hz[b] + 1] for b in range(n_ics)]
        Vx, offs = trop.backward_pass_batch(TO_states, TO_controls)
        Vx = Vx.cpu().numpy()
        r = rl.rtg_batch(TO_states, TO_cost)
        cut = lambda t: [t.cpu().numpy()[offs[e]:offs[e + 1]] for e in range(n_ics)]
        buf.add(TO_states, cut(r['partial']), cut(r['state_next']), [Vx[offs[e]:offs[e + 1]] for e in range(n_ics)], cut(r['done']), cut(r['term']))
        # ------------------------------------------------------------------ oracle modules
        ap = onn.to_torch(actor)

        def actor_eval(x):
            with torch.no_grad():
                return onn.actor_forward(ap, torch.tensor(x, dtype=torch.float32), conf).numpy()[0]
        lists = [[] for _ in range(6)]
        for b in range(n_ics):
            _, st, ct, T, ok = ortg.create_to_init(conf, oenv, actor_eval, ep, X0[b])
            assert ok == 1 and T == hz[b]
            cost = [-oenv.reward(conf.cost_weights_running, st[t], ct[t]) for t in range(T)] + [-oenv.reward(conf.cost_weights_terminal, st[T])]
            dV = obw.backward_pass(oenv, T + 1, st, ct)
            state_arr, partial, total, s_next, done, rwrd, term, ep_ret = ortg.rl_solve(conf, st, cost)
            for l, v in zip(lists, (state_arr, partial, s_next, dV, done, term)):
                l.append(v)
        obuf.add(*lists)
        # ------------------------------------------------------------------ the experiences both sides stored
        n_new = int(sum(hz + 1))
        got = buf.storage_mat[n_rows:n_rows + n_new].cpu().numpy()
        want = obuf.storage_mat[n_rows:n_rows + n_new]
        ns = conf.nb_state
        tol = 1e-9 if ep == 0 else 2e-3                        # episode 1 rolls out the actor both sides have trained (fp32, 1e-4 apart)
        assert rel(got[:, :ns], want[:, :ns]) < tol                                   # states
        assert rel(got[:, ns], want[:, ns]) < max(tol, 1e-6)                          # partial reward-to-go (float32-rounded)
        assert rel(got[:, ns + 1:2 * ns + 1], want[:, ns + 1:2 * ns + 1]) < tol       # n-step next states
        assert rel(got[:, 2 * ns + 1:3 * ns + 1], want[:, 2 * ns + 1:3 * ns + 1]) < max(tol, 1e-6) * 10   # dVdx
        assert np.array_equal(got[:, 3 * ns + 1:], want[:, 3 * ns + 1:])              # done, term
        n_rows += n_new
        assert buf.next_idx == obuf.next_idx == n_rows
        # ------------------------------------------------------------------ updates (same index draws: same seeds)
        np.random.seed(77 + ep); random.seed(78 + ep)
        counter = rl.learn_and_update(counter, buf, ep)
        np.random.seed(77 + ep); random.seed(78 + ep)
        for _ in range(int(conf.UPDATE_LOOPS[ep])):
            s_, r_, s1_, dv_, d_, term_, w_, idx_ = obuf.sample()
            o = onn.update(critic, target, actor, oc, oa, conf, w_S, oenv, (s_, r_, s1_, dv_, d_, term_, w_))
            if alpha != 0:
                obuf.update_priorities(idx_, o['rtg'], o['V'], o['V_target'])
        assert counter == 3 * (ep + 1) and rl.update_graph is not None
        wt = 2e-4 if ep == 0 else 5e-3
        for m, ref in zip(rl.critic_model.get_weights() + rl.actor_model.get_weights() + rl.target_critic.get_weights(), critic + actor + target):
            assert rel(m, ref) < wt
        if alpha != 0:
            np.testing.assert_allclose(buf._it_sum._value.cpu().numpy()[1], obuf._it_sum.val[1], rtol=1e-3 if ep == 0 else 2e-2)
