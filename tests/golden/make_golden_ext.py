#!/usr/bin/env python
"""PINNING KIT -- golden vectors for the rows of SURVEY.md section 8 that stay "parity unpinned" in this repository because
their arithmetic lives in third-party modules the build container does not have:

    pinocchio (pin3x-jnrh2023 == 2.9.2)         A3 Env.simulate, A6 derivative, A7 augmented_derivative, A11 EE position
                                                for double_integrator / manipulator / ur5, and their reward functions end to end
    tensorflow == 2.11 + tf_siren == 0.0.5      N2-N4 forward passes, N6 compute_critic_grad, N7 compute_actor_grad,
                                                N8 Adam steps, N9 Polyak, with w_S in {1e-2, 0} and conf.MC in {0, 1}

Run it ONCE where the reference runs (the reference's own environment: `pip install -r requirements.txt`):

    cd <checkout of nadimkanazi/cacto>
    python <this repo>/tests/golden/make_golden_ext.py --out <this repo>/tests/golden [--systems manipulator ur5 ...]

It executes the UNMODIFIED reference modules (environment.py, NeuralNetwork.py, RL.py, conf_*.py) on seeded inputs and writes
`ext_env_<system>.npz` / `ext_nn_<system>.npz`.  `tests/test_golden_ext.py` picks the files up when they exist: the oracle is
then checked against them on the CPU (`-m "not gpu"`) and the CUDA kernels on the GPU (`-m gpu`); until then those tests skip
and DESIGN.md lists the rows as unpinned.  This script cannot run in the build container (no tensorflow / pinocchio wheels, no
network) and has therefore only been checked for syntax and against the reference's call signatures (file:line below).

Schema (all arrays C-order, fp64 unless noted)
  ext_env_<system>.npz
    state [N, ns], action [N, na]                     seeded U(x_init_min, x_init_max) / U(u_min, u_max) * 0.5, time on the dt grid
    next [N, ns]                                      Env.simulate(state, action)                     environment.py:80-91
    Fu_norm [N, ns, na]                               Env.derivative(state, action)                   environment.py:93-109
    Fx [N, nx, nx], Fu [N, nx, na]                    Env.augmented_derivative(state, action)         environment.py:111-132
    ee [N, 3]                                         Env.get_end_effector_position(state)            environment.py:146-156
    reward_run [N], reward_ter [N]                    Env.reward(w_running | w_terminal, state, action | None)
    next_batch_f32 [N, ns] f32, Fu_batch_f32 [N, ns, na] f32, reward_batch_f32 [N, 1] f32
                                                      simulate_batch / derivative_batch / reward_batch on float32 inputs (quirk Q13)
  ext_nn_<system>.npz   (one group per case c in {sobolev, value_only, mc}: keys prefixed "<c>/")
    actor_w{i}, critic_w{i}, target_w{i}              initial weights, Keras order (kernel (in, out), bias)
    s, pr, sn, dv, d, term, w                         the minibatch (f32; term f64) in replay-buffer layout
    actor_out [B, na] f32, critic_out [B, 1] f32      NN.eval                                         NeuralNetwork.py:130-138
    critic_grad{i}, rtg, V, Vt                        NN.compute_critic_grad                          NeuralNetwork.py:150-178
    actor_grad{i}                                     NN.compute_actor_grad                           NeuralNetwork.py:180-232
    step{k}_critic_w{i}, step{k}_actor_w{i}, step{k}_target_w{i}   after k = 1, 2 calls of RL_AC.update (+ update_target when not MC)
                                                                                                       RL.py:101-118, 134-135
"""
import argparse
import importlib
import os
import random
import sys

import numpy as np

ENV_CLASS = {'single_integrator': 'SingleIntegrator', 'double_integrator': 'DoubleIntegrator', 'car': 'Car', 'car_park': 'CarPark',
             'manipulator': 'Manipulator', 'ur5': 'UR5'}
PINOCCHIO_SYSTEMS = ('double_integrator', 'manipulator', 'ur5')


def seeded_states(conf, n, rng):
    lo, hi = np.asarray(conf.x_init_min, float), np.asarray(conf.x_init_max, float)
    s = rng.uniform(lo, hi, (n, conf.nb_state))
    s[:, -1] = conf.dt * np.round(s[:, -1] / conf.dt)
    a = rng.uniform(np.asarray(conf.u_min, float), np.asarray(conf.u_max, float), (n, conf.nb_action)) * 0.5
    return s, a


def env_goldens(system, out_dir, n=64):
    import environment                                   # the reference's module (needs pinocchio for these systems)
    conf = importlib.import_module('conf_' + system)
    env = getattr(environment, ENV_CLASS[system])(conf)
    rng = np.random.default_rng(20)
    s, a = seeded_states(conf, n, rng)
    nx = conf.nb_state - 1
    rec = dict(state=s, action=a, next=np.zeros_like(s), Fu_norm=np.zeros((n, conf.nb_state, conf.nb_action)), Fx=np.zeros((n, nx, nx)),
               Fu=np.zeros((n, nx, conf.nb_action)), ee=np.zeros((n, 3)), reward_run=np.zeros(n), reward_ter=np.zeros(n))
    for i in range(n):
        rec['next'][i] = env.simulate(s[i], a[i])
        rec['Fu_norm'][i] = env.derivative(s[i], a[i])
        rec['Fx'][i], rec['Fu'][i] = env.augmented_derivative(s[i], a[i])
        rec['ee'][i] = np.asarray(env.get_end_effector_position(s[i])).reshape(-1)[:3]
        rec['reward_run'][i] = env.reward(conf.cost_weights_running, s[i], a[i])
        rec['reward_ter'][i] = env.reward(conf.cost_weights_terminal, s[i])
    import tensorflow as tf
    s32, a32 = s.astype(np.float32), a.astype(np.float32)
    rec['next_batch_f32'] = np.asarray(env.simulate_batch(s32, a32), dtype=np.float32)
    rec['Fu_batch_f32'] = np.asarray(env.derivative_batch(s32, a32), dtype=np.float32)
    w = np.tile(np.asarray(conf.cost_weights_running, float), (n, 1))
    rec['reward_batch_f32'] = np.asarray(env.reward_batch(w, s32, tf.convert_to_tensor(a32)), dtype=np.float32)
    np.savez_compressed(os.path.join(out_dir, f'ext_env_{system}.npz'), **rec)
    print('wrote ext_env_%s.npz' % system)


def nn_goldens(system, out_dir, B=64):
    import tensorflow as tf
    import environment
    from NeuralNetwork import NN
    from RL import RL_AC
    rec = {}
    for case, w_S, mc in (('sobolev', 1e-2, 0), ('value_only', 0.0, 0), ('mc', 1e-2, 1)):
        conf = importlib.reload(importlib.import_module('conf_' + system))
        conf.MC = mc
        conf.BATCH_SIZE = B
        random.seed(0); np.random.seed(0); tf.random.set_seed(0)
        env = getattr(environment, ENV_CLASS[system])(conf)
        nn = NN(env, conf, w_S)
        rl = RL_AC(env, nn, conf, 0)
        rl.setup_model()
        # decorrelate target and critic so that V_target matters
        rl.target_critic.set_weights([t + 0.01 * np.random.default_rng(5).normal(size=t.shape).astype(np.float32) for t in rl.target_critic.get_weights()])
        rng = np.random.default_rng(31)
        s, _ = seeded_states(conf, B, rng)
        sn, _ = seeded_states(conf, B, rng)
        s, sn = s.astype(np.float32), sn.astype(np.float32)
        pr = rng.uniform(-5, 0, (B, 1)).astype(np.float32)
        dv = rng.normal(size=(B, conf.nb_state)).astype(np.float32)
        dv[:, -1] = 0
        dv[0, 0] = 0.0                                   # the slog gate at exactly 0
        d = (rng.uniform(size=(B, 1)) < 0.5).astype(np.float32)
        term = (rng.uniform(size=(B, 1)) < 0.2).astype(np.float64)
        w = rng.uniform(0.2, 1.5, (B, 1)).astype(np.float32)
        P = case + '/'
        for name, model in (('actor', rl.actor_model), ('critic', rl.critic_model), ('target', rl.target_critic)):
            for i, a in enumerate(model.get_weights()):
                rec[f'{P}{name}_w{i}'] = a
        for k, v in (('s', s), ('pr', pr), ('sn', sn), ('dv', dv), ('d', d), ('term', term), ('w', w)):
            rec[P + k] = v
        T = lambda x: tf.convert_to_tensor(x, dtype=tf.float32)          # replay_buffer.convert_sample_to_tensor (:74-83); term stays NumPy f64
        rec[P + 'actor_out'] = nn.eval(rl.actor_model, s).numpy()
        rec[P + 'critic_out'] = nn.eval(rl.critic_model, s).numpy()
        cg, rtg, V, Vt = nn.compute_critic_grad(rl.critic_model, rl.target_critic, T(s), T(sn), T(pr), T(dv), T(d), T(w))
        for i, g in enumerate(cg):
            rec[f'{P}critic_grad{i}'] = g.numpy()
        rec[P + 'rtg'], rec[P + 'V'], rec[P + 'Vt'] = np.asarray(rtg), V.numpy(), Vt.numpy()
        ag = nn.compute_actor_grad(rl.actor_model, rl.critic_model, T(s), term, B)
        for i, g in enumerate(ag):
            rec[f'{P}actor_grad{i}'] = g.numpy()
        for k in (1, 2):
            rl.update(T(s), T(sn), T(pr), T(dv), T(d), term, T(w), B)
            if not conf.MC:
                rl.update_target(rl.target_critic.variables, rl.critic_model.variables)
            for name, model in (('actor', rl.actor_model), ('critic', rl.critic_model), ('target', rl.target_critic)):
                for i, a in enumerate(model.get_weights()):
                    rec[f'{P}step{k}_{name}_w{i}'] = a
    np.savez_compressed(os.path.join(out_dir, f'ext_nn_{system}.npz'), **rec)
    print('wrote ext_nn_%s.npz' % system)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--out', required=True)
    ap.add_argument('--reference', default=os.getcwd(), help='checkout of nadimkanazi/cacto (default: the current directory)')
    ap.add_argument('--systems', nargs='*', default=['single_integrator', 'double_integrator', 'car', 'manipulator', 'ur5'])
    args = ap.parse_args()
    sys.path.insert(0, args.reference)
    os.makedirs(args.out, exist_ok=True)
    for system in args.systems:
        if system in PINOCCHIO_SYSTEMS:
            env_goldens(system, args.out)
        nn_goldens(system, args.out)


if __name__ == '__main__':
    main()
