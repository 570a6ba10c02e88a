"""Generate the golden fixtures in tests/golden/ by EXECUTING THE UNMODIFIED REFERENCE
(/root/reference) in the build container.  Run:  python tests/golden/make_golden.py

TensorFlow / Pinocchio / CasADi are absent, so tests/golden/_ref_stubs.py provides NumPy
stand-ins for the few calls made on the executed paths.  What runs from the reference, as is:

* conf_<system>.py (all six)                       -> conf_constants.json
* urdf/*.urdf (parsed with xml.etree)              -> urdf_tables.json
* environment.py: SingleIntegrator/Car/CarPark simulate, derivative, augmented_derivative,
  reward, reward_batch, get_end_effector_position; DoubleIntegrator/Manipulator/UR5 reward
  with the EE position injected (Pinocchio FK is not available)        -> env_<system>.npz
* segment_tree.py, replay_buffer.py (ReplayBuffer and PrioritizedReplayBuffer, the latter
  with the three SURVEY quirk fixes Q1-Q3 injected from outside)       -> per_*.npz
* RL.py: RL_AC.RL_Solve and RL_AC.create_TO_init                       -> rtg_*.npz, toinit_*.npz
* Results */NNs/*.h5 Keras weight files (read without h5py)            -> h5_*.npz

The fixtures travel to the GPU box; /root/reference does not.
"""
import json
import os
import random
import re
import struct
import sys
import types
import xml.etree.ElementTree as ET

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)
import _ref_stubs  # noqa: E402

SYSTEMS = ['single_integrator', 'double_integrator', 'car', 'car_park', 'manipulator', 'ur5']
ENV_CLASS = dict(single_integrator='SingleIntegrator', double_integrator='DoubleIntegrator', car='Car',
                 car_park='CarPark', manipulator='Manipulator', ur5='UR5')


# ------------------------------------------------------------------------------------ confs
def dump_confs():
    out = {}
    skip = {'robot', 'simu', 'cmodel', 'cdata', 'init_states_sim', 'fig_ax_lim', 'CAMERA_TRANSFORM'}
    for s in SYSTEMS:
        m = _ref_stubs.import_conf(s)
        d = {}
        for k, v in vars(m).items():
            if k.startswith('__') or isinstance(v, types.ModuleType) or k in skip or 'path' in k.lower():
                continue
            if isinstance(v, np.ndarray):
                v = v.tolist()
            elif isinstance(v, (np.floating, np.integer)):
                v = v.item()
            if callable(v):
                continue
            d[k] = v
        out[s] = d
    txt = json.dumps(out, indent=1, sort_keys=True).replace('Infinity', '"inf"').replace('-"inf"', '"-inf"')
    open(os.path.join(HERE, 'conf_constants.json'), 'w').write(txt)


# ------------------------------------------------------------------------------------ URDF
def urdf_tables():
    out = {}
    for name in ('planar_manipulator_3dof', 'double_integrator', 'ur5_robot'):
        root = ET.parse(os.path.join(_ref_stubs.REF, 'urdf', name + '.urdf')).getroot()
        links = {}
        for l in root.findall('link'):
            ine = l.find('inertial')
            if ine is not None:
                o = ine.find('origin')
                I = ine.find('inertia').attrib
                links[l.attrib['name']] = dict(
                    mass=float(ine.find('mass').attrib['value']),
                    com=[float(x) for x in (o.attrib.get('xyz', '0 0 0') if o is not None else '0 0 0').split()],
                    com_rpy=[float(x) for x in (o.attrib.get('rpy', '0 0 0') if o is not None else '0 0 0').split()],
                    inertia=[float(I[k]) for k in ('ixx', 'iyy', 'izz', 'ixy', 'ixz', 'iyz')])
            else:
                links[l.attrib['name']] = None
        joints = []
        for j in root.findall('joint'):
            o = j.find('origin')
            ax = j.find('axis')
            joints.append(dict(name=j.attrib['name'], type=j.attrib['type'], parent=j.find('parent').attrib['link'],
                               child=j.find('child').attrib['link'],
                               xyz=[float(x) for x in o.attrib.get('xyz', '0 0 0').split()],
                               rpy=[float(x) for x in o.attrib.get('rpy', '0 0 0').split()],
                               axis=[float(x) for x in ax.attrib['xyz'].split()] if ax is not None else None))
        out[name] = dict(links=links, joints=joints)
    json.dump(out, open(os.path.join(HERE, 'urdf_tables.json'), 'w'), indent=1, sort_keys=True)


# ------------------------------------------------------------------------------------ environments
def env_goldens():
    import environment as ref_env
    from oracle import robots
    rng = np.random.default_rng(1234)
    for s in SYSTEMS:
        conf = _ref_stubs.import_conf(s)
        env = getattr(ref_env, ENV_CLASS[s])(conf)
        ns, na, nx = conf.nb_state, conf.nb_action, conf.nx
        N = 24
        lo = np.asarray(conf.x_init_min, dtype=float)
        hi = np.asarray(conf.x_init_max, dtype=float)
        states = rng.uniform(lo, hi, (N, ns))
        if s == 'car_park':                      # v and delta are 0 in the init box; spread them
            states[:, 3] = rng.uniform(-3, 3, N)
            states[:, 4] = rng.uniform(-0.5, 0.5, N)
        states[:, -1] = conf.dt * np.round(states[:, -1] / conf.dt)
        actions = rng.uniform(np.asarray(conf.u_min, float), np.asarray(conf.u_max, float), (N, na))
        term = (rng.uniform(size=(N, 1)) < 0.3).astype(float)
        weights = term.dot(np.reshape(conf.cost_weights_terminal, [1, -1])) + (1 - term).dot(np.reshape(conf.cost_weights_running, [1, -1]))
        out = dict(states=states, actions=actions, weights=weights)
        pinocchio_backed = s in ('double_integrator', 'manipulator', 'ur5')
        if pinocchio_backed:
            chain = robots.CHAINS[s]
            # FK injected: only the reward formula is pinned for these systems.
            env.get_end_effector_position = lambda st, recompute=True, _c=chain, _nq=conf.nq: np.array(_c.ee_position(np.asarray(st[:_nq], float)))
            out['ee_injected'] = np.array([env.get_end_effector_position(x) for x in states])
        else:
            out['simulate'] = np.array([env.simulate(x, u) for x, u in zip(states, actions)])
            out['derivative'] = np.array([env.derivative(x, u) for x, u in zip(states, actions)])
            fx, fu = zip(*[env.augmented_derivative(x, u) for x, u in zip(states, actions)])
            out['Fx'] = np.array(fx, dtype=float)
            out['Fu'] = np.array(fu, dtype=float)
            out['ee'] = np.array([env.get_end_effector_position(x) for x in states])
            st32, ac32 = states.astype(np.float32), actions.astype(np.float32)
            out['simulate_batch'] = np.asarray(env.simulate_batch(st32, ac32))
            out['derivative_batch'] = np.asarray(env.derivative_batch(st32, ac32))
        out['reward_sa'] = np.array([float(env.reward(w, x, u)) for w, x, u in zip(weights, states, actions)])
        out['reward_s'] = np.array([float(env.reward(w, x)) for w, x in zip(weights, states)])
        import tensorflow as tf
        out['reward_batch'] = np.asarray(env.reward_batch(weights, states.astype(np.float32), tf.convert_to_tensor(actions, dtype=np.float32)))
        np.savez_compressed(os.path.join(HERE, f'env_{s}.npz'), **out)


# ------------------------------------------------------------------------------------ TO backward pass
def backward_pass_goldens():
    """TO_Casadi.backward_pass (TO.py:119-202) executed UNMODIFIED, together with the reference's *_CAMS cost models
    (environment_TO.py) and Env.augmented_derivative (environment.py), on top of the symbolic stub of _casadi_stub.py.
    Only the systems whose models need no Pinocchio: single integrator, car and car_park."""
    import importlib
    import _casadi_stub
    _casadi_stub.install()
    import environment as ref_env
    import environment_TO as ref_env_TO
    importlib.reload(ref_env_TO)
    import TO as ref_TO
    importlib.reload(ref_TO)
    rng = np.random.default_rng(77)
    out = {}
    for s, cams in (('single_integrator', 'SingleIntegrator_CAMS'), ('car', 'Car_CAMS'), ('car_park', 'CarPark_CAMS')):
        conf = _ref_stubs.import_conf(s)
        env = getattr(ref_env, ENV_CLASS[s])(conf)
        to = ref_TO.TO_Casadi(env, conf, getattr(ref_env_TO, cams), w_S=1e-2)
        to.runningSingleModel = to.CAMS('running_model', conf)          # TO.py:43,45 (set inside TO_System_Solve)
        to.terminalModel = to.CAMS('terminal_model', conf)
        n, m = conf.nb_state - 1, conf.nb_action
        for k, T in enumerate((2, 9, 17)):
            x = rng.uniform(np.asarray(conf.x_init_min[:-1], float), np.asarray(conf.x_init_max[:-1], float))
            if s == 'car_park':                  # v and delta are 0 in the init box; spread them
                x[3], x[4] = rng.uniform(-2, 2), rng.uniform(-0.4, 0.4)
            X, U = [x], []
            for _ in range(T - 1):
                u = rng.uniform(np.asarray(conf.u_min, float), np.asarray(conf.u_max, float)) * 0.3
                U.append(u)
                X.append(np.asarray(env.simulate(np.append(X[-1], 0.0), u), dtype=float)[:-1])
            X, U = np.array(X), np.array(U).reshape(-1, m)
            Vx = to.backward_pass(T, X, U)
            out[f'{s}_{k}_X'], out[f'{s}_{k}_U'], out[f'{s}_{k}_Vx'] = X, U, np.asarray(Vx, dtype=float)
            # the cost model itself, on the same knots: -cost must be the reward (environment_TO.py vs environment.py)
            out[f'{s}_{k}_cost'] = np.array([float(to.runningSingleModel.cost(X[t], U[min(t, T - 2)])) for t in range(T)])
    np.savez_compressed(os.path.join(HERE, 'bp_cases.npz'), **out)


# ------------------------------------------------------------------------------------ PER
def _buffer_conf(R, B, ns, alpha=0.6, beta=0.6, eps=1e-2, fresh=0.95):
    return types.SimpleNamespace(REPLAY_SIZE=R, BATCH_SIZE=B, nb_state=ns, prioritized_replay_alpha=alpha,
                                 prioritized_replay_beta=beta, prioritized_replay_eps=eps, fresh_factor=fresh)


def _episodes(rng, n_ep, ns, tmin=3, tmax=12):
    eps = []
    for _ in range(n_ep):
        T = int(rng.integers(tmin, tmax))
        eps.append((rng.normal(size=(T, ns)), rng.uniform(-5, 0, T).astype(np.float32).astype(float), rng.normal(size=(T, ns)),
                    np.concatenate([rng.normal(size=(T, ns - 1)), np.zeros((T, 1))], 1),
                    (rng.uniform(size=T) < 0.5).astype(float), (np.arange(T) == T - 1).astype(float)))
    return tuple(zip(*eps))


def per_goldens():
    import segment_tree as ref_st
    import replay_buffer as ref_rb

    # --- raw segment tree: random writes, range reductions, prefix-sum searches
    rng = np.random.default_rng(7)
    cap = 64
    st_sum, st_min = ref_st.SumSegmentTree(cap), ref_st.MinSegmentTree(cap)
    w_idx = rng.integers(0, cap, 200)
    w_val = rng.uniform(1e-3, 2.0, 200)
    for i, v in zip(w_idx, w_val):
        st_sum[int(i)] = float(v)
        st_min[int(i)] = float(v)
    ranges = [(0, None), (0, cap - 1), (0, 17), (5, 40), (31, 33), (0, 1), (63, 64), (10, -3)]
    q = rng.uniform(0, st_sum.sum(), 64)
    np.savez_compressed(os.path.join(HERE, 'per_segment_tree.npz'), cap=cap, w_idx=w_idx, w_val=w_val,
                        sum_tree=np.array(st_sum._value), min_tree=np.array(st_min._value),
                        ranges=np.array([(a, -999 if b is None else b) for a, b in ranges]),
                        range_sum=np.array([st_sum.sum(a, b) for a, b in ranges]),
                        range_min=np.array([st_min.min(a, b) for a, b in ranges]),
                        queries=q, found=np.array([st_sum.find_prefixsum_idx(float(x)) for x in q]))

    # --- uniform buffer with wrap-around
    ns = 5
    conf = _buffer_conf(40, 8, ns)
    rb = ref_rb.ReplayBuffer(conf)
    rng = np.random.default_rng(11)
    rec = {}
    for r in range(4):
        ep = _episodes(rng, 3, ns)
        rb.add(*ep)
        for k, a in enumerate(ep):
            rec[f'add{r}_{k}'] = np.concatenate(a, axis=0)
        rec[f'add{r}_lens'] = np.array([len(a) for a in ep[0]])
        np.random.seed(100 + r)
        s = rb.sample()
        np.random.seed(100 + r)
        max_idx = conf.REPLAY_SIZE if rb.full else rb.next_idx
        rec[f'idx{r}'] = np.random.randint(0, max_idx, size=conf.BATCH_SIZE)
        for k in range(7):
            rec[f'sample{r}_{k}'] = np.asarray(s[k])
        rec[f'storage{r}'] = rb.storage_mat.copy()
        rec[f'next_idx{r}'] = rb.next_idx
    np.savez_compressed(os.path.join(HERE, 'per_uniform.npz'), R=conf.REPLAY_SIZE, B=conf.BATCH_SIZE, ns=ns, **rec)

    # --- prioritized buffer: quirk fixes injected from outside, sources untouched
    class _SumTree(ref_st.SumSegmentTree):          # Q2: element-wise gather for ndarray indices
        def __getitem__(self, idx):
            if isinstance(idx, np.ndarray):
                return np.array([ref_st.SumSegmentTree.__getitem__(self, int(i)) for i in idx])
            return ref_st.SumSegmentTree.__getitem__(self, idx)
    ref_rb.SumSegmentTree = _SumTree                 # Q1: unqualified names
    ref_rb.MinSegmentTree = ref_st.MinSegmentTree

    def run_per(tag, R, B, ns, rounds, n_ep, tmin, tmax, seed, keep_trees):
        conf = _buffer_conf(R, B, ns)
        pb = ref_rb.PrioritizedReplayBuffer(conf)
        pb.RB_type = 'PER'                           # Q3
        rng = np.random.default_rng(seed)
        rec = dict(R=R, B=B, ns=ns, rounds=rounds)
        for r in range(rounds):
            ep = _episodes(rng, n_ep, ns, tmin, tmax)
            pb.add(*ep)
            for k, a in enumerate(ep):
                rec[f'add{r}_{k}'] = np.concatenate(a, axis=0)
            rec[f'add{r}_lens'] = np.array([len(a) for a in ep[0]])
            for it in range(2):
                random.seed(1000 * r + it)
                s = pb.sample()
                random.seed(1000 * r + it)
                rec[f'u{r}_{it}'] = np.array([random.random() for _ in range(B)])
                idx = s[7]
                rec[f'idx{r}_{it}'] = idx
                rec[f'w{r}_{it}'] = np.asarray(s[6])
                for k in range(6):
                    rec[f'sample{r}_{it}_{k}'] = np.asarray(s[k])
                rtg = rng.normal(size=(B, 1)).astype(np.float32)
                V = rng.normal(size=(B, 1)).astype(np.float32)
                rec[f'rtg{r}_{it}'], rec[f'V{r}_{it}'] = rtg, V
                pb.update_priorities(idx, rtg, V)
                rec[f'maxp{r}_{it}'] = pb._max_priority
                if keep_trees:
                    rec[f'sum{r}_{it}'] = np.array(pb._it_sum._value)
                    rec[f'min{r}_{it}'] = np.array(pb._it_min._value)
                rec[f'expc{r}_{it}'] = pb.exp_counter.copy()
        rec['sum_final'] = np.array(pb._it_sum._value)
        rec['min_final'] = np.array(pb._it_min._value)
        np.savez_compressed(os.path.join(HERE, f'per_{tag}.npz'), **rec)

    run_per('small', R=48, B=16, ns=4, rounds=4, n_ep=3, tmin=3, tmax=12, seed=21, keep_trees=True)
    run_per('medium', R=2048, B=256, ns=7, rounds=3, n_ep=10, tmin=60, tmax=101, seed=22, keep_trees=False)


# ------------------------------------------------------------------------------------ RL_Solve / create_TO_init
def rtg_goldens():
    import RL as ref_rl
    rng = np.random.default_rng(5)
    rec = {}
    cases = [(100, 50, 0, 7), (37, 50, 0, 7), (200, 50, 0, 5), (1, 25, 0, 3), (60, 25, 1, 6), (500, 125, 0, 6), (50, 50, 0, 13)]
    for k, (T, n, MC, ns) in enumerate(cases):
        conf = types.SimpleNamespace(REPLAY_SIZE=16, nb_state=ns, env_RL=0, MC=MC, nsteps_TD_N=n)
        rl = ref_rl.RL_AC(None, None, conf, 0)
        rl.NSTEPS_SH = T
        states = rng.normal(size=(T + 1, ns))
        cost = rng.uniform(0, 3, T + 1) * 10.0 ** rng.integers(-6, 1, T + 1)
        out = rl.RL_Solve(rng.normal(size=(T, 2)), states, cost)
        rec[f'c{k}_meta'] = np.array([T, n, MC, ns])
        rec[f'c{k}_states'], rec[f'c{k}_cost'] = states, cost
        rec[f'c{k}_partial'], rec[f'c{k}_total'], rec[f'c{k}_snext'] = out[1], out[2], out[3]
        rec[f'c{k}_done'], rec[f'c{k}_rwrd'], rec[f'c{k}_term'], rec[f'c{k}_ret'] = out[4], out[5], out[6], out[7]
    rec['ncases'] = len(cases)
    np.savez_compressed(os.path.join(HERE, 'rtg_cases.npz'), **rec)


def toinit_goldens():
    """create_TO_init through the reference loop.  ep = 0 (zero controls) is reference-only;
    for ep = 1 the actor forward is the oracle's (TensorFlow is absent), so only the loop
    logic (horizon, indexing, time column) is pinned there."""
    import RL as ref_rl
    import environment as ref_env
    import torch
    from oracle import nn as onn
    rec = {}
    for s in ('single_integrator', 'car', 'car_park'):
        conf = _ref_stubs.import_conf(s)
        env = getattr(ref_env, ENV_CLASS[s])(conf)
        actor = onn.init_actor(conf.nb_state, conf.nb_action, seed=3)
        ap = onn.to_torch(actor)

        class _NN:
            def eval(self, model, x):
                import tensorflow as tf
                with torch.no_grad():
                    return tf.convert_to_tensor(onn.actor_forward(ap, torch.tensor(np.asarray(x), dtype=torch.float32), conf).numpy())
        rl = ref_rl.RL_AC(env, _NN(), conf, 0)
        random.seed(9)
        for k in range(3):
            ICS = env.reset()
            if k == 2:
                ICS[-1] = 0.0
            for ep in (0, 1):
                _, st, ct, T, ok = rl.create_TO_init(ep, ICS)
                rec[f'{s}_{k}_{ep}_ics'] = ICS
                rec[f'{s}_{k}_{ep}_states'], rec[f'{s}_{k}_{ep}_controls'] = st, ct
                rec[f'{s}_{k}_{ep}_T'] = T
        for i, a in enumerate(actor):
            rec[f'{s}_actor_{i}'] = a
    np.savez_compressed(os.path.join(HERE, 'toinit_cases.npz'), **rec)


# ------------------------------------------------------------------------------------ Keras .h5
def read_keras_h5(path):
    """Contiguous little-endian fp32 datasets of a Keras-2.11 save_weights file, in object
    header order = [kernel, bias] per layer group (groups sorted by name)."""
    b = open(path, 'rb').read()
    ds = []
    for m in re.finditer(rb'\x08\x00\x18\x00.\x00\x00\x00\x03\x01', b, re.S):
        addr, size = struct.unpack('<QQ', b[m.end():m.end() + 16])
        if 0 < size and addr + size <= len(b) and size % 4 == 0:
            ds.append(np.frombuffer(b[addr:addr + size], '<f4').copy())
    assert len(ds) % 2 == 0
    return [(ds[i], ds[i + 1]) for i in range(0, len(ds), 2)]


def order_layers(pairs, ns):
    """Chain (kernel, bias) pairs into network order starting from fan_in = ns (backtracking:
    a 128->1 head and a 128->128 layer both fit after a 128-wide layer)."""
    def rec(rem, fan_in):
        if not rem:
            return []
        for i, (k, bb) in enumerate(rem):
            if k.size == fan_in * bb.size:
                tail = rec(rem[:i] + rem[i + 1:], bb.size)
                if tail is not None:
                    return [k.reshape(fan_in, bb.size), bb] + tail
        return None
    out = rec(list(pairs), ns)
    if out is None:
        raise ValueError('cannot chain layers')
    return out


def h5_goldens():
    base = _ref_stubs.REF
    jobs = [('si_try0', 'Results Single Integrator/Results set test/NNs/N_try_0', '0', 3),
            ('di_try6_final', 'Results Double Integrator/Results set test/NNs/N_try_6', 'final', 5)]
    for tag, d, step, ns in jobs:
        rec = {}
        for net in ('actor', 'critic', 'target_critic'):
            layers = order_layers(read_keras_h5(os.path.join(base, d, f'{net}_{step}.h5')), ns)
            for i, a in enumerate(layers):
                rec[f'{net}_{i}'] = a
        np.savez_compressed(os.path.join(HERE, f'h5_{tag}.npz'), ns=ns, **rec)


if __name__ == '__main__':
    assert _ref_stubs.have_reference(), 'needs /root/reference (build container only)'
    _ref_stubs.install()
    dump_confs()
    urdf_tables()
    env_goldens()
    per_goldens()
    backward_pass_goldens()
    rtg_goldens()
    toinit_goldens()
    h5_goldens()
    print('golden fixtures written to', HERE)
