"""Configuration objects for the six CACTO systems.

The reference keeps one Python module of constants per system (conf_<system>.py, imported by
name in main.py:108-112) and every hot-path class reads ``self.conf.X``.  The classes in this
package accept either a reference conf module or the namespaces built here, which carry the
same attribute names and values but none of the Pinocchio / CasADi objects (``robot``,
``simu``, ``cmodel``): the rigid-body parameters the kernels need come from
``cacto_b200.robots``.

Values are checked against a dump of the reference's conf modules in
tests/golden/conf_constants.json (tests/test_conf.py).
"""
import math
from types import SimpleNamespace

import numpy as np

_PI = math.pi

# Entries common to every system (conf_*.py "CACTO parameters" / "NN" blocks).
_COMMON = dict(
    CRITIC_LEARNING_RATE=5e-4, ACTOR_LEARNING_RATE=1e-3, REPLAY_SIZE=2 ** 16, MC=0, UPDATE_RATE=0.001,
    critic_type='sine', NH1=256, NH2=256, NORMALIZE_INPUTS=1,
    kreg_l1_A=1e-2, kreg_l2_A=1e-2, breg_l1_A=1e-2, breg_l2_A=1e-2,
    kreg_l1_C=1e-2, kreg_l2_C=1e-2, breg_l1_C=1e-2, breg_l2_C=1e-2,
    prioritized_replay_alpha=0, prioritized_replay_beta=0.6, prioritized_replay_beta_iters=None,
    prioritized_replay_eps=1e-2, fresh_factor=0.95,
    offset_cost_fun=0, scale_cost_fun=1e-5, env_RL=0, profile=0, save_flag=1,
    simulate_coulomb_friction=0, simulation_type='euler', integration_scheme='E-Euler',
)

# 2-D obstacle set shared by SI / DI / car / manipulator: (XC, YC, A, B) x 3
_ELL_2D = ((-2.0, 0.0, 6, 10), (3.0, 4.0, 12, 4), (3.0, -4.0, 12, 4))

_SYSTEMS = dict(
    single_integrator=dict(
        EP_UPDATE=200, NUPDATES=100000, loops_stop=25000, NSTEPS=100, BATCH_SIZE=128, td_div=4, LR_SCHEDULE=0,
        save_interval=5000, ell=_ELL_2D, w=(100, 10, 5e5, 5e6, 0), alpha=50, alpha2=5, target=(-7.0, 0.0),
        dt=0.05, nq=None, nv=None, nx=2, na=2, init_lo=(-15, -15), init_hi=(15.0, 15.0), t_min=0.0,
        norm=(15, 15), u_lo=(-6.0, -6.0), u_hi=(6.0, 6.0)),
    double_integrator=dict(
        EP_UPDATE=200, NUPDATES=50000, loops_stop=19000, NSTEPS=200, BATCH_SIZE=128, td_div=4, LR_SCHEDULE=0,
        save_interval=5000, ell=_ELL_2D, w=(100, 10, 5e5, 5e6, 0), alpha=50, alpha2=5, target=(-7.0, 0.0),
        dt=0.05, nq=2, nv=2, nx=4, na=2, init_lo=(-15.0, -15.0, -6.0, -6.0), init_hi=(15.0, 15.0, 6.0, 6.0),
        t_min=0.05, norm=(15, 15, 6, 6), u_lo=(-2.0, -2.0), u_hi=(2.0, 2.0),
        prioritized_replay_eps=1e-4, fresh_factor=1),
    car=dict(
        EP_UPDATE=250, NUPDATES=260000, loops_stop=40000, NSTEPS=500, BATCH_SIZE=64, td_div=4, LR_SCHEDULE=0,
        save_interval=10000, ell=_ELL_2D, w=(100.0, 10.0, 5e5, 5e6, 0), alpha=50, alpha2=5, target=(-7.0, 0.0),
        dt=0.05, nq=None, nv=None, nx=5, na=2, init_lo=(-15.0, -15.0, -_PI, -10.0, -3.0),
        init_hi=(15.0, 15.0, _PI, 10.0, 3.0), t_min=0.0, norm=(15.0, 15.0, _PI, 10.0, 3.0),
        u_lo=(-2, -1), u_hi=(2, 1)),
    car_park=dict(
        EP_UPDATE=200, NUPDATES=260000, loops_stop=40000, NSTEPS=100, BATCH_SIZE=64, td_div=2, LR_SCHEDULE=0,
        save_interval=10000, ell=((-10, 6.75, 17, 4.5), (10, 6.75, 17, 4.5), (0, -2, 40, 4)),
        w=(100.0, 10.0, 1e6, 5e4, 100.0), alpha=50, alpha2=1, target=(0.0, 6.75),
        dt=0.05, nq=None, nv=None, nx=5, na=2, init_lo=(-10.0, 1.5, -_PI / 6, 0.0, 0.0),
        init_hi=(10.0, 3.0, _PI / 6, 0.0, 0.0), t_min=0.0, norm=(10.0, 3.0, _PI, 10.0, _PI / 6),
        u_lo=(-3, -1), u_hi=(3, 1)),
    manipulator=dict(
        EP_UPDATE=200, NUPDATES=380000, loops_stop=50000, NSTEPS=100, BATCH_SIZE=64, td_div=2, LR_SCHEDULE=1,
        save_interval=15000, ell=_ELL_2D, w=(100, 1, 5e5, 5e6, 1e4), alpha=50, alpha2=50, target=(-20.0, 0.0),
        dt=0.05, nq=3, nv=3, nx=6, na=3, init_lo=(-_PI,) * 3 + (-_PI / 4,) * 3, init_hi=(_PI,) * 3 + (_PI / 4,) * 3,
        t_min=0.0, norm=(15, 15, 15, 10, 10, 10), u_lo=(-200.0,) * 3, u_hi=(200.0,) * 3),
    ur5=dict(
        EP_UPDATE=200, NUPDATES=380000, loops_stop=50000, NSTEPS=100, BATCH_SIZE=64, td_div=4, LR_SCHEDULE=0,
        save_interval=5000, w=(100, 1, 5e5, 5e6, 0), alpha=50, alpha2=5, target=(0.0, 0.425, 0.2),
        dt=0.01, nq=6, nv=6, nx=12, na=6, init_lo=(-_PI,) * 6 + (-_PI / 4,) * 6, init_hi=(_PI,) * 6 + (_PI / 4,) * 6,
        t_min=0.0, norm=(10,) * 12, u_lo=(-150, -150, -150, -28, -28, -28), u_hi=(150, 150, 150, 28, 28, 28)),
)

# UR5 ellipsoids: centre (x, y, z), axes (A, B, C)  (conf_ur5.py obstacle block)
_UR5_ELL = (((0.0, 0.25, 0.2), (0.5, 0.2, 0.34)), ((0.2, 0.425, 0.2), (0.4, 0.14, 0.34)),
            ((-0.2, 0.425, 0.2), (0.4, 0.14, 0.34)))

SYSTEM_IDS = tuple(_SYSTEMS)


def get_conf(system_id, **overrides):
    """Build the conf namespace of ``system_id`` ('single_integrator', 'double_integrator',
    'car', 'car_park', 'manipulator', 'ur5').  Keyword overrides replace attributes
    (e.g. ``BATCH_SIZE=4096``, ``prioritized_replay_alpha=0.6``)."""
    s = dict(_SYSTEMS[system_id])
    c = dict(_COMMON)
    c['system_id'] = system_id
    # overrides of PRIMARY constants (REPLAY_SIZE, BATCH_SIZE, NSTEPS, dt, learning rates, ...) come first so that everything the
    # reference's conf modules derive from them (nsteps_TD_N, LR-schedule boundaries, x_init_max, state_norm_arr, ...) follows
    for k, v in overrides.items():
        if k in s:
            s[k] = v
        elif k in c:
            c[k] = v
    for k in ('prioritized_replay_eps', 'fresh_factor'):
        if k in s:
            c[k] = s.pop(k)
    for k in ('EP_UPDATE', 'NUPDATES', 'NSTEPS', 'BATCH_SIZE', 'LR_SCHEDULE', 'save_interval', 'dt',
              'nq', 'nv', 'nx', 'na', 'alpha', 'alpha2'):
        c[k] = s[k]
    c['UPDATE_LOOPS'] = np.arange(1000, s['loops_stop'], 3000)
    c['NLOOPS'] = len(c['UPDATE_LOOPS'])
    c['NEPISODES'] = int(c['EP_UPDATE'] * c['NLOOPS'])
    c['nsteps_TD_N'] = int(c['NSTEPS'] / s['td_div'])

    # learning-rate schedule (PiecewiseConstantDecay boundaries / values, RL.py:82-85)
    bnd = [m * c['REPLAY_SIZE'] / c['BATCH_SIZE'] for m in (200, 300, 400, 500)]
    c['boundaries_schedule_LR_C'] = list(bnd)
    c['boundaries_schedule_LR_A'] = list(bnd)
    c['values_schedule_LR_C'] = [c['CRITIC_LEARNING_RATE'] / d for d in (1, 2, 4, 8, 16)]
    c['values_schedule_LR_A'] = [c['ACTOR_LEARNING_RATE'] / d for d in (1, 2, 4, 8, 16)]

    # cost function
    w_d, w_u, w_peak, w_ob, w_v = s['w']
    c.update(w_d=w_d, w_u=w_u, w_peak=w_peak, w_ob=w_ob, w_v=w_v, w_b=1 / w_u)
    run = [w_d, w_peak, 0., w_ob, w_ob, w_ob, w_u]
    ter = [w_d, w_peak, w_v, w_ob, w_ob, w_ob, 0]
    if system_id == 'ur5':
        cen = [x for e in _UR5_ELL for x in e[0]]
        axes = [x for e in _UR5_ELL for x in e[1]]
        c['obs_param'] = np.array(cen + axes)
        for k, (ce, ax) in enumerate(_UR5_ELL, 1):
            c.update({f'XC{k}': ce[0], f'YC{k}': ce[1], f'ZC{k}': ce[2], f'A{k}': ax[0], f'B{k}': ax[1], f'C{k}': ax[2]})
    else:
        ell = s['ell']
        c['obs_param'] = np.array([float(x) for e in ell for x in e[:2]] + [float(x) for e in ell for x in e[2:]])
        for k, e in enumerate(ell, 1):
            c.update({f'XC{k}': e[0], f'YC{k}': e[1], f'A{k}': e[2], f'B{k}': e[3]})
    if system_id == 'car_park':
        L, W = 4.35, 2
        c.update(L=L, W=W, L_delta=2.63, tau_delta=1, k_db=50, delta_bound=_PI / 3, w_delta_bound=0)
        run.append(0)
        ter.append(0)
        c['check_points_BF'] = np.array([[-L / 2, W / 2], [-L / 2 + L / 3, W / 2], [-L / 2 + 2 / 3 * L, W / 2], [L / 2, W / 2],
                                         [L / 2, 0], [L / 2, -W / 2], [-L / 2 + 2 / 3 * L, -W / 2], [-L / 2 + L / 3, -W / 2],
                                         [-L / 2, -W / 2], [-L / 2, 0]])
    c['cost_weights_running'] = np.array(run, dtype=float)
    c['cost_weights_terminal'] = np.array(ter, dtype=float)
    c['soft_max_param'] = np.array([c['alpha'], c['alpha2']])
    c['cost_funct_param'] = np.array([c['offset_cost_fun'], c['scale_cost_fun']])
    c['TARGET_STATE'] = np.array(s['target'], dtype=float)

    # state / action spaces
    nb_state = s['nx'] + 1
    c['nb_state'] = nb_state
    c['nb_action'] = s['na']
    t_hi = (c['NSTEPS'] - 1) * c['dt']
    c['x_init_min'] = np.array(list(s['init_lo']) + [s['t_min']], dtype=float)
    c['x_init_max'] = np.array(list(s['init_hi']) + [t_hi], dtype=float)
    c['x_min'] = np.array([-np.inf] * s['nx'] + [s['t_min']])
    c['x_max'] = np.array([np.inf] * nb_state)
    c['state_norm_arr'] = np.array(list(s['norm']) + [int(c['NSTEPS'] * c['dt'])])
    c['u_min'] = np.array(s['u_lo'])
    c['u_max'] = np.array(s['u_hi'])
    c['tau_coulomb_max'] = np.zeros(s['na'])
    c['end_effector_frame_id'] = 'EE'
    c.update(overrides)
    return SimpleNamespace(**c)
