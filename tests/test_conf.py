"""cacto_b200.conf against the dump of the reference's conf_<system>.py modules."""
import json
import os

import numpy as np
import pytest

from cacto_b200.conf import SYSTEM_IDS, get_conf
from conftest import GOLDEN

REF = json.load(open(os.path.join(GOLDEN, 'conf_constants.json')))


def _num(v):
    if isinstance(v, str):
        return {'inf': np.inf, '-inf': -np.inf}.get(v, v)
    if isinstance(v, list):
        return [_num(x) for x in v]
    return v


@pytest.mark.parametrize('system', SYSTEM_IDS)
def test_conf_matches_reference_dump(system):
    conf = get_conf(system)
    ref = REF[system]
    skipped = {'weight', 'test_set', 'test_set_rec', 'N_try_rec', 'update_step_counter_rec', 'URDF_FILENAME', 'q_init', 'v_init',
               'plot_flag', 'plot_rollout_interval', 'plot_rollout_interval_diff_loc', 'x_base', 'y_base', 'x_des', 'y_des',
               'z_des', 'tau_lower_bound', 'tau_upper_bound', 'use_viewer', 'simulate_real_time', 'show_floor', 'PRINT_T',
               'DISPLAY_T', 'bound_actions', 'omega_lower_bound', 'omega_upper_bound', 'jerk_lower_bound', 'jerk_upper_bound',
               'acc_lower_bound', 'acc_upper_bound', 'delta_dot_lower_bound', 'delta_dot_upper_bound', 'ell1_center',
               'ell2_center', 'ell3_center'}
    checked = 0
    for k, v in ref.items():
        if k in skipped:
            continue
        assert hasattr(conf, k), f'{system}: missing {k}'
        mine = getattr(conf, k)
        v = _num(v)
        if isinstance(v, (list, float, int)) and not isinstance(v, bool) and v is not None:
            np.testing.assert_array_equal(np.asarray(mine, dtype=float), np.asarray(v, dtype=float), err_msg=f'{system}.{k}')
        else:
            assert mine == v, f'{system}.{k}: {mine!r} != {v!r}'
        checked += 1
    assert checked > 60


def test_overrides_of_primary_constants_reach_derived_values():
    """conf_*.py derive the LR-schedule boundaries from REPLAY_SIZE / BATCH_SIZE, nsteps_TD_N, x_init_max[-1] and
    state_norm_arr[-1] from NSTEPS and dt: an override of a primary constant must propagate (BASELINE configs 2 / 3 / 5 override
    BATCH_SIZE)."""
    from cacto_b200.conf import get_conf
    a, b = get_conf('manipulator'), get_conf('manipulator', BATCH_SIZE=4096)
    assert a.boundaries_schedule_LR_C[0] == 200 * 2 ** 16 / 64
    assert b.boundaries_schedule_LR_C == [m * 2 ** 16 / 4096 for m in (200, 300, 400, 500)]
    c = get_conf('car', NSTEPS=200, dt=0.1, REPLAY_SIZE=2 ** 12)
    assert c.nsteps_TD_N == 50 and c.x_init_max[-1] == 199 * 0.1 and c.state_norm_arr[-1] == 20
    assert c.boundaries_schedule_LR_A[0] == 200 * 2 ** 12 / 64
    d = get_conf('ur5', nsteps_TD_N=7)                       # a derived value can still be overridden directly
    assert d.nsteps_TD_N == 7
