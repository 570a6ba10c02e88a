"""Env.reset (environment.py:46-55): the host RNG stream of the reference -- random.uniform for the time first, then one draw
per state component, time snapped to the dt grid.  Pinned by the initial conditions the reference's own reset() produced for
tests/golden/toinit_cases.npz (make_golden.py::toinit_goldens: random.seed(9), three resets per system)."""
import random

import numpy as np
import pytest

from cacto_b200.conf import SYSTEM_IDS, get_conf
from conftest import golden
from oracle import systems as osys


def _env(system):
    from cacto_b200 import environment as genv
    return genv.make_env(get_conf(system))


@pytest.mark.parametrize('system', ['single_integrator', 'car', 'car_park'])
def test_reset_reproduces_reference_initial_conditions(system):
    g = golden('toinit_cases.npz')
    env = _env(system)
    random.seed(9)
    for k in range(3):
        ics = env.reset()
        ref = g[f'{system}_{k}_0_ics']
        if k == 2:                       # the generator overwrote the time of the third case with 0
            np.testing.assert_array_equal(ics[:-1], ref[:-1])
        else:
            np.testing.assert_array_equal(ics, ref)


@pytest.mark.parametrize('system', SYSTEM_IDS)
def test_reset_matches_oracle_and_bounds(system):
    conf = get_conf(system)
    env, oenv = _env(system), osys.make_env(conf)
    random.seed(4)
    a = [env.reset() for _ in range(50)]
    random.seed(4)
    b = [oenv.reset() for _ in range(50)]
    np.testing.assert_array_equal(np.array(a), np.array(b))
    a = np.array(a)
    assert (a >= np.asarray(conf.x_init_min) - 1e-12).all() and (a <= np.asarray(conf.x_init_max) + 1e-12).all()
    k = a[:, -1] / conf.dt
    assert np.abs(k - np.round(k)).max() < 1e-9              # time on the dt grid
