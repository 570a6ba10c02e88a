"""oracle.rtg pinned bit-exactly against RL_AC.RL_Solve / create_TO_init run in the build container."""
import numpy as np
import pytest
import torch
from types import SimpleNamespace

from cacto_b200.conf import get_conf
from conftest import golden
from oracle import nn as onn
from oracle import rtg, systems


def test_rl_solve_matches_reference():
    g = golden('rtg_cases.npz')
    for k in range(int(g['ncases'])):
        T, n, MC, ns = [int(x) for x in g[f'c{k}_meta']]
        conf = SimpleNamespace(nb_state=ns, MC=MC, nsteps_TD_N=n)
        st, partial, total, snext, done, rwrd, term, ret = rtg.rl_solve(conf, g[f'c{k}_states'], g[f'c{k}_cost'])
        np.testing.assert_array_equal(partial, g[f'c{k}_partial'])
        np.testing.assert_array_equal(total, g[f'c{k}_total'])
        np.testing.assert_array_equal(snext, g[f'c{k}_snext'])
        np.testing.assert_array_equal(done, g[f'c{k}_done'])
        np.testing.assert_array_equal(term, g[f'c{k}_term'])
        np.testing.assert_array_equal(rwrd, g[f'c{k}_rwrd'])
        assert ret == float(g[f'c{k}_ret'])


@pytest.mark.parametrize('system', ['single_integrator', 'car', 'car_park'])
def test_create_to_init_matches_reference(system):
    g = golden('toinit_cases.npz')
    conf = get_conf(system)
    env = systems.make_env(conf)
    ap = onn.to_torch([g[f'{system}_actor_{i}'] for i in range(6)])

    def actor_eval(x):
        with torch.no_grad():
            return onn.actor_forward(ap, torch.tensor(x, dtype=torch.float32), conf).numpy()[0]
    for k in range(3):
        for ep in (0, 1):
            ics = g[f'{system}_{k}_{ep}_ics']
            _, st, ct, T, ok = rtg.create_to_init(conf, env, actor_eval, ep, ics)
            assert ok == 1 and T == int(g[f'{system}_{k}_{ep}_T'])
            np.testing.assert_array_equal(ct, g[f'{system}_{k}_{ep}_controls'])
            np.testing.assert_array_equal(st, g[f'{system}_{k}_{ep}_states'])
