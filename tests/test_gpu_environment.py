"""CUDA system models (through the C ABI) against the oracle and the reference-generated goldens.
Tolerances are BASELINE.md's: 1e-6 relative in fp64, 1e-5 relative in fp32 (relative to the largest
magnitude of the compared block, since Jacobians hold structural zeros)."""
import numpy as np
import pytest
import torch

from cacto_b200.conf import SYSTEM_IDS, get_conf
from conftest import golden
from oracle import systems as osys

pytestmark = pytest.mark.gpu
RTOL = {torch.float64: 1e-6, torch.float32: 1e-5}


def _close(a, b, rtol):
    a = a.detach().cpu().double().numpy() if isinstance(a, torch.Tensor) else np.asarray(a, dtype=float)
    b = np.asarray(b, dtype=float)
    assert a.shape == b.shape, (a.shape, b.shape)
    scale = max(np.abs(b).max(), 1e-30)
    err = np.abs(a - b).max() / scale
    assert err <= rtol, f'rel err {err:.3e} > {rtol}'


def _samples(conf, n, seed):
    rng = np.random.default_rng(seed)
    lo, hi = np.asarray(conf.x_init_min, float), np.asarray(conf.x_init_max, float)
    s = rng.uniform(lo, hi, (n, conf.nb_state))
    if conf.system_id == 'car_park':
        s[:, 3] = rng.uniform(-3, 3, n)
        s[:, 4] = rng.uniform(-0.5, 0.5, n)
    s[:, -1] = conf.dt * np.round(s[:, -1] / conf.dt)
    a = rng.uniform(np.asarray(conf.u_min, float), np.asarray(conf.u_max, float), (n, conf.nb_action))
    return s, a


@pytest.fixture(scope='module', params=SYSTEM_IDS)
def pair(request):
    from cacto_b200 import environment as genv
    conf = get_conf(request.param)
    return conf, genv.make_env(conf), osys.make_env(conf)


@pytest.mark.parametrize('dtype', [torch.float64, torch.float32])
def test_simulate_and_derivative_batch(pair, dtype):
    conf, env, ora = pair
    n = 8 if conf.system_id == 'ur5' else 300          # 300: more than one CTA + a ragged tail
    s, a = _samples(conf, n, 1)
    s, a = s.astype(np.float32), a.astype(np.float32)   # quirk Q13: inputs are f32-rounded
    S = torch.tensor(s, dtype=dtype, device='cuda')
    A = torch.tensor(a, dtype=dtype, device='cuda')
    nxt = env.simulate_batch(S, A)
    assert nxt.dtype == dtype and nxt.shape == (n, conf.nb_state)
    ref = np.array([ora.simulate(x.astype(np.float64), u.astype(np.float64)) for x, u in zip(s, a)])
    # velocity increments dt*acc sit next to O(1) states: compare per column
    for j in range(conf.nb_state):
        _close(nxt[:, j], ref[:, j], RTOL[dtype])
    Fu = env.derivative_batch(S, A)
    assert Fu.shape == (n, conf.nb_state, conf.nb_action)
    refFu = np.array([ora.derivative(x.astype(np.float64), u.astype(np.float64)) for x, u in zip(s, a)])
    _close(Fu, refFu, RTOL[dtype])


@pytest.mark.parametrize('dtype', [torch.float64, torch.float32])
def test_augmented_derivative_batch(pair, dtype):
    conf, env, ora = pair
    n = 6 if conf.system_id == 'ur5' else 200
    s, a = _samples(conf, n, 2)
    Fx, Fu = env.augmented_derivative_batch(torch.tensor(s, dtype=dtype, device='cuda'), torch.tensor(a, dtype=dtype, device='cuda'))
    rx, ru = zip(*[ora.augmented_derivative(x, u) for x, u in zip(s, a)])
    rx, ru = np.array(rx), np.array(ru)
    I = np.eye(conf.nx)
    # the interesting part of Fx is dt * d(acc)/dx next to the identity: compare Fx - I
    _close((Fx.double().cpu().numpy() - I), rx - I, 20 * RTOL[dtype] if dtype == torch.float32 else RTOL[dtype])
    _close(Fu, ru, RTOL[dtype])


def test_single_sample_api_matches_oracle(pair):
    conf, env, ora = pair
    s, a = _samples(conf, 3, 3)
    for x, u in zip(s, a):
        np.testing.assert_allclose(env.simulate(x, u), ora.simulate(x, u), rtol=1e-9, atol=1e-12)
        np.testing.assert_allclose(env.derivative(x, u), ora.derivative(x, u), rtol=1e-7, atol=1e-14)
        fx, fu = env.augmented_derivative(x[:-1], u)
        rx, ru = ora.augmented_derivative(x, u)
        np.testing.assert_allclose(fx, rx, rtol=1e-6, atol=1e-9)
        np.testing.assert_allclose(fu, ru, rtol=1e-7, atol=1e-14)
        np.testing.assert_allclose(env.get_end_effector_position(x), ora.get_end_effector_position(x), rtol=1e-10, atol=1e-12)
        for w in (conf.cost_weights_running, conf.cost_weights_terminal):
            assert env.reward(w, x, u) == pytest.approx(ora.reward(w, x, u), rel=1e-8, abs=1e-12)
            assert env.reward(w, x) == pytest.approx(ora.reward(w, x), rel=1e-8, abs=1e-12)
        nxt, r = env.step(conf.cost_weights_running, x, u)
        assert r == pytest.approx(ora.reward(conf.cost_weights_running, x, u), rel=1e-8, abs=1e-12)


def test_against_reference_generated_goldens(pair):
    """Outputs of the reference's own environment.py (tests/golden/env_<system>.npz)."""
    conf, env, _ = pair
    g = golden(f'env_{conf.system_id}.npz')
    S, A, W = g['states'], g['actions'], g['weights']
    Sd, Ad = torch.tensor(S, device='cuda'), torch.tensor(A, device='cuda')
    if 'simulate' in g:
        _close(env.simulate_batch(Sd, Ad), g['simulate'], 1e-12)
        _close(env.derivative_batch(Sd, Ad), g['derivative'], 1e-12)
        Fx, Fu = env.augmented_derivative_batch(Sd, Ad)
        _close(Fx, g['Fx'], 1e-12)
        _close(Fu, g['Fu'], 1e-12)
        _close(env.get_end_effector_position_batch(Sd), g['ee'], 1e-12)
        _close(env.simulate_batch(Sd.float(), Ad.float()), g['simulate_batch'], 1e-5)
    else:
        _close(env.get_end_effector_position_batch(Sd), g['ee_injected'], 1e-12)
    r_sa = np.array([env.reward(w, s, a) for w, s, a in zip(W, S, A)])
    r_s = np.array([env.reward(w, s) for w, s in zip(W, S)])
    np.testing.assert_allclose(r_sa, g['reward_sa'], rtol=1e-7, atol=1e-12)
    np.testing.assert_allclose(r_s, g['reward_s'], rtol=1e-7, atol=1e-12)
    rb = env.reward_batch(W, Sd.float(), Ad.float())
    assert rb.shape == (len(S), 1)
    _close(rb, g['reward_batch'], 2e-5)


def test_reward_batch_gradient(pair):
    conf, env, ora = pair
    s, a = _samples(conf, 64, 4)
    term = (np.random.default_rng(0).uniform(size=(64, 1)) < 0.3).astype(float)
    W = term.dot(np.reshape(conf.cost_weights_terminal, [1, -1])) + (1 - term).dot(np.reshape(conf.cost_weights_running, [1, -1]))
    g = env.reward_batch_da(W, torch.tensor(s, device='cuda'), torch.tensor(a, device='cuda'))
    _close(g, ora.reward_batch_da(W, a), 1e-9)


def test_empty_batch_and_bad_shapes(pair):
    conf, env, _ = pair
    z = env.simulate_batch(torch.zeros((0, conf.nb_state), device='cuda'), torch.zeros((0, conf.nb_action), device='cuda'))
    assert z.shape == (0, conf.nb_state)
    with pytest.raises(ValueError):
        env.simulate_batch(torch.zeros((4, conf.nb_state + 1), device='cuda'), torch.zeros((4, conf.nb_action), device='cuda'))
