"""WarmStartFeeder host logic with an injected rollout function (no GPU needed) and, on the GPU box, the real fused rollout."""
import numpy as np
import pytest

from cacto_b200.conf import get_conf


def _to_solve(ics, states, controls, T):          # stands in for TO_Casadi.TO_Solve: must see exactly the warm-start of `ics`
    assert states.shape[0] == T + 1 and controls.shape[0] == T and states.flags['C_CONTIGUOUS']
    return float(ics[0]), float(states[:, 0].sum()), float(controls.sum()), T


def test_feeder_orders_results_and_skips_dead_episodes():
    from cacto_b200.feeder import WarmStartFeeder
    conf = get_conf('single_integrator')
    ns, na, Tm = conf.nb_state, conf.nb_action, conf.NSTEPS
    calls = []

    def fake_rollout(ics, ep):
        B = len(ics)
        calls.append(B)
        hz = np.array([0 if i % 7 == 3 else 1 + (i % 5) for i in range(B)], dtype=np.int32)
        ok = np.array([0 if i % 11 == 5 else 1 for i in range(B)], dtype=np.int32)
        st = np.zeros((Tm + 1, ns, B)); ct = np.zeros((Tm, na, B))
        for i in range(B):
            st[:, 0, i] = ics[i, 0] + np.arange(Tm + 1)
            ct[:, :, i] = ics[i, 0]
        return st, ct, ok, hz
    E = 50
    ICS = np.zeros((E, ns)); ICS[:, 0] = np.arange(E) * 10.0
    f = WarmStartFeeder(None, _to_solve, nb_cpus=2, chunk=16, rollout_fn=fake_rollout)
    res = f.run(ICS, ep=1)
    assert calls == [16, 16, 16, 2] and len(res) == E
    for e in range(E):
        i = e % 16
        T = 0 if i % 7 == 3 else 1 + (i % 5)
        if T == 0 or i % 11 == 5:
            assert res[e] is None
        else:
            x0 = ICS[e, 0]
            assert res[e] == (x0, float((x0 + np.arange(T + 1)).sum()), float(x0 * T * na), T)


@pytest.mark.gpu
@pytest.mark.parametrize('compact', [True, False])
def test_feeder_with_gpu_rollouts(compact):
    from cacto_b200 import environment as genv
    from cacto_b200.NeuralNetwork import NN
    from cacto_b200.RL import RL_AC
    from cacto_b200.feeder import WarmStartFeeder
    conf = get_conf('manipulator')
    env = genv.make_env(conf)
    rl = RL_AC(env, NN(env, conf, 1e-2, seed=0), conf, 0)
    rl.setup_model()
    rng = np.random.default_rng(0)
    ICS = rng.uniform(conf.x_init_min, conf.x_init_max, (300, conf.nb_state))
    ICS[:, -1] = conf.dt * np.round(ICS[:, -1] / conf.dt)
    ICS[7, -1] = conf.NSTEPS * conf.dt                        # horizon 0 -> skipped
    res = WarmStartFeeder(rl, _to_solve, nb_cpus=2, chunk=128, compact=compact).run(ICS, ep=1)
    ref = rl.rollout_batch(ICS, 1)
    S = ref['states'].permute(2, 0, 1).cpu().numpy(); C = ref['controls'].permute(2, 0, 1).cpu().numpy(); hz = ref['horizon'].cpu().numpy()
    assert res[7] is None
    for e in (0, 1, 127, 128, 299):
        T = int(hz[e])
        # a rollout's K-chunk accumulation order depends on which of the two tile pipelines runs it: equal to fp32 rounding only
        assert res[e][3] == T and res[e][1] == pytest.approx(S[e, :T + 1, 0].sum(), rel=1e-5, abs=1e-4) and res[e][2] == pytest.approx(C[e, :T].sum(), rel=1e-5, abs=1e-4)


@pytest.mark.gpu
def test_feeder_compact_transfers_hand_over_identical_warm_starts():
    """The compact PCIe format (default) is invisible to the TO workers: same arrays, bit for bit, as the full fp64 format."""
    from cacto_b200 import environment as genv
    from cacto_b200.NeuralNetwork import NN
    from cacto_b200.RL import RL_AC
    from cacto_b200.feeder import WarmStartFeeder
    conf = get_conf('manipulator')
    env = genv.make_env(conf)
    rl = RL_AC(env, NN(env, conf, 1e-2, seed=0), conf, 0)
    rl.setup_model()
    rng = np.random.default_rng(1)
    ICS = rng.uniform(conf.x_init_min, conf.x_init_max, (200, conf.nb_state))
    ICS[:, -1] = conf.dt * np.round(ICS[:, -1] / conf.dt)
    got = {}
    for compact in (True, False):
        f = WarmStartFeeder(rl, None, chunk=200, compact=compact)
        got[compact] = [t for t in f._tasks(ICS, f._gpu_rollout(ICS, 1))]
    for a, b in zip(got[True], got[False]):
        assert (a is None) == (b is None)
        if a is not None:
            assert a[4] == b[4] and np.array_equal(a[2], b[2]) and np.array_equal(a[3], b[3])
            assert a[2].dtype == np.float64 and a[3].dtype == np.float64 and a[2].flags['C_CONTIGUOUS'] and a[3].flags['C_CONTIGUOUS']
