"""One small pass over the hand-rolled kernels for compute-sanitizer (memcheck / racecheck / synccheck / initcheck, ONE tool per
run): K1 on the tcgen05 engines, K3 on the fp32-FMA and the tcgen05 engine (+ Adam / Polyak), K4 PER round, K5 reward-to-go.
    compute-sanitizer --tool memcheck python profiles/scripts/sanitize_case.py [k1|k3|k3tc|k4k5|all]"""
import sys
import numpy as np
import torch
sys.path.insert(0, '.')
sys.path.insert(0, 'tests')
from types import SimpleNamespace
from test_gpu_nn import make

which = sys.argv[1] if len(sys.argv) > 1 else 'all'
conf, env, nn, rl, batch = make('manipulator', 256)
s, pr, sn, dv, d, term, w = batch
rng = np.random.default_rng(0)
if which in ('k1', 'all'):
    X0 = rng.uniform(conf.x_init_min, conf.x_init_max, (300, conf.nb_state))
    X0[:, -1] = conf.dt * np.round(X0[:, -1] / conf.dt)
    X0[:, -1] = np.maximum(X0[:, -1], (conf.NSTEPS - 12) * conf.dt)          # short horizons: the sanitizer slows kernels ~50x
    for eng in ('tc', 'tf32', 'fma'):
        out = rl.rollout_batch(X0, 1, engine=eng)
        assert out['success'].cpu().numpy().all()
    print('k1 ok')
if which in ('k3', 'all'):
    nn.update_engine = 'fma'
    rl.update(s[:64], sn[:64], pr[:64], dv[:64], d[:64], term[:64], w[:64], fuse_target=True)
    rl.update(s, sn, pr, dv, d, term, w, fuse_target=True)
    torch.cuda.synchronize()
    print('k3 fma ok')
if which in ('k3tc', 'all'):
    nn.update_engine = 'tc'
    rl.update(s, sn, pr, dv, d, term, w, fuse_target=True)
    rl.update(s[:130], sn[:130], pr[:130], dv[:130], d[:130], term[:130], w[:130], fuse_target=True)
    torch.cuda.synchronize()
    print('k3 tc ok')
if which in ('k4k5', 'all'):
    from cacto_b200.replay_buffer import PrioritizedReplayBuffer
    ns = conf.nb_state
    bc = SimpleNamespace(REPLAY_SIZE=4096, BATCH_SIZE=256, nb_state=ns, prioritized_replay_alpha=0.6, prioritized_replay_beta=0.6,
                         prioritized_replay_eps=1e-2, fresh_factor=0.95)
    pb = PrioritizedReplayBuffer(bc)
    rows = rng.normal(size=(3000, 3 * ns + 3))
    cols = (rows[:, :ns], rows[:, ns], rows[:, ns + 1:2 * ns + 1], rows[:, 2 * ns + 1:3 * ns + 1], rows[:, 3 * ns + 1], rows[:, 3 * ns + 2])
    pb.add(*[(c,) for c in cols])
    for _ in range(2):
        o = pb.sample()
        pb.update_priorities(o[7], torch.randn(256, 1, device='cuda'), torch.randn(256, 1, device='cuda'))
    st = [rng.normal(size=(T + 1, ns)) for T in (100, 37, 1)]
    cost = [rng.uniform(0, 2, len(x)) for x in st]
    rl.rtg_batch(st, cost)
    torch.cuda.synchronize()
    print('k4 k5 ok')
print('done')
