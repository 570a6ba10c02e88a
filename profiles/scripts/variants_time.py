import sys
import numpy as np, torch
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
from test_gpu_nn_variants import make
for ct in ('sine', 'elu', 'sine-elu', 'relu'):
    for B in (64, 512):
        try:
            conf, env, nn, rl, batch = make('manipulator', B, critic_type=ct)
        except TypeError:
            from test_gpu_nn import make as mk
            conf, env, nn, rl, batch = mk('manipulator', B)
        s, pr, sn, dv, d, term, w = batch
        ug = rl.make_update_graph(B)
        for k_, t_ in zip(('state', 'state_next', 'partial_rtg', 'dVdx', 'done', 'term', 'weights'), (s, sn, pr, dv, d, term, w)):
            ug.io[k_].copy_(torch.as_tensor(t_))
        for _ in range(5): ug.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(50): ug.replay()
        e1.record(); torch.cuda.synchronize()
        print(ct, B, 'graph update %.1f us' % (e0.elapsed_time(e1) * 1e3 / 50), flush=True)
