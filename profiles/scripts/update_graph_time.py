# stable timing: the update as a CUDA graph (RL_AC.make_update_graph), manipulator
import sys, time
import numpy as np, torch
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
from test_gpu_nn import make
for B in [int(a) for a in sys.argv[1:]] or [4096, 16384]:
    conf, env, nn, rl, batch = make('manipulator', B)
    s, pr, sn, dv, d, term, w = batch
    ug = rl.make_update_graph(B)
    for k_, t_ in zip(('state', 'state_next', 'partial_rtg', 'dVdx', 'done', 'term', 'weights'), (s, sn, pr, dv, d, term, w)):
        ug.io[k_].copy_(torch.as_tensor(t_))
    for _ in range(10): ug.replay()
    torch.cuda.synchronize()
    best = 1e9
    for rep in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(50): ug.replay()
        e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) * 1e3 / 50)
    print('B=%d graph update %.1f us' % (B, best), flush=True)
