import sys; sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import numpy as np, torch
from test_gpu_nn import make
from cacto_b200.NeuralNetwork import Network
for B in (64, 128, 512):
    for mode in ('fused', 'generic'):
        conf, env, nn, rl, batch = make('manipulator', B)
        s, pr, sn, dv, d, term, w = batch
        if mode == 'generic':
            for name in ('critic_model', 'target_critic'):
                old = getattr(rl, name)
                new = Network('critic_generic', conf.nb_state, conf.nb_action, old.dims, ['sin'] * 4 + ['linear'])
                new.set_weights(old.get_weights())
                setattr(rl, name, new)
        args = [torch.as_tensor(a).cuda() for a in (s, sn, pr, dv, d, term, w)]
        for _ in range(10): rl.update(*args, fuse_target=True)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(100): rl.update(*args, fuse_target=True)
        b.record(); torch.cuda.synchronize()
        eager = a.elapsed_time(b) * 10
        g = rl.make_update_graph(B)
        for k_, t_ in zip(('state', 'state_next', 'partial_rtg', 'dVdx', 'done', 'term', 'weights'), args): g.io[k_].copy_(t_)
        for _ in range(20): g.replay()
        torch.cuda.synchronize(); a.record()
        for _ in range(500): g.replay()
        b.record(); torch.cuda.synchronize()
        print(f'B={B:4d} {mode:8s} eager {eager:7.1f} us/update, graph {a.elapsed_time(b)*2:7.1f} us/update', flush=True)
