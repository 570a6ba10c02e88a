"""Device time per update: sequential UpdateGraph vs PipelinedUpdateGraph (conf batch), and the e2e loop (sample + replay + loss read-back)."""
import sys, time; sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import numpy as np, torch
from test_gpu_update_graph import build, fill
from cacto_b200.replay_buffer import ReplayBuffer
for system, B in (('manipulator', 64), ('manipulator', 128), ('ur5', 64), ('manipulator', 1024)):
    conf, rl = build(system, BATCH_SIZE=B)
    buf = ReplayBuffer(conf); fill(buf, conf, 60000)
    gs, gp = rl.make_update_graph(), rl.make_pipelined_update_graph()
    buf.sample(out=gs.io); buf.sample(out=gp.ios[0]); buf.sample(out=gp.ios[1])
    def dev_time(g, n=400):
        for _ in range(20): g.replay()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(n): g.replay()
        b.record(); torch.cuda.synchronize()
        return a.elapsed_time(b) / n * 1e3
    ts, tp = dev_time(gs), dev_time(gp); gp.flush()
    def e2e(g, n=600):
        pin = torch.zeros(1).pin_memory()
        for _ in range(20): buf.sample(out=g.io); g.replay()
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for _ in range(n):
            buf.sample(out=g.io); g.replay(); pin.copy_(rl.NN.last_critic_loss, non_blocking=True)
        torch.cuda.synchronize()
        return (time.perf_counter() - t0) / n * 1e6
    es, ep = e2e(gs), e2e(gp); gp.flush()
    print(f'{system} B={B}: graph replay {ts:.1f} us sequential, {tp:.1f} us pipelined ({ts / tp:.2f} x); e2e loop {es:.1f} -> {ep:.1f} us per update', flush=True)
