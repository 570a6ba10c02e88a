import sys, time
import numpy as np, torch
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
from test_gpu_nn import make
from cacto_b200.replay_buffer import ReplayBuffer
from cacto_b200.ops import ops
B = 64
conf, env, nn, rl, batch = make('manipulator', B)
ns = conf.nb_state
buf = ReplayBuffer(conf)
buf.add_rows(torch.randn((conf.REPLAY_SIZE + 8, 3 * ns + 3), dtype=torch.float64, device='cuda'))
ug = rl.make_update_graph(B)
N = 300
def timeit(fn):
    for _ in range(20): fn()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(N): fn()
    h = time.perf_counter() - t0
    torch.cuda.synchronize(); tot = time.perf_counter() - t0
    return h / N * 1e6, tot / N * 1e6
idx_np = np.random.randint(0, 65536, size=B)
print('np.random.randint           host %.1f us' % timeit(lambda: np.random.randint(0, 65536, size=B))[0])
print('as_tensor + .to(cuda)        host %.1f us  total %.1f' % timeit(lambda: torch.as_tensor(idx_np).to('cuda', non_blocking=True)))
idx_dev = torch.as_tensor(idx_np).to('cuda')
print('_gather (op)                 host %.1f us  total %.1f' % timeit(lambda: buf._gather(idx_dev, ug.io)))
print('weights.fill_                host %.1f us  total %.1f' % timeit(lambda: ug.io['weights'].fill_(1.0)))
print('buf.sample(out)              host %.1f us  total %.1f' % timeit(lambda: buf.sample(out=ug.io)))
print('ug.replay()                  host %.1f us  total %.1f' % timeit(lambda: ug.replay()))
def full():
    buf.sample(out=ug.io); ug.replay(); return float(nn.last_critic_loss)
print('sample + replay + loss read  host %.1f us  total %.1f' % timeit(full))
