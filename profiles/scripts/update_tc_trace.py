import os, sys, ctypes as C
os.environ['CACTO_B200_LIB'] = '/root/repo/scratch/libcacto_trace.so'
import numpy as np, torch
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
from test_gpu_nn import make
from cacto_b200 import _lib
conf, env, nn, rl, batch = make('manipulator', 16384)
s, pr, sn, dv, d, term, w = batch
nn.update_engine = 'tc'
s, sn, pr, dv, d, w = [torch.as_tensor(x, device='cuda') for x in (s, sn, pr, dv, d, w)]
term = torch.as_tensor(term, device='cuda')
for _ in range(2):
    rl.update(s, sn, pr, dv, d, term, w, fuse_target=True)
torch.cuda.synchronize()
lib = C.CDLL(os.environ['CACTO_B200_LIB'])
N = lib.cacto_debug_tcu_trace_n()
buf = torch.zeros(2 * N, dtype=torch.int64, device='cuda')
lib.cacto_debug_tcu_trace(C.c_void_p(buf.data_ptr()))
# one critic gradient only: fwd (no events) then bwd<0> (events), adj (issuer events only), bwd<2> (events)
import cacto_b200._lib as L
which = sys.argv[1] if len(sys.argv) > 1 else 'critic'
if which == 'fwd':
    V, dV = None, None
    # forward-only: one FP-kind sweep through the actor gradient would also run others; call the critic grad with w_S = 0 (fwd + B + wgrad) and read role 0 before B overwrites: B has no FEV codes >= 30
    nn.w_S = 0.0
g = nn.compute_critic_grad(rl.critic_model, rl.target_critic, s, sn, pr, dv, d, w)
torch.cuda.synchronize()
t = buf.cpu().numpy().reshape(2, N)
for role in (0, 1):
    ev = [(int(x) >> 8, int(x) & 255) for x in t[role] if x != 0]
    print('role', role, 'events', len(ev))
    t0 = ev[0][0] if ev else 0
    prev = t0
    for c, code in ev[:120]:
        print('  %8d (+%6d) code %d' % (c - t0, c - prev, code))
        prev = c
