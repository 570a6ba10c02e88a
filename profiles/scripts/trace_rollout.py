"""Event timeline of CTA 0 of k_rollout_tc16 (debug build with -DT16_TRACE): python profiles/scripts/trace_rollout.py"""
import os, sys, ctypes; sys.path.insert(0, '/root/repo')
os.environ['CACTO_B200_LIB'] = '/root/repo/scratch/libcacto_trace.so'
import numpy as np, torch
from cacto_b200.conf import get_conf
from cacto_b200 import environment as genv
from cacto_b200.NeuralNetwork import NN
from cacto_b200.RL import RL_AC
B = 131072
conf = get_conf('manipulator'); env = genv.make_env(conf)
rl = RL_AC(env, NN(env, conf, 1e-2, seed=0), conf, 0); rl.setup_model()
rng = np.random.default_rng(1000)
X0 = rng.uniform(conf.x_init_min, conf.x_init_max, (B, conf.nb_state)); X0[:, -1] = 0.0
ics = torch.as_tensor(X0).cuda(); T = conf.NSTEPS
hz = torch.full((B,), T, dtype=torch.int32, device='cuda')
states = torch.empty((T + 1, conf.nb_state, B), dtype=torch.float64, device='cuda')
controls = torch.empty((T, conf.nb_action, B), dtype=torch.float64, device='cuda')
flags = torch.empty(B, dtype=torch.int32, device='cuda')
rl._launch_rollout(1, ics, hz, T, states, controls, flags, None, B, 'tc'); torch.cuda.synchronize()
lib = ctypes.CDLL(os.environ['CACTO_B200_LIB'])
n = lib.cacto_debug_t16_trace_n()
buf = torch.zeros(5 * n, dtype=torch.int64, device='cuda')
lib.cacto_debug_t16_trace.argtypes = [ctypes.c_void_p]
assert lib.cacto_debug_t16_trace(buf.data_ptr()) == 0
rl._launch_rollout(1, ics, hz, T, states, controls, flags, None, B, 'tc'); torch.cuda.synchronize()
ev = buf.cpu().numpy().reshape(5, n)
np.save('/root/repo/gpurun_out/t16_trace.npy', ev)
print('saved', [(int((ev[r] != 0).sum())) for r in range(5)])
