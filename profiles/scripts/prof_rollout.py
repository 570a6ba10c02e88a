"""One bench-size rollout launch per engine for ncu: python profiles/scripts/prof_rollout.py [engine] [B]"""
import sys; sys.path.insert(0, '/root/repo')
import numpy as np, torch
from cacto_b200.conf import get_conf
from cacto_b200 import environment as genv
from cacto_b200.NeuralNetwork import NN
from cacto_b200.RL import RL_AC
engine = sys.argv[1] if len(sys.argv) > 1 else 'tc'
B = int(sys.argv[2]) if len(sys.argv) > 2 else 131072
conf = get_conf('manipulator'); env = genv.make_env(conf)
rl = RL_AC(env, NN(env, conf, 1e-2, seed=0), conf, 0); rl.setup_model()
rng = np.random.default_rng(1000)
X0 = rng.uniform(conf.x_init_min, conf.x_init_max, (B, conf.nb_state)); X0[:, -1] = 0.0
ics = torch.as_tensor(X0).cuda(); T = conf.NSTEPS
hz = torch.full((B,), T, dtype=torch.int32, device='cuda')
states = torch.empty((T + 1, conf.nb_state, B), dtype=torch.float64, device='cuda')
controls = torch.empty((T, conf.nb_action, B), dtype=torch.float64, device='cuda')
flags = torch.empty(B, dtype=torch.int32, device='cuda')
for _ in range(2):
    rl._launch_rollout(1, ics, hz, T, states, controls, flags, None, B, engine)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record(); rl._launch_rollout(1, ics, hz, T, states, controls, flags, None, B, engine); b.record(); torch.cuda.synchronize()
print(engine, B, 'ms', a.elapsed_time(b), 'ok', bool(flags.all()), flush=True)
