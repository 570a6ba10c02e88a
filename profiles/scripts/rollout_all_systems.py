"""Rollout throughput of every system at its BASELINE config size (full horizon), engines tc / fma."""
import sys; sys.path.insert(0, '/root/repo')
import numpy as np, torch
from cacto_b200.conf import get_conf
from cacto_b200 import environment as genv
from cacto_b200.NeuralNetwork import NN
from cacto_b200.RL import RL_AC
for sysid, B in (('single_integrator', 65536), ('double_integrator', 65536), ('car', 262144), ('car_park', 262144), ('manipulator', 131072), ('ur5', 32768)):
    conf = get_conf(sysid); env = genv.make_env(conf)
    rl = RL_AC(env, NN(env, conf, 1e-2, seed=0), conf, 0); rl.setup_model()
    rng = np.random.default_rng(1)
    X0 = rng.uniform(conf.x_init_min, conf.x_init_max, (B, conf.nb_state)); X0[:, -1] = 0.0
    ics = torch.as_tensor(X0).cuda(); T = conf.NSTEPS
    hz = torch.full((B,), T, dtype=torch.int32, device='cuda')
    states = torch.empty((T + 1, conf.nb_state, B), dtype=torch.float64, device='cuda')
    controls = torch.empty((T, conf.nb_action, B), dtype=torch.float64, device='cuda')
    flags = torch.empty(B, dtype=torch.int32, device='cuda')
    res = {}
    for engine in ('tc', 'fma'):
        for _ in range(2):
            rl._launch_rollout(1, ics, hz, T, states, controls, flags, None, B, engine)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(3):
            rl._launch_rollout(1, ics, hz, T, states, controls, flags, None, B, engine)
        b.record(); torch.cuda.synchronize()
        res[engine] = a.elapsed_time(b) / 3
    print(f'{sysid:18s} B={B:7d} T={T:3d}: tc {res["tc"]:8.2f} ms = {B*T/res["tc"]/1e6:7.2f} G env-steps/s | fma {res["fma"]:8.2f} ms = {B*T/res["fma"]/1e6:6.2f} G/s | ok {bool(flags.all())}', flush=True)
