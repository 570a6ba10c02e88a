import sys, time
import numpy as np, torch
sys.path.insert(0, '/root/repo')
from cacto_b200.conf import get_conf
from cacto_b200 import environment as genv
from cacto_b200.TO import TO_Casadi
for system, E in (('ur5', 256), ('manipulator', 256), ('car', 256)):
    conf = get_conf(system); env = genv.make_env(conf)
    T = conf.NSTEPS; nx, na = conf.nb_state - 1, conf.nb_action
    rng = np.random.default_rng(0)
    S = [rng.uniform(np.asarray(conf.x_init_min, float)[:nx], np.asarray(conf.x_init_max, float)[:nx], (T + 1, nx)) * 0.3 for _ in range(E)]
    U = [rng.uniform(-1, 1, (T, na)) for _ in range(E)]
    to = TO_Casadi(env, conf, None, 1e-2)
    for _ in range(2): to.backward_pass_batch(S, U)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(3): to.backward_pass_batch(S, U)
    torch.cuda.synchronize()
    print(system, E, 'trajectories x', T, ': %.1f ms per call (host staging included)' % ((time.perf_counter() - t0) / 3 * 1e3), flush=True)
