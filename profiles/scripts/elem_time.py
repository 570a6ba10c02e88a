"""CUDA-event timing of the element-wise kernels (K1' step, K2 derivative / augmented) for every system at B = 4 Mi samples
(UR5 256 Ki), fp32 and fp64, through the Python mirror: python profiles/scripts/elem_time.py"""
import sys; sys.path.insert(0, '/root/repo')
import torch
from cacto_b200.conf import get_conf
from cacto_b200 import environment as genv

def timed(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n * 1e-3

for system in ('manipulator', 'car', 'single_integrator', 'double_integrator', 'ur5'):
    conf = get_conf(system); env = genv.make_env(conf); ns, na, nx = conf.nb_state, conf.nb_action, conf.nb_state - 1
    B = 1 << 18 if system == 'ur5' else 1 << 22
    for dt, sz in ((torch.float32, 4), (torch.float64, 8)):
        s = torch.rand((B, ns), device='cuda', dtype=dt) * 2 - 1; a = torch.rand((B, na), device='cuda', dtype=dt) * 2 - 1
        t_step = timed(lambda: env.simulate_batch(s, a))
        t_der = timed(lambda: env.derivative_batch(s, a))
        t_aug = timed(lambda: env.augmented_derivative_batch(s, a))
        by_step, by_der, by_aug = sz * (2 * ns + na), sz * (ns + na + ns * na), sz * (ns + na + nx * nx + nx * na)
        print(f'{system:18s} {"f32" if sz == 4 else "f64"}  step {t_step*1e6:7.1f} us {B*by_step/t_step/1e9:7.0f} GB/s | derivative {t_der*1e6:7.1f} us '
              f'{B*by_der/t_der/1e9:7.0f} GB/s | augmented {t_aug*1e6:7.1f} us {B*by_aug/t_aug/1e9:7.0f} GB/s', flush=True)
