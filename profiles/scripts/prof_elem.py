"""One launch of each element-wise kernel (manipulator, B = 4 Mi samples) for ncu: python profiles/scripts/prof_elem.py"""
import sys; sys.path.insert(0, '/root/repo')
import torch
from cacto_b200.conf import get_conf
from cacto_b200 import environment as genv
B = 1 << 22
conf = get_conf('manipulator'); env = genv.make_env(conf); ns, na = conf.nb_state, conf.nb_action
for dt in (torch.float32, torch.float64):
    s = torch.rand((B, ns), device='cuda', dtype=dt) * 2 - 1; a = torch.rand((B, na), device='cuda', dtype=dt) * 2 - 1
    for _ in range(2):
        env.simulate_batch(s, a); env.derivative_batch(s, a); env.augmented_derivative_batch(s, a)
    torch.cuda.synchronize()
print('ok')
