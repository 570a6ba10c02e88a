import sys
import numpy as np, torch
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
from test_gpu_nn import make
system, B, eng = sys.argv[1], int(sys.argv[2]), sys.argv[3]
conf, env, nn, rl, batch = make(system, B)
s, pr, sn, dv, d, term, w = batch
nn.update_engine = eng
s, sn, pr, dv, d, w = [torch.as_tensor(x, device='cuda') for x in (s, sn, pr, dv, d, w)]
term = torch.as_tensor(term, device='cuda')
for _ in range(3):
    rl.update(s, sn, pr, dv, d, term, w, fuse_target=True)
torch.cuda.synchronize()
print('done')
