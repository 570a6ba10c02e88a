import sys; sys.path.insert(0,'/root/repo')
import numpy as np, torch
from cacto_b200.conf import get_conf
from cacto_b200 import environment as genv
from cacto_b200.NeuralNetwork import NN
from cacto_b200.RL import RL_AC
B=int(sys.argv[1]) if len(sys.argv)>1 else 64
conf=get_conf('manipulator'); env=genv.make_env(conf); nn=NN(env,conf,1e-2,seed=0); rl=RL_AC(env,nn,conf,0); rl.setup_model()
rng=np.random.default_rng(0); ns=7
s=rng.uniform(conf.x_init_min,conf.x_init_max,(B,ns)).astype(np.float32); sn=rng.uniform(conf.x_init_min,conf.x_init_max,(B,ns)).astype(np.float32)
pr=rng.uniform(-5,0,(B,1)).astype(np.float32); dv=rng.normal(size=(B,ns)).astype(np.float32); dv[:,-1]=0
d=(rng.uniform(size=(B,1))<0.5).astype(np.float32); term=(rng.uniform(size=(B,1))<0.1).astype(np.float64); w=np.ones((B,1),np.float32)
t=lambda a: torch.tensor(a,device='cuda')
args=(t(s),t(sn),t(pr),t(dv),t(d),t(term),t(w))
for _ in range(10): rl.update(*args, fuse_target=True)
torch.cuda.synchronize()
a,b=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(50): rl.update(*args, fuse_target=True)
b.record(); torch.cuda.synchronize()
print('B',B,'us/update eager',a.elapsed_time(b)*1000/50)
g=rl.make_update_graph(B)
for k_,t_ in zip(('state','state_next','partial_rtg','dVdx','done','term','weights'),args): g.io[k_].copy_(t_)
for _ in range(20): g.replay()
torch.cuda.synchronize()
a.record()
for _ in range(500): g.replay()
b.record(); torch.cuda.synchronize()
print('B',B,'us/update graph',a.elapsed_time(b)*1000/500)
