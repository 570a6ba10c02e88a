import sys, time; sys.path.insert(0,'/root/repo')
import numpy as np, torch
from cacto_b200.conf import get_conf
from cacto_b200 import environment as genv
from cacto_b200.NeuralNetwork import NN
from cacto_b200.RL import RL_AC
conf=get_conf('manipulator'); env=genv.make_env(conf); rl=RL_AC(env,NN(env,conf,1e-2,seed=0),conf,0); rl.setup_model()
B,T,ns,na=131072,100,7,3
rng=np.random.default_rng(0); X0=rng.uniform(conf.x_init_min,conf.x_init_max,(B,ns)); X0[:,-1]=0
ics=torch.as_tensor(X0).pin_memory()
st=torch.empty((T+1,ns,B),dtype=torch.float64).pin_memory(); ct=torch.empty((T,na,B),dtype=torch.float64).pin_memory(); fl=torch.empty(B,dtype=torch.int32).pin_memory()
for mode in ('pipelined','zero_copy','pipelined'):
    for _ in range(2): rl.rollout_to_host(ics,1,st,ct,fl,mode=mode)
    torch.cuda.synchronize(); t0=time.perf_counter()
    for _ in range(5): rl.rollout_to_host(ics,1,st,ct,fl,mode=mode)
    torch.cuda.synchronize(); dt=(time.perf_counter()-t0)/5
    print(mode, 'ms', dt*1e3, 'env-steps/s', B*T/dt, 'check', float(st[50,0,:10].sum()), int(fl.sum()))
