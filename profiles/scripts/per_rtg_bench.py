import sys, time, random; sys.path.insert(0,'/root/repo')
import numpy as np, torch
from types import SimpleNamespace
from cacto_b200.replay_buffer import PrioritizedReplayBuffer, ReplayBuffer
from cacto_b200.rtg import rtg_batch
from oracle import per as oper, rtg as ortg
ns=7
rng=np.random.default_rng(0)
for B in (64, 4096):
    conf=SimpleNamespace(REPLAY_SIZE=2**16,BATCH_SIZE=B,nb_state=ns,prioritized_replay_alpha=0.6,prioritized_replay_beta=0.6,prioritized_replay_eps=1e-2,fresh_factor=0.95)
    rows=rng.normal(size=(2**16+8,3*ns+3))
    cols=(rows[:,:ns],rows[:,ns],rows[:,ns+1:2*ns+1],rows[:,2*ns+1:3*ns+1],rows[:,3*ns+1],rows[:,3*ns+2])
    pb=PrioritizedReplayBuffer(conf); pb.add(*[(c,) for c in cols])
    ob=oper.PrioritizedReplayBuffer(conf); t0=time.perf_counter(); ob.add(*[(c,) for c in cols]); t_add=time.perf_counter()-t0
    rtg=torch.randn(B,1,device='cuda'); V=torch.randn(B,1,device='cuda')
    random.seed(0)
    for _ in range(5):
        o=pb.sample(); pb.update_priorities(o[7],rtg,V)
    torch.cuda.synchronize(); n=50; t0=time.perf_counter()
    for _ in range(n):
        o=pb.sample(); pb.update_priorities(o[7],rtg,V)
    torch.cuda.synchronize(); gpu=(time.perf_counter()-t0)/n
    rn,Vn=rtg.cpu().numpy(),V.cpu().numpy()
    n2=3 if B>64 else 50; t0=time.perf_counter()
    for _ in range(n2):
        o=ob.sample(); ob.update_priorities(o[7],rn,Vn)
    cpu=(time.perf_counter()-t0)/n2
    print(f'PER round (sample+update_priorities) cap 65536 B={B}: GPU path {gpu*1e6:.0f} us/round, CPU oracle {cpu*1e6:.0f} us/round  (oracle add of 65536 rows {t_add:.2f} s)')
    # kernel-only timings
    idx=torch.randint(0,2**16,(B,),device='cuda'); val=torch.rand(B,device='cuda',dtype=torch.float64)+0.1
    from cacto_b200._lib import lib, ptr, stream_ptr
    def upd(): lib.cacto_segtree_update(ptr(pb._it_sum._value),ptr(pb._it_min._value),pb._capacity,ptr(idx),ptr(val),B,ptr(pb._stamp),stream_ptr())
    u=torch.rand(B,device='cuda',dtype=torch.float64); io=torch.empty(B,dtype=torch.int64,device='cuda'); lf=torch.empty(B,dtype=torch.float64,device='cuda')
    def smp(): lib.cacto_segtree_sample(ptr(pb._it_sum._value),ptr(pb._it_min._value),pb._capacity,2**16,ptr(u),B,ptr(io),ptr(lf),ptr(pb._totals),stream_ptr())
    for name,fn in (('segtree_update',upd),('segtree_sample',smp)):
        for _ in range(5): fn()
        a,b=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True); a.record()
        for _ in range(100): fn()
        b.record(); torch.cuda.synchronize(); print(f'   {name} kernel B={B}: {a.elapsed_time(b)*10:.1f} us')
conf=SimpleNamespace(nb_state=ns,MC=0,nsteps_TD_N=50)
for E in (200, 20000):
    states=torch.randn((E*101,ns),dtype=torch.float64,device='cuda'); cost=torch.rand(E*101,dtype=torch.float64,device='cuda')
    lens=[101]*E
    sl=[None]*E
    # host-list API cost included once; device tensor path for kernel timing
    out=rtg_batch(conf,[states[i*101:(i+1)*101].cpu().numpy() for i in range(min(E,200))],[cost[i*101:(i+1)*101].cpu().numpy() for i in range(min(E,200))])
    off=torch.arange(0,(E+1)*101,101,dtype=torch.int64,device='cuda')
    f64=dict(dtype=torch.float64,device='cuda'); tot=E*101
    o=[torch.empty(tot,**f64),torch.empty(tot,**f64),torch.empty((tot,ns),**f64),torch.empty(tot,**f64),torch.empty(tot,**f64),torch.empty(E,**f64)]
    rw=-cost
    def k(): lib.cacto_rtg_window(ptr(off),E,ptr(rw),ptr(states),ns,50,0,ptr(o[0]),ptr(o[1]),ptr(o[2]),ptr(o[3]),ptr(o[4]),ptr(o[5]),stream_ptr())
    for _ in range(3): k()
    a,b=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True); a.record()
    for _ in range(20): k()
    b.record(); torch.cuda.synchronize(); ms=a.elapsed_time(b)/20
    by=8*tot*(1+ns) + 8*tot*(ns+4)
    print(f'rtg_window E={E} T=100 n=50: {ms*1e3:.1f} us/launch, {E/ms*1e3:.3e} trajectories/s, {by/ms/1e6:.1f} GB/s algorithmic')
st=np.random.default_rng(1).normal(size=(101,ns)); c=np.random.default_rng(2).uniform(0,2,101)
t0=time.perf_counter()
for _ in range(20): ortg.rl_solve(conf,st,c)
print(f'CPU oracle RL_Solve T=100: {(time.perf_counter()-t0)/20*1e6:.0f} us/trajectory')

# ---- K6 backward pass: E TO solutions of T knots
from cacto_b200.conf import get_conf
from cacto_b200 import environment as genv
from cacto_b200.TO import TO_Casadi
from oracle import systems as osys, backward as obw
for sysid, E, T in (('manipulator', 256, 100), ('manipulator', 8192, 100), ('ur5', 256, 100), ('car', 256, 500)):
    conf = get_conf(sysid); env = genv.make_env(conf); oenv = osys.make_env(conf); to = TO_Casadi(env, conf)
    rng = np.random.default_rng(0)
    nx, na = conf.nb_state - 1, conf.nb_action
    X = [rng.uniform(np.asarray(conf.x_init_min[:-1], float), np.asarray(conf.x_init_max[:-1], float), (T + 1, nx)) for _ in range(E)]
    U = [rng.uniform(conf.u_min, conf.u_max, (T, na)) * 0.3 for _ in range(E)]
    to.backward_pass_batch(X, U); torch.cuda.synchronize()
    t0 = time.perf_counter(); Vx, off = to.backward_pass_batch(X, U); torch.cuda.synchronize(); gpu = time.perf_counter() - t0
    # kernel-only (data resident)
    from cacto_b200._lib import lib, ptr, stream_ptr
    K = E * (T + 1)
    Xd = torch.as_tensor(np.concatenate(X)).cuda(); Ud = torch.zeros((K, na), dtype=torch.float64, device='cuda'); offd = torch.as_tensor(off).cuda()
    ws = torch.empty(int(lib.cacto_backward_pass_workspace_bytes(nx, na, K)) // 8 + 1, dtype=torch.float64, device='cuda'); out = torch.zeros((K, nx + 1), dtype=torch.float64, device='cuda')
    def run(): lib.cacto_backward_pass(env._p, ptr(offd), E, ptr(Xd), ptr(Ud), K, 1e-9, ptr(ws), ptr(out), stream_ptr())
    run(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True); a.record(); run(); b.record(); torch.cuda.synchronize()
    t0 = time.perf_counter(); obw.backward_pass(oenv, 6, X[0][:6], U[0][:5]); cpu_knot = (time.perf_counter() - t0) / 6
    print(f'K6 backward pass {sysid} E={E} T={T}: host-to-device call {gpu*1e3:.1f} ms, kernels {a.elapsed_time(b):.2f} ms ({K/a.elapsed_time(b)/1e3:.2f} M knots/s); oracle {cpu_knot*1e3:.1f} ms per knot on one core')
