import sys, time
import numpy as np, torch
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
from test_gpu_nn import make, rel

def run(system, B, w_S=1e-2, **over):
    conf, env, nn, rl, batch = make(system, B, w_S=w_S, **over)
    s, pr, sn, dv, d, term, w = batch
    target = [t + 0.01 * np.random.default_rng(5).normal(size=t.shape).astype(np.float32) for t in rl.target_critic.get_weights()]
    rl.target_critic.set_weights(target)
    out = {}
    for eng in ('fma', 'tc'):
        nn.update_engine = eng
        g, rtg, V, Vt = nn.compute_critic_grad(rl.critic_model, rl.target_critic, s, sn, pr, dv, d, w)
        torch.cuda.synchronize()
        cg = [x.clone() for x in g]
        loss = float(nn.last_critic_loss)
        ga, act = nn.compute_actor_grad(rl.actor_model, rl.critic_model, s, term, None, return_actions=True)
        torch.cuda.synchronize()
        out[eng] = dict(cg=cg, rtg=rtg.clone(), V=V.clone(), Vt=Vt.clone(), loss=loss, ag=[x.clone() for x in ga], act=act.clone())
    a, b = out['tc'], out['fma']
    r = lambda x, y: rel(x, y.cpu().numpy())
    print(f'--- {system} B={B} w_S={w_S} {over}')
    print('  rtg %.2e V %.2e Vt %.2e loss %.3e/%.3e act %.2e' % (r(a['rtg'], b['rtg']), r(a['V'], b['V']), r(a['Vt'], b['Vt']), a['loss'], b['loss'], r(a['act'], b['act'])))
    print('  critic grads', ' '.join('%.1e' % r(x, y) for x, y in zip(a['cg'], b['cg'])))
    print('  actor grads ', ' '.join('%.1e' % r(x, y) for x, y in zip(a['ag'], b['ag'])), flush=True)
    # timing (device-resident inputs)
    dev='cuda'
    s, sn, pr, dv, d, w = [torch.as_tensor(x, device=dev) for x in (s, sn, pr, dv, d, w)]
    term = torch.as_tensor(term, device=dev)
    for eng in ('fma', 'tc'):
        nn.update_engine = eng
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        for it in range(3):
            if it == 1: ev[0].record()
            for _ in range(5):
                rl.update(s, sn, pr, dv, d, term, w, fuse_target=True)
        ev[1].record(); torch.cuda.synchronize()
        print('  %s update: %.1f us' % (eng, ev[0].elapsed_time(ev[1]) * 1e3 / 10), flush=True)

cases = [('manipulator', 4096, 1e-2, {}), ('manipulator', 3000, 1e-2, {}), ('manipulator', 4096, 0.0, {}), ('manipulator', 4096, 1e-2, dict(MC=1)),
         ('car', 16384, 1e-2, {}), ('ur5', 2304, 1e-2, {}), ('single_integrator', 2048, 1e-2, {}), ('manipulator', 16384, 1e-2, {})]
if len(sys.argv) > 1:
    cases = cases[:int(sys.argv[1])]
for sy, B, w, o in cases:
    run(sy, B, w, **o)
