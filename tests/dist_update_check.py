"""Run under torchrun on >= 2 GPUs of one box: the data-parallel update with the peer-memory Adam kernel (reduce='peer') against
the same update over an NCCL all-reduce (reduce='nccl') and against the full-batch single-GPU update; then times both.
Prints 'peer-vs-nccl ok' from rank 0.  Not a pytest file (tests/test_gpu_peer_reduce.py launches it when 2 GPUs are visible)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))

import numpy as np
import torch
import torch.distributed as dist


def main():
    rank, world, local = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ.get('LOCAL_RANK', 0))
    torch.cuda.set_device(local)
    dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    from test_gpu_peer_reduce import batch, flat, make_rl
    from cacto_b200 import environment as genv
    from cacto_b200.NeuralNetwork import NN
    from cacto_b200.RL import RL_AC
    from cacto_b200.conf import get_conf
    system = os.environ.get('CACTO_CHECK_SYSTEM', 'manipulator')
    conf = get_conf(system)
    env = genv.make_env(conf)
    mk = lambda reduce: RL_AC(env, NN(env, conf, 1e-2, seed=0), conf, 0, dist=dist, reduce=reduce)
    peer, nccl = mk('peer'), mk('nccl')
    peer.setup_model()
    nccl.setup_model()
    assert peer._peer is not None and nccl._peer is None
    _, ref = make_rl(system)
    Bl = 32
    for it in range(4):
        g = batch(conf, Bl * world, 30 + it)                  # same seed on every rank: the global minibatch
        sl = slice(rank * Bl, (rank + 1) * Bl)
        a = [g[k][sl] for k in ('state', 'state_next', 'partial_rtg', 'dVdx', 'done', 'term', 'weights')]
        for rl in (peer, nccl):
            rl.update(*a, fuse_target=True)
        ref.update(*[g[k] for k in ('state', 'state_next', 'partial_rtg', 'dVdx', 'done', 'term', 'weights')], fuse_target=True)
    torch.cuda.synchronize()
    wp, wn, wr = flat(peer), flat(nccl), flat(ref)
    scale = float(wr.abs().max())
    assert float((wp - wr).abs().max()) <= 1e-4 * scale, ('peer vs full batch', float((wp - wr).abs().max()))
    assert float((wp - wn).abs().max()) <= 1e-4 * scale, ('peer vs nccl', float((wp - wn).abs().max()))
    gathered = [torch.empty_like(wp) for _ in range(world)]
    dist.all_gather(gathered, wp)
    assert all(torch.equal(x, gathered[0]) for x in gathered), 'peer replicas must be bit-identical'

    # CUDA-graph replays (capture must leave the training state and the exchange protocol intact), then timing of eager
    # launches and replays: device time, max over ranks
    out = {}
    graphs = {}
    for name, rl in (('peer', peer), ('nccl', nccl)):
        graphs[name] = rl.make_update_graph(Bl)
    assert torch.equal(flat(peer), wp), 'capture changed the training state'
    for it in range(3):
        g = batch(conf, Bl * world, 40 + it)
        for name in ('peer', 'nccl'):
            for k in ('state', 'state_next', 'partial_rtg', 'dVdx', 'done', 'term', 'weights'):
                graphs[name].io[k].copy_(g[k][sl])
            graphs[name].replay()
    torch.cuda.synchronize()
    wp, wn = flat(peer), flat(nccl)
    assert float((wp - wn).abs().max()) <= 1e-4 * scale, ('graph: peer vs nccl', float((wp - wn).abs().max()))
    dist.all_gather(gathered, wp)
    assert all(torch.equal(x, gathered[0]) for x in gathered), 'peer replicas must be bit-identical after graph replays'
    for name, rl in (('peer', peer), ('nccl', nccl)):
        graph = graphs[name]
        for mode, fn in (('eager', lambda: rl._update_static(graph.io)), ('graph', graph.replay)):
            for _ in range(20):
                fn()
            dist.barrier()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(200):
                fn()
            e1.record()
            torch.cuda.synchronize()
            t = torch.tensor([e0.elapsed_time(e1) / 200 * 1e3], device='cuda')
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            out['%s_%s_us' % (name, mode)] = round(float(t[0]), 1)
    if rank == 0:
        print('world', world, 'system', system, 'local batch', Bl, out, flush=True)
        print('peer-vs-nccl ok', flush=True)
    del graphs, graph                      # captured NCCL nodes must be gone before the communicator is torn down
    torch.cuda.synchronize()
    dist.barrier()
    sys.stdout.flush()
    os._exit(0)                            # skip interpreter teardown: nothing left to check, and NCCL teardown after graph capture can block


if __name__ == '__main__':
    main()
