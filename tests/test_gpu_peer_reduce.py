"""Data-parallel update with the gradient all-reduce fused into the Adam kernel over peer memory (update.cu: k_adam_peer,
cacto_b200/parallel.py: PeerRegion / PeerReduce).

Single-GPU part: W simulated ranks share ONE device -- each rank is an RL_AC with its own stream and its own block of a
``PeerRegion.local_group`` -- so that the arrival-flag protocol, the rank-ordered sum and the deferred clearing of the
gradient blocks run exactly as across GPUs.  Multi-GPU part (skipped with fewer than 2 devices): real processes, CUDA-IPC
mapped blocks, checked against the NCCL path by tests/dist_update_check.py under torchrun."""
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

from cacto_b200.conf import get_conf

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class SimDist:
    """The slice of the torch.distributed surface RL_AC touches; the simulated ranks are seeded identically, so broadcast
    has nothing to do."""

    def __init__(self, rank, world):
        self.rank, self.world = rank, world

    def get_rank(self):
        return self.rank

    def get_world_size(self):
        return self.world

    def is_initialized(self):
        return True

    def broadcast(self, tensor, src=0):
        return tensor


def make_rl(system, dist=None, region=None, **over):
    from cacto_b200 import environment as genv
    from cacto_b200.NeuralNetwork import NN
    from cacto_b200.RL import RL_AC
    conf = get_conf(system, **over)
    env = genv.make_env(conf)
    rl = RL_AC(env, NN(env, conf, 1e-2, seed=0), conf, 0, dist=dist, peer_region=region, peer_max_ctas=16)
    rl.setup_model()
    return conf, rl


def batch(conf, B, seed):
    rng = np.random.default_rng(seed)
    ns = conf.nb_state
    lo, hi = np.asarray(conf.x_init_min, float), np.asarray(conf.x_init_max, float)
    dv = rng.normal(size=(B, ns))
    dv[:, -1] = 0
    f = lambda a, dt=torch.float32: torch.tensor(np.asarray(a), dtype=dt, device='cuda')
    return dict(state=f(rng.uniform(lo, hi, (B, ns))), state_next=f(rng.uniform(lo, hi, (B, ns))), partial_rtg=f(rng.uniform(-5, 0, (B, 1))),
                dVdx=f(dv), done=f(rng.uniform(size=(B, 1)) < 0.5), term=f(rng.uniform(size=(B, 1)) < 0.1, torch.float64),
                weights=f(rng.uniform(0.5, 1.5, (B, 1))))


def flat(rl):
    return torch.cat([n.params for n in (rl.critic_model, rl.target_critic, rl.actor_model)]).clone()


IO_KEYS = ('state', 'state_next', 'partial_rtg', 'dVdx', 'done', 'term', 'weights')


SIM_CASES = [('manipulator', 2, False), ('double_integrator', 4, False), ('ur5', 3, True), ('manipulator', 2, True)]


def simulated_ranks_case(system, world, static):
    """static = the allocation-free launch sequence that RL.UpdateGraph captures (gradient blocks cleared only by the peer
    kernel of the other network); otherwise the eager RL_AC.update."""
    from cacto_b200.parallel import PeerReduce, PeerRegion
    conf, ref = make_rl(system)
    regions = PeerRegion.local_group(PeerReduce.region_bytes(ref.critic_model.n, ref.actor_model.n), world)
    ranks = [make_rl(system, SimDist(r, world), regions[r])[1] for r in range(world)]
    assert all(rl._peer is not None for rl in ranks)
    streams = [torch.cuda.Stream() for _ in range(world)]
    Bl = 16
    ios = []
    for r in range(world):                                   # per-rank buffers; also warms the allocator pool of every stream so that
        with torch.cuda.stream(streams[r]):                  # no cudaMalloc happens while another rank's kernel waits for this one
            io = batch(conf, Bl, 0)
            io.update(rtg=torch.zeros((Bl, 1), device='cuda'), V=torch.zeros((Bl, 1), device='cuda'), V_target=torch.zeros((Bl, 1), device='cuda'))
            ios.append(io)
            for o, n in ((ranks[r].critic_optimizer, ranks[r].critic_model), (ranks[r].actor_optimizer, ranks[r].actor_model)):
                o.moments(n)
                o._device_state(n.params.device)
            scratch = [torch.empty((Bl, 1), device='cuda') for _ in range(8)] + [torch.empty((Bl, conf.nb_state), device='cuda') for _ in range(8)]
            del scratch
    torch.cuda.synchronize()
    for it in range(3):
        g = batch(conf, Bl * world, 10 + it)
        ref.update(*[g[k] for k in IO_KEYS], fuse_target=True)
        for r in range(world):
            for k in IO_KEYS:
                ios[r][k].copy_(g[k][r * Bl:(r + 1) * Bl])
        torch.cuda.synchronize()
        for r, rl in enumerate(ranks):                       # no host synchronisation between the ranks' launches
            with torch.cuda.stream(streams[r]):
                if static:
                    rl._update_static(ios[r])
                else:
                    rl.update(*[ios[r][k] for k in IO_KEYS], fuse_target=True)
        torch.cuda.synchronize()
    w0 = flat(ranks[0])
    for rl in ranks[1:]:
        assert torch.equal(flat(rl), w0), 'replicas must stay bit-identical (rank-ordered sum)'
    wr = flat(ref)
    assert float((w0 - wr).abs().max()) <= 1e-4 * float(wr.abs().max()), float((w0 - wr).abs().max())
    for rg in regions:
        rg.close()


def test_simulated_ranks_match_full_batch_update():
    """Runs SIM_CASES through ``simulated_ranks_case`` in a child process (tests/peer_sim_check.py): a rank that never arrives makes
    the kernel trap after 20 s, which poisons the CUDA context of the process -- the child keeps that away from the rest of the suite."""
    r = subprocess.run([sys.executable, os.path.join(ROOT, 'tests', 'peer_sim_check.py')], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert r.stdout.count('case ok') == len(SIM_CASES) and 'peer-sim ok' in r.stdout, r.stdout[-2000:]


@pytest.mark.skipif(torch.cuda.device_count() < 2 or os.environ.get('CACTO_B200_MULTI_GPU_TESTS') != '1',
                    reason='needs 2 GPUs of one box and CACTO_B200_MULTI_GPU_TESTS=1 (spawns torchrun; run by hand: see profiles/r1_dist_update_check_n2.txt)')
def test_two_processes_peer_reduce_equals_nccl():
    port = 29500 + os.getpid() % 2000
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node=2', '--master-addr', '127.0.0.1', '--master-port',
           str(port), os.path.join(ROOT, 'tests', 'dist_update_check.py')]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert 'peer-vs-nccl ok' in r.stdout
