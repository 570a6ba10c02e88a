"""world_size-2 gloo test (CPU) of the data-parallel host logic: block partition of units, and the rule
'gradient sums scaled by 1/B_global, all-reduced with SUM, equal the full-batch gradient' that RL_AC.update
relies on.  Gradients come from the oracle (the CUDA kernels need a GPU); the partition/collective code under
test is cacto_b200.parallel, the same functions the GPU path calls over NCCL."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from cacto_b200.parallel import allreduce_sum, global_batch, shard_range


def test_shard_range_partitions_exactly():
    for n in (0, 1, 7, 64, 131072, 1000003):
        for world in (1, 2, 3, 8):
            spans = [shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from cacto_b200.conf import get_conf
    from oracle import nn as onn
    conf = get_conf('double_integrator')
    ns, B = conf.nb_state, 24
    rng = np.random.default_rng(0)                      # identical global minibatch on every rank
    s = rng.uniform(conf.x_init_min, conf.x_init_max, (B, ns)).astype(np.float32)
    sn = rng.uniform(conf.x_init_min, conf.x_init_max, (B, ns)).astype(np.float32)
    pr = rng.uniform(-5, 0, (B, 1)).astype(np.float32)
    dv = rng.normal(size=(B, ns)).astype(np.float32)
    d = (rng.uniform(size=(B, 1)) < 0.5).astype(np.float32)
    w = rng.uniform(0.5, 1.5, (B, 1)).astype(np.float32)
    critic = onn.init_critic_sine(ns, seed=1)
    lo, hi = shard_range(B, rank, world)
    assert global_batch(B // world, dist) == B
    g_local = onn.critic_grad(critic, critic, conf, 1e-2, s[lo:hi], sn[lo:hi], pr[lo:hi], dv[lo:hi], d[lo:hi], w[lo:hi])[0]
    flat = torch.cat([torch.tensor(g).reshape(-1) for g in g_local]) * ((hi - lo) / B)     # sum over shard / B_global
    allreduce_sum(flat, dist)
    g_full = onn.critic_grad(critic, critic, conf, 1e-2, s, sn, pr, dv, d, w)[0]
    ref = torch.cat([torch.tensor(g).reshape(-1) for g in g_full])
    err = float((flat - ref).abs().max() / ref.abs().max())
    np.save(os.path.join(out_dir, f'err{rank}.npy'), np.array([err]))
    np.save(os.path.join(out_dir, f'grad{rank}.npy'), flat.numpy())
    dist.destroy_process_group()


def test_sharded_gradient_allreduce_equals_full_batch(tmp_path):
    world, port = 2, _free_port()
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    errs = [float(np.load(tmp_path / f'err{r}.npy')[0]) for r in range(world)]
    assert max(errs) < 1e-5, errs
    g0, g1 = np.load(tmp_path / 'grad0.npy'), np.load(tmp_path / 'grad1.npy')
    np.testing.assert_array_equal(g0, g1)               # every replica holds the same reduced gradient
