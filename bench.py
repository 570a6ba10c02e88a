#!/usr/bin/env python
"""bench.py -- CACTO hot-path throughput on B200 (BASELINE.json metric: manipulator rollout env-steps/s
+ Sobolev actor-critic updates/s).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--engine tc|fma]

One "step" = one pass of the fused rollout kernel (K1) over this rank's batch of synthetic initial
conditions: BASELINE config[3] (3-DOF planar manipulator, 1 M rollouts x 100 steps sharded over 8 GPUs)
= 131072 rollouts x 100 env-steps per GPU (weak scaling: per-GPU work fixed).  `value` counts env-steps
with the inputs resident in HBM; `e2e` runs the same pass through the public Python API
(RL_AC.rollout_to_host) with the initial conditions in pinned HOST memory and the fp64 warm-start
trajectories delivered to pinned host memory inside the timed region.  The secondary number (Sobolev
critic+actor updates/s, conf batch 64 per GPU) is in `extra`.  `cpu_baseline` times the oracle's restatement
of the reference's rollout loop (RL.py:221-231 over multiprocessing.Pool like main.py:220-225) on the host
cores, on a bounded sample.  `--impl reference` prints that CPU arm alone as the reference line.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SYSTEM = 'manipulator'
ROLLOUTS_PER_GPU = 131072
P_ACTOR_MACS = 256 * 7 + 65536 + 256 * 3          # P_a for the manipulator
F_DYN = 250                                        # flops of the planar-3R forward dynamics step
FLOPS_PER_ENV_STEP = 2 * P_ACTOR_MACS + F_DYN
# dram__bytes_read.sum + dram__bytes_write.sum of the rollout kernel per launch at the default workload, from the committed
# `ncu --set full` captures under profiles/ (None until an engine has been captured)
TRAFFIC_BYTES = {'tc': 8695296 + 1001016000, 'tf32': 8953600 + 998347520}
TRAFFIC_SOURCE = {'tc': 'profiles/r2_rollout_tc16_ncu_raw.csv (ncu --set full capture of the same launch at the end of round 2, profiles/scripts/prof_rollout.py; '
                        'not re-measured in-run; algorithmic output 1.056 GB allocated, 1.00 GB written: rows past a horizon are not stored)',
                  'tf32': 'profiles/r1_rollout_tc_ncu_raw.csv (ncu --set full capture; not re-measured in-run)'}


# ----------------------------------------------------------------------------------------- CPU arm
def _cpu_worker(args):
    """One worker of the reference-structured CPU rollout: B=1 actor forward + one simulate per step."""
    import torch
    torch.set_num_threads(1)
    from cacto_b200.conf import get_conf
    from oracle import nn as onn, rtg as ortg, systems as osys
    seed, n_rollouts, weights = args
    conf = get_conf(SYSTEM)
    env = osys.make_env(conf)
    ap = onn.to_torch(weights)
    rng = np.random.default_rng(seed)

    def actor_eval(x):
        with torch.no_grad():
            return onn.actor_forward(ap, torch.tensor(x, dtype=torch.float32), conf).numpy()[0]
    steps = 0
    for _ in range(n_rollouts):
        x0 = rng.uniform(conf.x_init_min, conf.x_init_max)
        x0[-1] = 0.0
        _, st, ct, T, ok = ortg.create_to_init(conf, env, actor_eval, 1, x0)
        steps += T
    return steps


def cpu_rollout_rate(cores, rollouts_per_core):
    """env-steps/s of the CPU arm on `cores` processes (fork; must run before CUDA is initialised)."""
    import multiprocessing as mp
    from oracle import nn as onn
    weights = onn.init_actor(7, 3, seed=0)
    ctx = mp.get_context('fork')
    with ctx.Pool(cores) as pool:
        pool.map(_cpu_worker, [(100 + i, 1, weights) for i in range(cores)])          # warm-up: imports, page-in
        t0 = time.perf_counter()
        steps = sum(pool.map(_cpu_worker, [(i, rollouts_per_core, weights) for i in range(cores)]))
        dt = time.perf_counter() - t0
    return steps / dt, steps, dt


def _update_batch(conf, B, seed):
    """Synthetic replay rows of SURVEY.md 8(d): s, s_next ~ U(x_init), partial_rtg ~ U(-5, 0), dVdx ~ N(0, 1) (time column 0),
    done ~ Bernoulli(0.5), term ~ Bernoulli(0.01), weights 1."""
    rng = np.random.default_rng(seed)
    ns = conf.nb_state
    lo, hi = np.asarray(conf.x_init_min, float), np.asarray(conf.x_init_max, float)
    s = rng.uniform(lo, hi, (B, ns)).astype(np.float32)
    sn = rng.uniform(lo, hi, (B, ns)).astype(np.float32)
    pr = rng.uniform(-5, 0, (B, 1)).astype(np.float32)
    dv = rng.normal(size=(B, ns)).astype(np.float32)
    dv[:, -1] = 0
    d = (rng.uniform(size=(B, 1)) < 0.5).astype(np.float32)
    term = (rng.uniform(size=(B, 1)) < 0.01).astype(np.float64)
    w = np.ones((B, 1), np.float32)
    return s, pr, sn, dv, d, term, w


def cpu_update_baseline(system, B, cores, n=3, dyn_sample=None):
    """C4 of BASELINE.md: the oracle's restatement of RL_AC.update (torch-CPU fp32 autograd with create_graph for the Sobolev
    term, TF-style Adam, Polyak; per-sample Python loops for simulate_batch / derivative_batch as environment.py:134-144) on
    `cores` torch threads.  With ``dyn_sample`` the per-sample dynamics loops run on that many rows and are scaled to B (stated)."""
    import torch
    from cacto_b200.conf import get_conf
    from oracle import nn as onn, systems as osys
    torch.set_num_threads(cores)
    conf = get_conf(system)
    env = osys.make_env(conf)
    critic, actor = onn.init_critic_sine(conf.nb_state, seed=0), onn.init_actor(conf.nb_state, conf.nb_action, seed=1)
    target = [c.copy() for c in critic]
    oc, oa = onn.Adam(critic, conf.CRITIC_LEARNING_RATE), onn.Adam(actor, conf.ACTOR_LEARNING_RATE)
    batch = _update_batch(conf, B, 0)
    if dyn_sample is None or dyn_sample >= B:
        onn.update(critic, target, actor, oc, oa, conf, 1e-2, env, batch)
        t0 = time.perf_counter()
        for _ in range(n):
            onn.update(critic, target, actor, oc, oa, conf, 1e-2, env, batch)
        dt = (time.perf_counter() - t0) / n
        return 1.0 / dt, f'{n} full updates of batch {B} ({dt:.3f} s each)'
    # bounded: network part on the full batch with the per-sample dynamics replaced by a scaled sample
    s, pr, sn, dv, d, term, w = batch
    t0 = time.perf_counter()
    env.simulate_batch(s[:dyn_sample], np.zeros((dyn_sample, conf.nb_action), np.float32))
    env.derivative_batch(s[:dyn_sample], np.zeros((dyn_sample, conf.nb_action), np.float32))
    t_dyn = (time.perf_counter() - t0) * B / dyn_sample

    class _Env:                       # stand-in dynamics terms of the right shape: only the (scaled) cost of the real loops is charged
        def simulate_batch(self, st, a):
            return np.asarray(st, np.float32)

        def derivative_batch(self, st, a):
            return np.full((len(st), conf.nb_state, conf.nb_action), 1e-3, np.float32)
    fake = _Env()
    onn.update(critic, target, actor, oc, oa, conf, 1e-2, fake, batch)
    t0 = time.perf_counter()
    for _ in range(n):
        onn.update(critic, target, actor, oc, oa, conf, 1e-2, fake, batch)
    t_net = (time.perf_counter() - t0) / n
    dt = t_net + t_dyn
    return 1.0 / dt, (f'batch {B}: torch-CPU network part measured on the full batch ({t_net:.3f} s, {cores} threads) + per-sample '
                      f'simulate/derivative loops measured on {dyn_sample} rows and scaled to {B} ({t_dyn:.1f} s, 1 core)')


def cpu_misc_baselines():
    """C2, C3, C5, C6 of BASELINE.md on one core (the reference's per-sample / per-knot Python loops), small bounded samples."""
    from types import SimpleNamespace
    from cacto_b200.conf import get_conf
    from oracle import per as oper, rtg as ortg, systems as osys
    out = {}
    conf = get_conf(SYSTEM)
    env = osys.make_env(conf)
    rng = np.random.default_rng(0)
    n = 128
    s = rng.uniform(conf.x_init_min, conf.x_init_max, (n, conf.nb_state)).astype(np.float32)
    a = rng.uniform(conf.u_min, conf.u_max, (n, conf.nb_action)).astype(np.float32)
    w = np.tile(np.asarray(conf.cost_weights_running, float), (n, 1))
    for name, fn in (('simulate_batch', lambda: env.simulate_batch(s, a)), ('derivative_batch', lambda: env.derivative_batch(s, a)),
                     ('reward_batch', lambda: env.reward_batch(w, s, a))):
        t0 = time.perf_counter(); fn(); out['C2_' + name + '_samples_per_s'] = n / (time.perf_counter() - t0)
    t0 = time.perf_counter()
    for i in range(32):
        env.augmented_derivative(s[i].astype(np.float64), a[i].astype(np.float64))
    out['C3_augmented_derivative_samples_per_s'] = 32 / (time.perf_counter() - t0)
    ns = conf.nb_state
    for B in (64, 4096):
        bc = SimpleNamespace(REPLAY_SIZE=2 ** 16, BATCH_SIZE=B, nb_state=ns, prioritized_replay_alpha=0.6, prioritized_replay_beta=0.6,
                             prioritized_replay_eps=1e-2, fresh_factor=0.95)
        ob = oper.PrioritizedReplayBuffer(bc)
        rows = rng.normal(size=(8192, 3 * ns + 3))
        cols = (rows[:, :ns], rows[:, ns], rows[:, ns + 1:2 * ns + 1], rows[:, 2 * ns + 1:3 * ns + 1], rows[:, 3 * ns + 1], rows[:, 3 * ns + 2])
        ob.add(*[(c,) for c in cols])
        rtg_, V_ = rng.normal(size=(B, 1)).astype(np.float32), rng.normal(size=(B, 1)).astype(np.float32)
        reps = 20 if B == 64 else 2
        t0 = time.perf_counter()
        for _ in range(reps):
            o = ob.sample()
            ob.update_priorities(o[7], rtg_, V_)
        out[f'C5_per_rounds_per_s_B{B}'] = reps / (time.perf_counter() - t0)
    st, c = rng.normal(size=(101, ns)), rng.uniform(0, 2, 101)
    t0 = time.perf_counter()
    for _ in range(20):
        ortg.rl_solve(conf, st, c)
    out['C6_rl_solve_trajectories_per_s'] = 20 / (time.perf_counter() - t0)
    out['cores'] = 1
    out['note'] = ('oracle restatements with the reference\'s per-sample / per-knot Python structure (environment.py:134-144, TO.py:181, '
                   'replay_buffer.py + segment_tree.py, RL.py:173-187), manipulator, capacity-2^16 trees')
    return out


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    vals = []
    per_core = 16                  # ~3 s per bench step on the GPU box's host cores
    for _ in range(args.warmup):
        cpu_rollout_rate(cores, 1)
    for _ in range(args.steps):
        vals.append(cpu_rollout_rate(cores, per_core))
    total_steps = sum(v[1] for v in vals)
    total_dt = sum(v[2] for v in vals)
    rate = total_steps / total_dt
    sample = (f'{cores * per_core} rollouts x 100 steps per bench step ({total_steps} env-steps in {total_dt:.1f} s), t0 = 0; oracle port '
              f'(B=1 torch-CPU actor forward + NumPy fp64 RNEA dynamics per step) over multiprocessing.Pool({cores})')
    emit({
        'impl': 'reference', 'metric': 'manipulator rollout env-steps/s', 'value': rate, 'unit': 'env-steps/s', 'n_gpus': args.gpus,
        'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': 1e3 * total_dt / max(1, args.steps), 'higher_is_better': True,
        'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32 actor / f64 dynamics', 'data': 'synthetic',
        'config': {'workload': 'BASELINE config[3]: manipulator policy rollouts (RL.py:221-231 structure), CPU', 'sample': sample},
        'cpu_baseline': {'value': rate, 'unit': 'env-steps/s', 'cores': cores, 'kind': 'port', 'sample': sample},
        'e2e': {'value': rate, 'unit': 'env-steps/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0})


# ----------------------------------------------------------------------------------------- clocks
class ClockSampler:
    Q = 'index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,' \
        'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap'

    def __init__(self, gpu_index):
        self.rows, self.proc, self.gpu = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', f'--query-gpu={self.Q}', '--format=csv,noheader,nounits', '-lms', '100', '-i',
                                          str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(',')])

    def stop(self):
        if self.proc is not None:
            self.proc.terminate()
        sm, mx, pw, reasons = [], [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2])); pw.append(float(r[3]))
                for name, v in zip(('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'), r[5:9]):
                    if v.lower().startswith('active'):
                        reasons.add(name)
            except Exception:
                pass
        if not sm:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': [], 'samples': 0}
        return {'sm_mhz': float(np.median(sm)), 'sm_max_mhz': float(max(mx)), 'reasons': sorted(reasons), 'samples': len(sm),
                'power_w_max': float(max(pw))}


# ----------------------------------------------------------------------------------------- GPU arm
def run_b200(args):
    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    K, W = args.steps, args.warmup

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:          # before CUDA init: the pool forks
        cores = os.cpu_count() or 1
        per_core = 48              # ~10 s of CPU work on the GPU box's host cores (the contract asks for a 10-30 s bounded sample)
        rate, steps, dt = cpu_rollout_rate(cores, per_core)
        cpu = {'value': rate, 'unit': 'env-steps/s', 'cores': cores, 'kind': 'port',
               'sample': f'{cores * per_core} rollouts x 100 steps ({steps} env-steps, {dt:.1f} s): oracle restatement of RL.py:221-231 '
                         f'(B=1 torch-CPU actor forward + NumPy fp64 RNEA dynamics per step; slower than Pinocchio C++ would be) over '
                         f'multiprocessing.Pool({cores})'}
        if not args.quick:         # BASELINE.md C4 (update) and C2 / C3 / C5 / C6 one-liners, ~15 s together
            try:
                r64, s64 = cpu_update_baseline(SYSTEM, 64, cores, n=3)
                r16, s16 = cpu_update_baseline(SYSTEM, 16384, cores, n=1, dyn_sample=128)
                cpu_update = {'64': {'value': r64, 'unit': 'updates/s', 'cores': cores, 'kind': 'port', 'sample': s64},
                              '16384': {'value': r16, 'unit': 'updates/s', 'cores': cores, 'kind': 'port', 'sample': s16}}
            except Exception as exc:
                cpu_update = {'error': f'{type(exc).__name__}: {exc}'}
            try:
                cpu_misc = cpu_misc_baselines()
            except Exception as exc:
                cpu_misc = {'error': f'{type(exc).__name__}: {exc}'}
            cpu['update'] = cpu_update
            cpu['misc'] = cpu_misc

    import torch
    import torch.distributed as dist
    assert torch.cuda.is_available(), 'bench.py needs a CUDA device (no CPU fallback on the product path)'
    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    from cacto_b200.parallel import bind_to_gpu_numa_node
    placement = bind_to_gpu_numa_node(local_rank)          # before any pinned allocation: first touch on the GPU's node
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    from cacto_b200 import _lib, environment as genv
    from cacto_b200.NeuralNetwork import NN
    from cacto_b200.RL import RL_AC
    from cacto_b200.conf import get_conf

    conf = get_conf(SYSTEM)
    env = genv.make_env(conf)
    rl = RL_AC(env, NN(env, conf, 1e-2, seed=0), conf, 0, dist=dist if world > 1 else None)
    rl.setup_model()
    rl.rollout_engine = args.engine
    B, T, ns, na = args.rollouts_per_gpu, conf.NSTEPS, conf.nb_state, conf.nb_action
    rng = np.random.default_rng(1000 + rank)
    X0 = rng.uniform(conf.x_init_min, conf.x_init_max, (B, ns))
    X0[:, -1] = 0.0                                            # full horizon: fixed work B x NSTEPS env-steps per step
    ics_host = torch.as_tensor(X0).pin_memory()
    ics = ics_host.to(dev)
    hz = torch.full((B,), T, dtype=torch.int32, device=dev)
    states = torch.empty((T + 1, ns, B), dtype=torch.float64, device=dev)
    controls = torch.empty((T, na, B), dtype=torch.float64, device=dev)
    flags = torch.empty(B, dtype=torch.int32, device=dev)
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)     # 256 MiB > 126 MB L2
    stream = torch.cuda.current_stream()
    lib, ptr = _lib.lib, _lib.ptr

    def rollout_step():      # the launches of RL_AC.rollout_batch on pre-allocated outputs (tc: W2 image refresh + rollout kernel)
        rl._launch_rollout(1, ics, hz, T, states, controls, flags, None, B)
    launches_per_step = {'tc': 3, 'tf32': 2, 'fma': 1}[args.engine]      # W2 image kernels + the rollout kernel

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def timed(fn, steps, do_flush=True):
        """Per-step CUDA-event durations (ms) on the launching stream; L2 flushed between steps."""
        evs = []
        for _ in range(steps):
            if do_flush:
                flush.fill_(1.0)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(stream); fn(); b.record(stream)
            evs.append((a, b))
        torch.cuda.synchronize()
        return [a.elapsed_time(b) for a, b in evs]

    def max_over_ranks(x):
        if world == 1:
            return float(x)
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t[0])

    # FP32 FMA peak, measured in-run (roofline denominator of the fp32-FMA kernels)
    pk_out = torch.zeros(4, dtype=torch.float32, device=dev)
    pk_iters, pk_blocks = 4096, 148 * 16
    for _ in range(2):
        lib.cacto_peak_fma_fp32(ptr(pk_out), pk_iters, pk_blocks, _lib.stream_ptr())
    pk_ms = min(timed(lambda: lib.cacto_peak_fma_fp32(ptr(pk_out), pk_iters, pk_blocks, _lib.stream_ptr()), 5, do_flush=False))
    fma_peak_tflops = pk_blocks * 256 * pk_iters * 16 * 2 / (pk_ms * 1e-3) / 1e12

    # ---- K1 rollouts: device-resident
    for _ in range(W):
        rollout_step()
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ms = timed(rollout_step, K)
    barrier()
    total_ms = max_over_ranks(float(sum(ms)))
    assert bool(flags.all()), 'a rollout hit NaN'
    value = B * T * K * world / (total_ms * 1e-3)
    kernel_ms = float(np.mean(ms))
    achieved_tflops = B * T * FLOPS_PER_ENV_STEP / (kernel_ms * 1e-3) / 1e12

    # ---- e2e: host ICS (pinned) -> H2D -> rollout -> fp64 warm-start trajectories in pinned host memory
    def e2e_leg(compact):
        # two sets of host buffers: batch k + 1 is queued before batch k is waited for (its first kernel and host-side bookkeeping
        # overlap the tail of batch k's copies); every batch is waited for and checked inside the timed region
        def host_set():
            if compact:
                return (torch.empty((T + 1, ns - 1, B), dtype=torch.float64).pin_memory(),
                        torch.empty((T, na, B), dtype=torch.float32).pin_memory(), torch.empty(B, dtype=torch.int32).pin_memory())
            return (torch.empty((T + 1, ns, B), dtype=torch.float64).pin_memory(),
                    torch.empty((T, na, B), dtype=torch.float64).pin_memory(), torch.empty(B, dtype=torch.int32).pin_memory())
        host_sets = (host_set(), host_set())
        for _ in range(max(1, W // 2)):
            for hs in host_sets:
                rl.rollout_to_host(ics_host, 1, hs[0], hs[1], hs[2], compact=compact)
        barrier()
        t0 = time.perf_counter()
        pending, ok_all = None, True
        for k in range(K):
            hs = host_sets[k & 1]
            nxt = rl.rollout_to_host(ics_host, 1, hs[0], hs[1], hs[2], wait=False, compact=compact)
            if pending is not None:
                pending[0].wait()
                ok_all = ok_all and bool(pending[1].all())
            pending = (nxt, hs[2])
        pending[0].wait()
        ok_all = ok_all and bool(pending[1].all())
        torch.cuda.synchronize()
        e2e_s = max_over_ranks(time.perf_counter() - t0)
        assert ok_all
        hs = host_sets[0]
        d2h_bytes = hs[0].numel() * 8 + hs[1].numel() * hs[1].element_size() + hs[2].numel() * 4
        del host_sets
        return B * T * K * world / e2e_s, d2h_bytes

    e2e_value, d2h = e2e_leg(False)
    # the same batches in the compact transfer format (no time row, controls as the fp32 values the actor produced): bit-
    # reconstructible on the host (RL.CompactRollouts, tests/test_gpu_rollout.py), 25 % fewer bytes over PCIe
    e2e_compact_value, d2h_compact = e2e_leg(True)
    h2d = ics_host.numel() * 8 + B * 4

    # ---- K3 Sobolev critic+actor updates/s (the second half of the metric): data resident, gradients summed across GPUs inside
    # the Adam kernels.  Legs: the reference's conf batch (64 per GPU, fp32-FMA tile kernels, latency-bound) and the BASELINE
    # config 2 / 3 batches (4096, 16384: tcgen05 engine from B = 2400); UR5 (config 5) with its global batches split over the ranks.
    from cacto_b200.replay_buffer import ReplayBuffer

    def update_leg(system, B_local, n_eager, n_graph, e2e_updates=0):
        cf = get_conf(system, BATCH_SIZE=B_local)
        ev_ = genv.make_env(cf)
        nn_ = NN(ev_, cf, 1e-2, seed=0)
        r_ = RL_AC(ev_, nn_, cf, 0, dist=dist if world > 1 else None)
        r_.setup_model()
        bt = [torch.as_tensor(x).to(dev) for x in _update_batch(cf, B_local, 7 + rank)]
        s_, pr_, sn_, dv_, d_, term_, w_ = bt
        r_.peer_barrier()
        for _ in range(5):
            r_.update(s_, sn_, pr_, dv_, d_, term_, w_, fuse_target=True, synced=True)
        barrier()
        a_, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a_.record(stream)
        for _ in range(n_eager):
            r_.update(s_, sn_, pr_, dv_, d_, term_, w_, fuse_target=True, synced=True)
        b_.record(stream)
        torch.cuda.synchronize()
        eager_us = max_over_ranks(a_.elapsed_time(b_)) * 1e3 / n_eager
        graph_us = None
        if world == 1 or r_._peer is not None:
            try:
                ug = r_.make_update_graph(B_local)
                for k_, t_ in (('state', s_), ('state_next', sn_), ('partial_rtg', pr_), ('dVdx', dv_), ('done', d_), ('term', term_), ('weights', w_)):
                    ug.io[k_].copy_(t_)
                for _ in range(10):
                    ug.replay()
                barrier()
                a_, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a_.record(stream)
                for _ in range(n_graph):
                    ug.replay()
                b_.record(stream)
                torch.cuda.synchronize()
                graph_us = max_over_ranks(a_.elapsed_time(b_)) * 1e3 / n_graph
                del ug
            except Exception as exc:           # report, do not hide
                graph_us = f'failed: {type(exc).__name__}: {exc}'
        # one GPU, fused engine: consecutive updates software-pipelined (RL.PipelinedUpdateGraph, learn_and_update's default there):
        # the actor step of update i beside the critic gradient of update i + 1, same weights as the sequential order
        pipe_us = None
        can_pipeline = world == 1 and r_.critic_model.kind == 'critic_sine' and not nn_._use_tc(B_local)
        if can_pipeline:
            try:
                pg = r_.make_pipelined_update_graph(B_local)
                for io_ in pg.ios:
                    for k_, t_ in (('state', s_), ('state_next', sn_), ('partial_rtg', pr_), ('dVdx', dv_), ('done', d_), ('term', term_), ('weights', w_)):
                        io_[k_].copy_(t_)
                for _ in range(10):
                    pg.replay()
                a_, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a_.record(stream)
                for _ in range(n_graph):
                    pg.replay()
                pg.flush()
                b_.record(stream)
                torch.cuda.synchronize()
                pipe_us = a_.elapsed_time(b_) * 1e3 / n_graph
                del pg
            except Exception as exc:           # report, do not hide
                pipe_us = f'failed: {type(exc).__name__}: {exc}'
        best_us = graph_us if isinstance(graph_us, float) else eager_us
        if isinstance(pipe_us, float):
            best_us = min(best_us, pipe_us)
        ns_, na_ = cf.nb_state, cf.nb_action
        P_c, P_a = 64 * ns_ + 28800, 256 * ns_ + 65536 + 256 * na_
        flops = 2.0 * B_local * (10 * P_c + 3 * P_a)                 # MACs ~ B (10 P_c + 3 P_a)
        ach = flops / (best_us * 1e-6) / 1e12
        tc = nn_._use_tc(B_local)
        leg = {'system': system, 'batch_per_gpu': B_local, 'global_batch': B_local * world, 'engine': 'tcgen05 (update_tc.cu)' if tc else 'fp32 FMA (update.cu)',
               'us_per_update_eager': eager_us, 'us_per_update_cuda_graph': graph_us, 'us_per_update_pipelined_graph': pipe_us, 'updates_per_s': 1e6 / best_us,
               'samples_per_s': B_local * world * 1e6 / best_us, 'algorithmic_flops_per_update': flops}
        if tc:
            leg['roofline'] = {'bound': 'tensor', 'achieved': ach, 'peak': bf16_peak, 'unit': 'TFLOP/s', 'frac': ach / bf16_peak,
                               'ceiling_3xfp16': bf16_peak / 3.0, 'frac_of_3xfp16_ceiling': ach / (bf16_peak / 3.0), 'peak_source': peak_src,
                               'note': 'layer-wise sweeps over 128-sample tiles: bound by the per-sample intermediates that travel through the '
                                       'workspace (L2 / HBM) and by CUDA-core epilogues, not by the tensor pipe (profiles/README.md)'}
            # the implementation's own traffic: floats written + read per sample by the sweeps and the weight-gradient GEMMs
            # (critic 2468 w + 3642 r, actor 1480 w + 2262 r: DESIGN.md section 4), against the measured HBM copy bandwidth
            ws_bytes = 4.0 * (2468 + 3642 + 1480 + 2262) * B_local
            hbm = peaks.get('hbm_gbs') or 6650.0
            leg['workspace_traffic'] = {'bytes_per_update': ws_bytes, 'achieved_gbs': ws_bytes / (best_us * 1e-6) / 1e9, 'hbm_peak_gbs': hbm,
                                        'frac_of_hbm': ws_bytes / (best_us * 1e-6) / 1e9 / hbm,
                                        'note': 'analytic workspace bytes (most of them L2 hits at these sizes: 126 MB L2), not a DRAM counter'}
        else:
            leg['roofline'] = {'bound': 'fp32_fma' if B_local >= 1024 else 'latency', 'achieved': ach, 'peak': fma_peak_tflops, 'unit': 'TFLOP/s',
                               'frac': ach / fma_peak_tflops, 'peak_source': 'cacto_peak_fma_fp32 measured in this run'}
        if e2e_updates:
            # end to end through the public API, as RL_AC.learn_and_update runs it (RL.py:120-143): ReplayBuffer.sample() draws the
            # indices on the host (np.random, quirk Q5), ships them, gathers the rows; update; the loss is read back every update
            buf = ReplayBuffer(cf)
            rows = torch.randn((cf.REPLAY_SIZE + 8, 3 * ns_ + 3), dtype=torch.float64, device=dev)      # wraps: the buffer reports full
            rows[:, :ns_] = s_.double()[torch.randint(0, B_local, (cf.REPLAY_SIZE + 8,), device=dev)]
            rows[:, ns_ + 1:2 * ns_ + 1] = rows[:, :ns_]
            rows[:, 3 * ns_ + 1] = (rows[:, 3 * ns_ + 1] > 0).double()
            rows[:, 3 * ns_ + 2] = (rows[:, 3 * ns_ + 2] > 2.3).double()
            buf.add_rows(rows)
            np.random.seed(0)
            ug_ = None
            if world == 1 or r_._peer is not None:
                try:                                             # learn_and_update's defaults: the update replayed as a CUDA graph,
                    ug_ = r_.make_pipelined_update_graph(B_local) if can_pipeline else r_.make_update_graph(B_local)   # pipelined on one GPU
                except Exception:
                    ug_ = None
            r_.peer_barrier()
            # every update's loss is read back; the read of update k is issued behind it (asynchronous copy into pinned memory)
            # and waited for after update k + 1 has been launched, so that the host's sampling work overlaps the device's update
            loss_pin = torch.zeros(2, dtype=torch.float32).pin_memory()
            loss_ev = [torch.cuda.Event(), torch.cuda.Event()]
            loss_host, n_read = float('nan'), 0
            rows_ = buf.index_stream(3) if ug_ is not None else None
            for it in range(3 + e2e_updates):
                if it == 3:
                    torch.cuda.synchronize()
                    t0_ = time.perf_counter()
                    if ug_ is not None:                          # inside the timed region, as learn_and_update does it: the index draws
                        rows_ = buf.index_stream(e2e_updates)                      # of the loop in chunks (ReplayBuffer.index_stream)
                if ug_ is not None:
                    buf.sample(next(rows_), out=ug_.io)
                    ug_.replay()
                else:
                    bs = buf.sample()
                    r_.update(bs[0], bs[2], bs[1], bs[3], bs[4], bs[5], bs[6], fuse_target=True, synced=True)
                loss_pin[it & 1:(it & 1) + 1].copy_(nn_.last_critic_loss.reshape(1), non_blocking=True)
                loss_ev[it & 1].record()
                if it > 0:
                    loss_ev[(it - 1) & 1].synchronize()
                    loss_host = float(loss_pin[(it - 1) & 1])    # device -> host read of update it - 1
                    n_read += 1
            if hasattr(ug_, 'flush'):
                ug_.flush()                                      # the outstanding actor step of the pipelined graph
                torch.cuda.synchronize()
            loss_ev[(3 + e2e_updates - 1) & 1].synchronize()
            loss_host = float(loss_pin[(3 + e2e_updates - 1) & 1])
            n_read += 1
            assert n_read == 3 + e2e_updates
            e2e_s_ = max_over_ranks(time.perf_counter() - t0_)
            leg['e2e'] = {'value': e2e_updates / e2e_s_, 'unit': 'updates/s', 'h2d_bytes_per_step': 8 * B_local, 'd2h_bytes_per_step': 4,
                          'note': 'as RL_AC.learn_and_update runs it: the index draws of the loop in chunks of ~32 k indices, one np.random call + one H2D each (ReplayBuffer.index_stream: same stream), per update a device gather into the graph inputs + '
                                  'the update replayed as a CUDA graph (' + ('software-pipelined over consecutive updates, RL.PipelinedUpdateGraph' if hasattr(ug_, 'flush') else 'sequential') + ') + the loss of EVERY update read back (copy issued behind the update, waited for after the next one is launched); '
                                  f'last loss {loss_host:.4g}'}
        del r_, nn_
        return leg

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))
    except Exception:
        pass
    bf16_peak = peaks.get('bf16_tflops') or 1590.0
    peak_src = 'MEASURED_PEAKS.json bf16_tflops (burst)' if peaks.get('bf16_tflops') else 'fallback 1590 TFLOP/s'
    upd_legs = [update_leg(SYSTEM, conf.BATCH_SIZE, 200, 1000, e2e_updates=200)]
    if not args.quick:
        upd_legs.append(update_leg(SYSTEM, 4096, 30, 100))
        upd_legs.append(update_leg(SYSTEM, 16384, 20, 50, e2e_updates=30))
        for Bg in (64, 4096, 16384):                  # BASELINE config 5: UR5, global batch split over the ranks
            if Bg % world == 0:
                upd_legs.append(update_leg('ur5', Bg // world, 20 if Bg > 64 else 100, 50 if Bg > 64 else 300))
    Bu = conf.BATCH_SIZE
    updates_per_s = 1e6 / upd_legs[0]['us_per_update_eager']
    graph_us0 = upd_legs[0]['us_per_update_cuda_graph']
    graph_updates_per_s = 1e6 / graph_us0 if isinstance(graph_us0, float) else graph_us0
    pipe_us0 = upd_legs[0].get('us_per_update_pipelined_graph')
    pipelined_updates_per_s = 1e6 / pipe_us0 if isinstance(pipe_us0, float) else pipe_us0

    # data-parallel correctness, visible to the driver: one update from identical weights through the NVLink peer-memory exchange and
    # through NCCL all-reduces must agree (1e-4, the parity gate), and the replicas must stay bit-identical
    dp_check = None
    if world > 1:
        try:
            ra = RL_AC(env, NN(env, conf, 1e-2, seed=0), conf, 0, dist=dist)
            ra.setup_model()
            rb = RL_AC(env, NN(env, conf, 1e-2, seed=0), conf, 0, dist=dist, reduce='nccl')
            rb.setup_model()
            for na_, nb_ in ((ra.actor_model, rb.actor_model), (ra.critic_model, rb.critic_model), (ra.target_critic, rb.target_critic)):
                nb_.params.copy_(na_.params)
                nb_.refresh_transposed()
            bt = [torch.as_tensor(x).to(dev) for x in _update_batch(conf, Bu, 100 + rank)]
            s_, pr_, sn_, dv_, d_, term_, w_ = bt
            for r_ in (ra, rb):
                for _ in range(2):
                    r_.update(s_, sn_, pr_, dv_, d_, term_, w_, fuse_target=True)
            torch.cuda.synchronize()
            worst = 0.0
            for na_, nb_ in ((ra.actor_model, rb.actor_model), (ra.critic_model, rb.critic_model), (ra.target_critic, rb.target_critic)):
                worst = max(worst, float((na_.params - nb_.params).abs().max() / nb_.params.abs().max()))
            gathered = [torch.empty_like(ra.actor_model.params) for _ in range(world)]
            dist.all_gather(gathered, ra.actor_model.params)
            identical = all(bool(torch.equal(gathered[0], g_)) for g_ in gathered[1:])
            ok = worst <= 1e-4 and identical
            flag = torch.tensor([1 if ok else 0], dtype=torch.int32, device=dev)
            dist.all_reduce(flag, op=dist.ReduceOp.MIN)
            exch = 'peer' if ra._peer is not None else 'nccl (peer unavailable)'
            dp_check = ('ok' if int(flag[0]) == 1 else 'FAILED') + f': {exch} vs nccl max rel diff {worst:.2e} after 2 updates, replicas bit-identical: {identical}'
            del ra, rb
        except Exception as exc:
            dp_check = f'FAILED: {type(exc).__name__}: {exc}'
    large_batch = {str(l['batch_per_gpu']): {'us_per_update': l['us_per_update_eager'], 'us_per_update_cuda_graph': l['us_per_update_cuda_graph'],
                                              'samples_per_s': l['samples_per_s'], 'engine': l['engine'],
                                              'algorithmic_tflops': l['roofline']['achieved']}
                   for l in upd_legs if l['system'] == SYSTEM and l['batch_per_gpu'] > 64}
    clocks = sampler.stop() if rank == 0 else None

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))
        except Exception:
            pass
        bf16 = peaks.get('bf16_tflops') or 1590.0
        peak_src = 'MEASURED_PEAKS.json bf16_tflops (burst)' if peaks.get('bf16_tflops') else 'fallback 1590 TFLOP/s'
        if args.engine == 'tc':
            roof = {'bound': 'tensor', 'achieved': achieved_tflops, 'peak': bf16, 'unit': 'TFLOP/s', 'frac': achieved_tflops / bf16,
                    'traffic': TRAFFIC_BYTES.get(args.engine), 'traffic_source': TRAFFIC_SOURCE.get(args.engine), 'peak_source': peak_src,
                    'ceiling_3xfp16': bf16 / 3.0, 'frac_of_3xfp16_ceiling': achieved_tflops / (bf16 / 3.0),
                    'note': 'k_rollout_tc16: 256x256 actor layer on tcgen05.mma kind::f16 with fp16 hi/lo operand splitting (3 UMMAs per logical '
                            'product, fp32 accumulation in TMEM, fp32-class accuracy: parity gate 1e-5), persistent CTAs with two free-running '
                            'tile pipelines. achieved counts ALGORITHMIC flops (2 P_a + F_dyn per env-step); the reachable ceiling is peak/3.',
                    'flops_per_env_step': FLOPS_PER_ENV_STEP, 'kernel_ms': kernel_ms}
        elif args.engine == 'tf32':
            roof = {'bound': 'tensor', 'achieved': achieved_tflops, 'peak': bf16, 'unit': 'TFLOP/s', 'frac': achieved_tflops / bf16,
                    'traffic': TRAFFIC_BYTES.get(args.engine), 'peak_source': peak_src,
                    'ceiling_3xtf32': bf16 / 6.0, 'frac_of_3xtf32_ceiling': achieved_tflops / (bf16 / 6.0),
                    'note': 'k_rollout_tc: 256x256 actor layer on tcgen05.mma kind::tf32 with 3xTF32 operand splitting; tf32 runs at half the '
                            'bf16 rate and 3 UMMAs are issued per logical product, so the reachable ceiling is peak/6.',
                    'flops_per_env_step': FLOPS_PER_ENV_STEP, 'kernel_ms': kernel_ms}
        else:
            roof = {'bound': 'fp32_fma', 'achieved': achieved_tflops, 'peak': fma_peak_tflops, 'unit': 'TFLOP/s',
                    'frac': achieved_tflops / fma_peak_tflops, 'traffic': None, 'peak_source': 'cacto_peak_fma_fp32 measured in this run',
                    'frac_of_bf16_tensor_peak': achieved_tflops / bf16,
                    'note': 'k_rollout: fp32 CUDA-core FMA contraction', 'flops_per_env_step': FLOPS_PER_ENV_STEP, 'kernel_ms': kernel_ms}
        line = {
            'metric': 'manipulator rollout env-steps/s', 'value': value, 'unit': 'env-steps/s', 'n_gpus': world, 'steps': K, 'warmup': W,
            'ms_per_step': total_ms / K, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
            'dtype': {'tc': 'f32 actor MLP (fp16-split x3 on tcgen05, fp32 accumulate) / f64 dynamics', 'tf32': 'f32 actor MLP (3xTF32 tensor cores) / f64 dynamics',
                      'fma': 'f32 actor MLP / f64 dynamics'}[args.engine],
            'data': 'synthetic',
            'config': {'workload': 'BASELINE config[3]: 3-DOF planar manipulator policy rollouts (create_TO_init), '
                                   f'{B} rollouts x {T} steps per GPU (1 M over 8 GPUs), seeded-init actor {ns}->256->256->{na}',
                       'rollouts_per_gpu': B, 'horizon': T, 'engine': args.engine,
                       'l2': 'flushed between timed steps (256 MiB write); outputs 1.06 GB/step',
                       'parallelism': f'dp{world} over independent rollouts, no collective on the rollout path',
                       'host_placement_rank0': placement},
            'roofline': roof,
            'e2e': {'value': e2e_value, 'unit': 'env-steps/s', 'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': d2h,
                    'note': 'RL_AC.rollout_to_host (pipelined): 8 sub-batches, the copy engine moves the fp64 trajectories of sub-batch k into pinned host memory '
                            'while sub-batch k+1 is rolled out; consecutive batches alternate between two sets of pinned host buffers (batch k+1 queued before batch k is waited for); '
                            'PCIe-bound (1.06 GB per step)'},
            'gpu_launches': K * launches_per_step,          # device-resident leg; the e2e leg launches 2 + 8 kernels per step
            'clocks': clocks,
            'extra': {'e2e_compact': {'value': e2e_compact_value, 'unit': 'env-steps/s', 'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': d2h_compact,
                                      'note': 'rollout_to_host(compact=True): the same batches without the time row (t_k = t_0 + k dt by repeated addition) and with '
                                              'the controls as the fp32 values the actor produced; RL.CompactRollouts rebuilds the fp64 arrays bit-identically '
                                              '(tests/test_gpu_rollout.py::test_rollout_to_host_compact_is_bit_reconstructible); not the headline e2e'},
                      'update': {'metric': 'Sobolev actor-critic updates/s (RL_AC.update: critic gradient, Adam + Polyak, actor gradient, Adam)',
                                 'legs': upd_legs, 'dp_check': dp_check,
                                 'cpu_baseline': (cpu or {}).get('update'),
                                 'note': 'updates_per_s = CUDA-graph replay where capture is possible, else eager; value of a leg counts updates of '
                                         'the GLOBAL batch (all ranks step together); roofline.achieved = algorithmic flops 2 B (10 P_c + 3 P_a) / time'},
                      'cpu_baselines_C2_C6': (cpu or {}).get('misc'),
                      'parity': 'oracle pinned by reference-executed goldens for PER / reward-to-go / analytic systems / rewards / SI-car-car_park '
                                'backward pass; UNPINNED for pinocchio-backed dynamics and tensorflow update semantics (absent here): '
                                'tests/golden/make_golden_ext.py is the kit that pins them where the wheels exist',
                      'sobolev_updates_per_s': updates_per_s, 'sobolev_updates_per_s_cuda_graph': graph_updates_per_s,
                      'sobolev_updates_per_s_pipelined_graph': pipelined_updates_per_s,   # one GPU: RL.PipelinedUpdateGraph (None with more ranks)
                      'update_batch_per_gpu': Bu, 'update_global_batch': Bu * world, 'update_kernels_per_update': 5,      # schedules (one launch for both optimizers), critic gradient, Adam + Polyak, actor gradient, Adam
                      
                      'update_gradient_exchange': ('none (1 GPU)' if world == 1 else
                                                   'NVLink peer-memory sum inside the Adam kernels (k_adam_peer)' if rl._peer is not None else
                                                   'NCCL all-reduce per network'),
                      'fp32_fma_peak_tflops_measured': fma_peak_tflops, 'sobolev_update_large_batch': large_batch},
        }
        if cpu is not None:
            line['cpu_baseline'] = cpu
        emit(line)
    if world > 1:
        dist.destroy_process_group()


_REAL_STDOUT = None


def emit(line):
    """The ONE JSON line goes to the process's original stdout; everything else any library prints to fd 1 during the run
    (NCCL writes its version banner there) has been redirected to stderr."""
    data = (json.dumps(line) + '\n').encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode()); sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


def main():
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)                      # fd 1 -> stderr for the duration of the run
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=5)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--engine', default='tc', choices=['tc', 'tf32', 'fma'])
    ap.add_argument('--rollouts-per-gpu', type=int, default=ROLLOUTS_PER_GPU)
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--quick', action='store_true', help='rollout legs and the conf-batch update only')
    args = ap.parse_args()
    if args.impl == 'reference':
        run_reference(args)
    else:
        run_b200(args)


if __name__ == '__main__':
    main()
