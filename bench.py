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
TRAFFIC_BYTES = {'tc': 8869376 + 998446080, 'tf32': 8953600 + 998347520}


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

    import torch
    import torch.distributed as dist
    assert torch.cuda.is_available(), 'bench.py needs a CUDA device (no CPU fallback on the product path)'
    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
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
    st_host = torch.empty((T + 1, ns, B), dtype=torch.float64).pin_memory()
    ct_host = torch.empty((T, na, B), dtype=torch.float64).pin_memory()
    fl_host = torch.empty(B, dtype=torch.int32).pin_memory()
    for _ in range(max(1, W // 2)):
        rl.rollout_to_host(ics_host, 1, st_host, ct_host, fl_host)
    barrier()
    t0 = time.perf_counter()
    for _ in range(K):
        rl.rollout_to_host(ics_host, 1, st_host, ct_host, fl_host)
    torch.cuda.synchronize()
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    assert bool(fl_host.all())
    e2e_value = B * T * K * world / e2e_s
    h2d = ics_host.numel() * 8 + B * 4
    d2h = (st_host.numel() + ct_host.numel()) * 8 + fl_host.numel() * 4

    # ---- K3 updates/s (secondary metric): conf batch per GPU, data resident, gradients summed across GPUs inside the Adam kernels
    Bu = conf.BATCH_SIZE
    g = torch.Generator(device='cpu').manual_seed(rank)
    lo, hi = torch.as_tensor(conf.x_init_min), torch.as_tensor(conf.x_init_max)
    s = (lo + (hi - lo) * torch.rand((Bu, ns), generator=g, dtype=torch.float64)).float().to(dev)
    sn = (lo + (hi - lo) * torch.rand((Bu, ns), generator=g, dtype=torch.float64)).float().to(dev)
    pr = (-5 * torch.rand((Bu, 1), generator=g)).to(dev)
    dv = torch.randn((Bu, ns), generator=g).to(dev); dv[:, -1] = 0
    d = (torch.rand((Bu, 1), generator=g) < 0.5).float().to(dev)
    term = (torch.rand((Bu, 1), generator=g) < 0.01).double().to(dev)
    w = torch.ones((Bu, 1), device=dev)

    def update_step():
        rl.update(s, sn, pr, dv, d, term, w, fuse_target=True)
    n_up = 200
    for _ in range(20):
        update_step()
    barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(stream)
    for _ in range(n_up):
        update_step()
    b.record(stream)
    torch.cuda.synchronize()
    updates_per_s = n_up / (max_over_ranks(a.elapsed_time(b)) * 1e-3)
    # the same update replayed as a CUDA graph (6 kernel nodes; data-parallel: the gradient exchange happens inside the two
    # Adam nodes over NVLink peer memory, so the graph holds no collective)
    graph_updates_per_s, n_gr = None, 1000
    if world == 1 or rl._peer is not None:
        try:
            ug = rl.make_update_graph(Bu)
            for k_, t_ in (('state', s), ('state_next', sn), ('partial_rtg', pr), ('dVdx', dv), ('done', d), ('term', term), ('weights', w)):
                ug.io[k_].copy_(t_)
            for _ in range(20):
                ug.replay()
            barrier()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(stream)
            for _ in range(n_gr):
                ug.replay()
            b.record(stream)
            torch.cuda.synchronize()
            graph_ms = a.elapsed_time(b)
            del ug
        except Exception as exc:           # report, do not hide
            graph_ms, graph_updates_per_s = -1.0, f'failed: {type(exc).__name__}: {exc}'
        graph_ms_all = max_over_ranks(graph_ms)
        if graph_ms > 0:
            graph_updates_per_s = n_gr / (graph_ms_all * 1e-3)
    # the update at the batch sizes of BASELINE configs 2 / 3 (PER batch 4096, critic batch 16384), single GPU, eager launches
    large_batch = {}
    if world == 1:
        for Bl in (4096, 16384):
            gl = torch.Generator(device='cpu').manual_seed(Bl)
            sl = (lo + (hi - lo) * torch.rand((Bl, ns), generator=gl, dtype=torch.float64)).float().to(dev)
            snl = (lo + (hi - lo) * torch.rand((Bl, ns), generator=gl, dtype=torch.float64)).float().to(dev)
            prl = (-5 * torch.rand((Bl, 1), generator=gl)).to(dev)
            dvl = torch.randn((Bl, ns), generator=gl).to(dev); dvl[:, -1] = 0
            dl = (torch.rand((Bl, 1), generator=gl) < 0.5).float().to(dev)
            tl = (torch.rand((Bl, 1), generator=gl) < 0.01).double().to(dev)
            wl = torch.ones((Bl, 1), device=dev)
            for _ in range(5):
                rl.update(sl, snl, prl, dvl, dl, tl, wl, fuse_target=True)
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(stream)
            for _ in range(30):
                rl.update(sl, snl, prl, dvl, dl, tl, wl, fuse_target=True)
            b.record(stream)
            torch.cuda.synchronize()
            us = a.elapsed_time(b) * 1e3 / 30
            large_batch[str(Bl)] = {'us_per_update': us, 'samples_per_s': Bl / (us * 1e-6),
                                    'algorithmic_tflops': Bl * 0.994e6 / (us * 1e-6) / 1e12}
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
                    'traffic': TRAFFIC_BYTES.get(args.engine), 'peak_source': peak_src,
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
                       'parallelism': f'dp{world} over independent rollouts, no collective on the rollout path'},
            'roofline': roof,
            'e2e': {'value': e2e_value, 'unit': 'env-steps/s', 'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': d2h,
                    'note': 'RL_AC.rollout_to_host (pipelined): 8 sub-batches, the copy engine moves the fp64 trajectories of sub-batch k into pinned host memory '
                            'while sub-batch k+1 is rolled out; PCIe-bound (1.06 GB per step)'},
            'gpu_launches': K * launches_per_step,          # device-resident leg; the e2e leg launches 2 + 8 kernels per step
            'clocks': clocks,
            'extra': {'sobolev_updates_per_s': updates_per_s, 'sobolev_updates_per_s_cuda_graph': graph_updates_per_s,
                      'update_batch_per_gpu': Bu, 'update_global_batch': Bu * world, 'update_kernels_per_update': 6,
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
    args = ap.parse_args()
    if args.impl == 'reference':
        run_reference(args)
    else:
        run_b200(args)


if __name__ == '__main__':
    main()
