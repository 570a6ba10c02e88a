"""Sobolev update at the reference's batch sizes: eager launches and CUDA-graph replays, CUDA events.
python profiles/scripts/update_small_batch.py [system]"""
import sys; sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import torch
from test_gpu_peer_reduce import batch, make_rl, IO_KEYS
system = sys.argv[1] if len(sys.argv) > 1 else 'manipulator'
for B in (64, 128, 256, 512):
    conf, rl = make_rl(system)
    g = batch(conf, B, 0)
    ug = rl.make_update_graph(B)
    for k in IO_KEYS: ug.io[k].copy_(g[k])
    out = {}
    for name, fn, n in (('eager', lambda: rl.update(*[g[k] for k in IO_KEYS], fuse_target=True), 200), ('graph', ug.replay, 1000)):
        for _ in range(20): fn()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(n): fn()
        b.record(); torch.cuda.synchronize()
        out[name] = a.elapsed_time(b) / n * 1e3
    print(f'{system} B={B}: eager {out["eager"]:.1f} us, graph {out["graph"]:.1f} us = {1e6/out["graph"]:.0f} updates/s', flush=True)
