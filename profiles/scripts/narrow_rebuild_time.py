"""Event timing of the two kernels added at the end of round 2: k_narrow_f64 (HBM-bound, 12 B per element) and k_segtree_rebuild."""
import sys; sys.path.insert(0, '/root/repo')
import json, torch
from cacto_b200._lib import lib, ptr, stream_ptr, check
peaks = json.load(open('/root/repo/MEASURED_PEAKS.json')) if __import__('os').path.exists('/root/repo/MEASURED_PEAKS.json') else {}
flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')
def timed(fn, reps=20, do_flush=True):
    for _ in range(3): fn()
    ts = []
    for _ in range(reps):
        if do_flush: flush.fill_(1)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b) * 1e3)
    ts.sort(); return ts[len(ts) // 2]
for n in (100 * 3 * 16384, 100 * 3 * 131072):
    src = torch.randn(n, dtype=torch.float32, device='cuda').double(); dst = torch.empty(n, dtype=torch.float32, device='cuda')
    us = timed(lambda: check(lib.cacto_narrow_f64_to_f32(ptr(src), ptr(dst), n, stream_ptr()), 'narrow'))
    assert torch.equal(dst.double(), src)
    print(f'k_narrow_f64 n={n}: {us:.1f} us, {12 * n / us / 1e3:.0f} GB/s algorithmic (L2 flushed), HBM peak {peaks}')
cap = 1 << 16
for B in (64, 4096):
    s = torch.zeros(2 * cap, dtype=torch.float64, device='cuda'); m = torch.full((2 * cap,), float('inf'), dtype=torch.float64, device='cuda')
    st = torch.full((cap,), -1, dtype=torch.int32, device='cuda')
    idx = torch.randint(0, cap, (B,), device='cuda'); val = torch.rand(B, dtype=torch.float64, device='cuda') + 0.1
    us = timed(lambda: check(lib.cacto_segtree_update(ptr(s), ptr(m), cap, ptr(idx), ptr(val), B, ptr(st), stream_ptr()), 'upd'), do_flush=False)
    print(f'k_segtree_rebuild cap=65536 B={B}: {us:.1f} us (trees L2-resident)')
