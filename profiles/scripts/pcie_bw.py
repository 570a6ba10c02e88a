import torch, time
n = 1056440320 // 8
d = torch.empty(n, dtype=torch.float64, device='cuda'); h = torch.empty(n, dtype=torch.float64).pin_memory()
for name, fn in (('D2H', lambda: h.copy_(d, non_blocking=True)), ('H2D', lambda: d.copy_(h, non_blocking=True))):
    for _ in range(2): fn()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(5): fn()
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 5
    print(name, f'{n*8/dt/1e9:.1f} GB/s', f'{dt*1e3:.2f} ms')
# 8 chunked copies on a side stream
chunks_d = list(d.chunk(8)); chunks_h = list(h.chunk(8))
s = torch.cuda.Stream()
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(5):
    with torch.cuda.stream(s):
        for a, b in zip(chunks_h, chunks_d): a.copy_(b, non_blocking=True)
torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 5
print('D2H 8 chunks', f'{n*8/dt/1e9:.1f} GB/s')
