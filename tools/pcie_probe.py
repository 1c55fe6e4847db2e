import torch, time
x = torch.empty(256 * 1024 * 1024 // 4).pin_memory()
d = torch.empty_like(x, device='cuda')
for name, fn in (('H2D', lambda: d.copy_(x, non_blocking=True)), ('D2H', lambda: x.copy_(d, non_blocking=True))):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(5): fn()
    torch.cuda.synchronize()
    print(name, f'{5 * 0.25 / (time.perf_counter() - t0):.1f} GiB/s')
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
y = torch.empty_like(x).pin_memory(); e = torch.empty_like(d)
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(5):
    with torch.cuda.stream(s1): d.copy_(x, non_blocking=True)
    with torch.cuda.stream(s2): y.copy_(e, non_blocking=True)
torch.cuda.synchronize()
print('both directions concurrently', f'{5 * 0.5 / (time.perf_counter() - t0):.1f} GiB/s total')
