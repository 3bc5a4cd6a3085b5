import torch, time
x = torch.empty(256**3 * 3, dtype=torch.float32).pin_memory()
y = torch.empty(256**3, dtype=torch.float32).pin_memory()
d = torch.empty_like(x, device='cuda'); e = torch.empty_like(y, device='cuda')
for name, fn, nbytes in (('H2D 201MB pinned', lambda: d.copy_(x, non_blocking=True), x.numel()*4), ('D2H 67MB pinned', lambda: y.copy_(e, non_blocking=True), y.numel()*4)):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(5): fn()
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / 5
    print('%s: %.2f ms, %.1f GB/s' % (name, dt * 1e3, nbytes / dt / 1e9))
t0 = time.perf_counter()
for _ in range(5):
    h = torch.empty(256**3, dtype=torch.float32, pin_memory=True); h.copy_(e, non_blocking=True); torch.cuda.synchronize()
print('alloc pinned + D2H: %.2f ms' % ((time.perf_counter() - t0) / 5 * 1e3))
t0 = time.perf_counter()
for _ in range(5):
    dd = x.to('cuda', non_blocking=True); torch.cuda.synchronize()
print('to(cuda) 201MB: %.2f ms' % ((time.perf_counter() - t0) / 5 * 1e3))
