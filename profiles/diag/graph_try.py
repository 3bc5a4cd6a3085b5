import os, sys, time
sys.path.insert(0, '/root/repo/torch-interpol_b200'); sys.path.insert(0, '/root/repo/tests'); sys.path.insert(0, '/root/repo')
import torch
import interpol_b200 as ib
from test_gpu_ops import smooth_grid
gen = torch.Generator().manual_seed(0)
for shape in ((64, 64, 64), (128, 128, 128)):
    vol = torch.randn([1, 1, *shape], generator=gen).cuda()
    grid = smooth_grid(shape, gen, amp=2.0).contiguous().cuda()
    kw = dict(interpolation=3, bound='dct2', extrapolate=True)
    def step():
        p = ib.grid_pull(vol, grid, **kw)
        g = ib.grid_grad(vol, grid, **kw)
        s = ib.grid_push(p, grid, **kw)
        c = ib.spline_coeff_nd(s, interpolation=3, bound='dct2', dim=3)
        return p, g, s, c
    s_ = torch.cuda.Stream()
    with torch.cuda.stream(s_):
        for _ in range(3): step()
    torch.cuda.current_stream().wait_stream(s_)
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        outs = step()
    # new contents, same buffers
    vol.copy_(torch.randn(vol.shape, generator=gen)); grid.add_(0.37)
    graph.replay(); torch.cuda.synchronize()
    want = step(); torch.cuda.synchronize()
    print(shape, 'graph == eager:', [bool(torch.equal(a, b)) or float((a - b).abs().max()) for a, b in zip(outs, want)])
    def wall(fn, n=300):
        for _ in range(20): fn()
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for _ in range(n): fn()
        torch.cuda.synchronize(); return (time.perf_counter() - t0) / n * 1e6
    print(shape, 'eager %.1f us per step, graph replay %.1f us per step' % (wall(step), wall(graph.replay)))
