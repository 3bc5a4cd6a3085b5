#!/usr/bin/env python
"""Host-side cost of one call: wall time per call of interpol_b200.grid_pull (public API) and of
interpol_b200.pushpull.grid_pull (binding layer) on problems small enough to be launch-bound."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'torch-interpol_b200')); sys.path.insert(0, ROOT)
import torch
import interpol_b200 as ib
from interpol_b200 import pushpull as pp


def wall(fn, n=2000):
    for _ in range(50):
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n * 1e6


for name, vshape, dim in (('2-D 256^2 linear (cfg 1)', (256, 256), 2), ('3-D 32^3 cubic', (32, 32, 32), 3), ('3-D 64^3 cubic', (64, 64, 64), 3)):
    vol = torch.randn([1, 1, *vshape], device='cuda')
    grid = ib.identity_grid(vshape, device='cuda')[None] + 0.3
    order = 1 if dim == 2 else 3
    a = wall(lambda: ib.grid_pull(vol, grid, interpolation=order, bound='dct2', extrapolate=True))
    b = wall(lambda: pp.grid_pull(vol, grid, [3], [order], 1))
    c = wall(lambda: ib.grid_push(vol, grid, interpolation=order, bound='dct2', extrapolate=True))
    print('%-26s grid_pull: API %.1f us, binding %.1f us per call; grid_push API %.1f us' % (name, a, b, c))
