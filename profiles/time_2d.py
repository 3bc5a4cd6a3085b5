#!/usr/bin/env python
"""2-D problems (generic one-thread-per-point kernels): device time and fraction of the HBM roofline."""
import os, sys, statistics
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'torch-interpol_b200')); sys.path.insert(0, ROOT)
import torch
from bench import measured_peak
import interpol_b200 as ib
from interpol_b200 import pushpull as pp

peak, _ = measured_peak()


def timeit(fn, n=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return statistics.median(ts)


for n in (1024, 4096):
    gen = torch.Generator(device='cuda').manual_seed(0)
    vol = torch.randn([1, 1, n, n], device='cuda', generator=gen)
    coarse = torch.randn([1, 2, 16, 16], device='cuda', generator=gen) * 3
    disp = torch.nn.functional.interpolate(coarse, size=[n, n], mode='bilinear', align_corners=True).permute(0, 2, 3, 1)
    grid = (ib.identity_grid([n, n], device='cuda')[None] + disp).contiguous()
    for order in (1, 3):
        for name, fn, by in (('pull', lambda: pp.grid_pull(vol, grid, [3], [order], 1), n * n * 16),
                             ('push', lambda: pp.grid_push(vol, grid, [n, n], [3], [order], 1), n * n * 16),
                             ('grad', lambda: pp.grid_grad(vol, grid, [3], [order], 1), n * n * 24)):
            ms = timeit(fn)
            print('2-D %d^2 order %d %-5s %8.3f ms  %8.0f Mpix/s  %5.1f %% of %.0f GB/s  [%s]'
                  % (n, order, name, ms, n * n / ms / 1e3, 100 * by / (ms * 1e-3) / 1e9 / peak, peak, ib.last_kernel()))
