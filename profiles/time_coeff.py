#!/usr/bin/env python
"""Per-axis timing of the spline prefilter (256^3 fp32, cubic, dct2): which of the three passes costs what."""
import os, sys, statistics
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'torch-interpol_b200')); sys.path.insert(0, ROOT)
import torch
import interpol_b200 as ib

def timeit(fn, n=20, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    return statistics.median(ts)

for C in (1, 4):
    x = torch.randn(1, C, 256, 256, 256, device='cuda')
    for order in (3, 5):
        line = []
        for d in (-3, -2, -1):
            t = timeit(lambda: ib.spline_coeff(x, interpolation=order, bound='dct2', dim=d, inplace=True))
            line.append('axis %d: %.3f ms (%s)' % (d, t, ib.last_kernel()))
        t = timeit(lambda: ib.spline_coeff_nd(x, interpolation=order, bound='dct2', dim=3, inplace=True))
        tc = timeit(lambda: x.clone())
        print('C=%d order %d | %s | nd in place %.3f ms | clone alone %.3f ms | copy roofline (1R+1W) %.3f ms' % (
            C, order, ' | '.join(line), t, tc, 2 * x.numel() * 4 / 6.5437e12 * 1e3))
