#!/usr/bin/env python
"""restrict 256^3 -> 128^3 (linear / cubic): adjoint separable passes vs the dense-grid path."""
import importlib, os, sys, statistics
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'torch-interpol_b200')); sys.path.insert(0, ROOT)
import torch
import interpol_b200 as ib
rs = importlib.import_module('interpol_b200.restrict')


def timeit(fn, reps=8):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    return statistics.median(ts)


x = torch.randn(1, 1, 256, 256, 256, device='cuda')
for order in (1, 3):
    fn = lambda: ib.restrict(x, factor=[2, 2, 2], anchor='e', interpolation=order, bound='dct2')
    rs.SEPARABLE = True
    sep = timeit(fn); a = fn()
    rs.SEPARABLE = False
    dense = timeit(fn); b = fn()
    rs.SEPARABLE = True
    print('restrict 256^3 / 2, order %d: separable %.3f ms, dense grid + push %.3f ms, ratio %.1fx, max rel diff %.1e' % (
        order, sep, dense, dense / sep, (a - b).abs().max().item() / b.abs().max().item()))
