#!/usr/bin/env python
"""resize 256^3 -> 384^3 / 128^3 (cubic, dct2, prefilter): separable passes vs the dense-grid path."""
import os, sys, statistics
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'torch-interpol_b200')); sys.path.insert(0, ROOT)
import torch
import interpol_b200 as ib
import importlib
rz = importlib.import_module('interpol_b200.resize')      # (the package attribute `resize` is the function)


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
for factor in (1.5, 0.5):
    for prefilter in (False, True):
        fn = lambda: ib.resize(x, factor=[factor] * 3, anchor='e', interpolation=3, bound='dct2', prefilter=prefilter)
        rz.SEPARABLE = True
        sep = timeit(fn); a = fn()
        rz.SEPARABLE = False
        dense = timeit(fn); b = fn()
        rz.SEPARABLE = True
        nout = a.numel()
        print('256^3 x %.1f prefilter=%d: separable %.3f ms (%.0f Mvox/s out), dense grid + pull %.3f ms, ratio %.1fx, max rel diff %.1e' % (
            factor, prefilter, sep, nout / sep / 1e3, dense, dense / sep, (a - b).abs().max().item() / b.abs().max().item()))
