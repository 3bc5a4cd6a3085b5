#!/usr/bin/env python
"""label-map pull, 256^3 int64 with 16 labels, linear: one-pass kernel vs the loop over labels."""
import os, sys, statistics
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'torch-interpol_b200')); sys.path.insert(0, ROOT)
import torch
from bench import make_workload
import interpol_b200 as ib
import interpol_b200.api as api

_, grid = make_workload(256, 'cuda')
lab = torch.randint(0, 16, [1, 1, 256, 256, 256], device='cuda')


def timeit(fn, reps=5):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    return statistics.median(ts)


fn = lambda: ib.grid_pull(lab, grid, interpolation=1, bound='dct2', extrapolate=True)
fused = timeit(fn); a = fn()
api.LABELS_FUSED = False
loop = timeit(fn); b = fn()
api.LABELS_FUSED = True
print('label pull 256^3, 16 labels, linear: one pass %.3f ms, loop over labels %.3f ms (%.0fx), mismatching voxels %d of %d' % (
    fused, loop, loop / fused, (a != b).sum().item(), a.numel()))
