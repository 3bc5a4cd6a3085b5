#!/usr/bin/env python
"""Sensitivity of the persistent pull to the local stretch of the deformation along z (shared-memory bank conflicts:
a row of 32 voxels whose support starts span more than 32 words costs two wavefronts per tap): 256^3 cubic dct2,
grid = centre + s * (identity - centre) along z, identity (+0.3) along x / y."""
import os, sys, statistics
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'torch-interpol_b200')); sys.path.insert(0, ROOT)
import torch
import interpol_b200 as ib
from interpol_b200 import pushpull as pp

n = 256
vol = torch.randn([1, 1, n, n, n], device='cuda')
ident = ib.identity_grid([n, n, n], device='cuda')[None]


def timeit(fn, reps=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


for axis in (2, 1):
    for s in (0.8, 0.95, 1.0, 1.05, 1.12, 1.3, 1.6):
        grid = ident.clone() + 0.3
        grid[..., axis] = (ident[..., axis] - n / 2) * s + n / 2 + 0.3
        grid = grid.contiguous()
        for name, fn in (('pull', lambda: pp.grid_pull(vol, grid, [3], [3], 1)), ('push', lambda: pp.grid_push(vol, grid, [n] * 3, [3], [3], 1))):
            ms = timeit(fn)
            print('stretch %.2f along %s  %-4s %.3f ms  [%s]' % (s, 'xyz'[axis], name, ms, ib.last_kernel()), flush=True)
