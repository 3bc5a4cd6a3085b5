#!/usr/bin/env python
"""BASELINE config 5, one GPU's shard: batch 8 x 192^3 fp32, cubic, per-dim bounds (dct2, dft, zero),
grid_pull + grid_push; checks tiled/persistent kernels against the generic ones on batch element 0."""
import os, sys, statistics
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'torch-interpol_b200')); sys.path.insert(0, ROOT)
import torch
from bench import make_workload
import interpol_b200 as ib
from interpol_b200 import pushpull as pp

B, n = 8, 192
vols, grids = zip(*[make_workload(n, 'cuda', seed=100 + b) for b in range(B)])
vol, grid = torch.cat(vols).contiguous(), torch.cat(grids).contiguous()
bound, order = [3, 6, 0], [3]


def timeit(fn, reps=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    return statistics.median(ts)


N = B * n ** 3
for name, fn in (('pull', lambda: pp.grid_pull(vol, grid, bound, order, 1)), ('push', lambda: pp.grid_push(vol, grid, [n] * 3, bound, order, 1))):
    ms = timeit(fn)
    print('%s  B=%d %d^3 bounds (dct2, dft, zero): %.3f ms  %.0f Mvox/s  %.1f %% of HBM roofline (20 B/voxel)  [%s]' % (
        name, B, n, ms, N / ms / 1e3, 100 * N * 20 / (ms * 1e-3) / 1e9 / 6543.7, ib.last_kernel()))
    out = fn()[:1]
    pp.flags = 1
    try:
        ref = (pp.grid_pull(vol[:1], grid[:1], bound, order, 1) if name == 'pull' else pp.grid_push(vol[:1], grid[:1], [n] * 3, bound, order, 1))
    finally:
        pp.flags = 0
    print('   max |tiled - generic| / max|generic| on batch element 0: %.2e' % ((out - ref).abs().max().item() / ref.abs().max().item()))
