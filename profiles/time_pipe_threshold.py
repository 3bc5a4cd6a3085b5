#!/usr/bin/env python
"""Persistent kernel vs one-tile-per-CTA kernel on small lattices (where does the dispatch threshold belong?):
20 launches back to back, cubic / linear dct2 pull and grad on the benchmark deformation at several edge lengths."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'torch-interpol_b200')); sys.path.insert(0, ROOT)
import torch
from bench import make_workload
import interpol_b200 as ib
from interpol_b200 import pushpull as pp


def timeit(fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


flush = torch.empty(64 << 20, dtype=torch.float32, device='cuda')
for n in (128, 160, 192, 224, 256):
    vol, grid = make_workload(n, 'cuda')
    tiles = (n // 8) ** 2 * ((n + 31) // 32)
    for order in (3, 1):
        for name, fn in (('pull', lambda: pp.grid_pull(vol, grid, [3], [order], 1)), ('grad', lambda: pp.grid_grad(vol, grid, [3], [order], 1))):
            res = []
            for flags in (4, 8):
                pp.flags = flags
                res.append((timeit(fn), ib.last_kernel()))
            pp.flags = 0
            print('%3d^3 (%5d tiles, %4.1f per SM) order %d %-4s tile kernel %.4f ms [%s]   persistent %.4f ms [%s]'
                  % (n, tiles, tiles / 148, order, name, res[0][0], res[0][1], res[1][0], res[1][1]), flush=True)
