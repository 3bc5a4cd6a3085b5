#!/usr/bin/env python
"""Back-to-back device time of the persistent pull / grad (20 launches between two events, no host gaps): the
benchmark deformation and an unstretched one (identity * 0.95 + 0.3: no shared-memory bank conflicts)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'torch-interpol_b200')); sys.path.insert(0, ROOT)
import torch
from bench import make_workload
import interpol_b200 as ib
from interpol_b200 import pushpull as pp

n = 256
vol, grid = make_workload(n, 'cuda')
ident = ib.identity_grid([n, n, n], device='cuda')[None]
flat = ((ident - n / 2) * 0.95 + n / 2 + 0.3).contiguous()


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


tag = os.environ.get('IB200_LIB', 'default').split('/')[-1]
for order in (3, 1):
    for gname, g in (('bench', grid), ('flat', flat)):
        for name, fn in (('pull', lambda: pp.grid_pull(vol, g, [3], [order], 1)), ('grad', lambda: pp.grid_grad(vol, g, [3], [order], 1))):
            print('%-16s order %d %-5s grid %-5s %.4f ms [%s]' % (tag, order, name, gname, timeit(fn), ib.last_kernel()), flush=True)
