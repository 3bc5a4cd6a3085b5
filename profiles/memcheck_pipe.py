#!/usr/bin/env python
"""Small problems for `compute-sanitizer --tool memcheck python profiles/memcheck_pipe.py`: every box mode
(plain, folded x / y / z, zero-bound clipping, dft fix-up, split tiles) of the persistent pull / grad kernels
(forced) and of the boxed push / count kernels (default), f32 and f16, plus the fused backward and a
displacement-field call."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'torch-interpol_b200')); sys.path.insert(0, os.path.join(ROOT, 'tests')); sys.path.insert(0, ROOT)
import torch
from test_gpu_ops import smooth_grid
from interpol_b200 import pushpull as pp
import interpol_b200 as ib

gen = torch.Generator().manual_seed(0)
shape = (40, 24, 64)
vol = torch.randn([1, 2, *shape], generator=gen).cuda()
seen = set()
for amp in (3.0, 25.0):
    grid = (smooth_grid(shape, gen, amp=amp) - 2.0).contiguous().cuda()
    for bound in ([3], [0], [6, 1, 6], [4, 5, 2]):
        pp.flags = 8
        try:
            for order in (1, 3):
                a = pp.grid_pull(vol, grid, bound, [order], 1); seen.add(ib.last_kernel())
                b = pp.grid_grad(vol, grid, bound, [order], 0); seen.add(ib.last_kernel())
                c = pp.grid_push(vol, grid, list(shape), bound, [order], 1); seen.add(ib.last_kernel())
                d = pp.grid_count(grid, list(shape), bound, [order], 2); seen.add(ib.last_kernel())
        finally:
            pp.flags = 0
        # boxed push / count in 16-bit storage (order 5: the cfg 4 instantiation), fused backward, displacement mode
        h = vol.half(); gh = grid.half()
        pp.grid_push(h, gh, list(shape), bound, [5], 1); seen.add(ib.last_kernel())
        pp.grid_count(gh, list(shape), bound, [5], 1); seen.add(ib.last_kernel())
        vr = vol.clone().requires_grad_(); gr = grid.clone().requires_grad_()
        pp.grid_pull_backward(torch.ones_like(vol), vr, gr, bound, [3], 1); seen.add(ib.last_kernel())
        pp.grid_pull(vol, grid * 0.1, bound, [3], 1, True); seen.add(ib.last_kernel())
torch.cuda.synchronize()
print('ran', sorted(seen))
