import sys, os
sys.path[:0] = ['/root/repo', '/root/repo/torch-interpol_b200', '/root/repo/tests']
import torch
import interpol_b200 as ib
from test_gpu_ops import smooth_grid
gen = torch.Generator().manual_seed(31)
shape = (40, 36, 44)
vol = torch.randn([1, 2, *shape], generator=gen)
grid = smooth_grid(shape, gen)
kw = dict(interpolation=3, bound='dct2', extrapolate=True)
for B, C in ((1, 2), (1, 1), (2, 2)):
    v = torch.randn([B, C, *shape], generator=gen).cuda(); g = smooth_grid(shape, gen, batch=B).cuda()
    out = ib.grid_push(v, g, **kw); print(B, C, 'push ->', ib.last_kernel(), g.stride(), v.stride())
    out = ib.grid_push(v, g.contiguous(), **kw); print(B, C, 'push contiguous grid ->', ib.last_kernel())
