import os, sys
ROOT='/root/repo'
sys.path.insert(0, os.path.join(ROOT, 'torch-interpol_b200')); sys.path.insert(0, os.path.join(ROOT, 'tests')); sys.path.insert(0, ROOT)
import random, torch, numpy as np
from test_gpu_ops import smooth_grid
import interpol_b200 as ib
from interpol_b200 import pushpull as pp
import oracle
oracle.set_num_threads(8)
def case(shape, vshape, B, order, bound, ex, amp, shift, seedcase, noise=None):
    # regenerate exactly like the fuzz would is complicated; just build a similar problem
    gen = torch.Generator().manual_seed(seedcase)
    grid = smooth_grid(shape, gen, amp=amp, batch=B)
    grid = grid * torch.tensor([vshape[d] / shape[d] for d in range(3)]) + shift
    grid = grid.contiguous()
    g = grid.cuda()
    fast = pp.grid_count(g, list(vshape), bound, [order], ex); kf = ib.last_kernel()
    pp.flags = 1
    slow = pp.grid_count(g, list(vshape), bound, [order], ex); ks = ib.last_kernel()
    pp.flags = 0
    want = oracle.grid_count(grid.double().numpy(), list(vshape), bound, [order], ex, nthreads=8)
    sc = np.abs(want).max()
    print(kf, ks, 'scale %.4g' % sc, 'fast err %.3g (rel %.2e)' % (np.abs(fast.double().cpu().numpy() - want).max(), np.abs(fast.double().cpu().numpy() - want).max() / sc),
          'generic err %.3g (rel %.2e)' % (np.abs(slow.double().cpu().numpy() - want).max(), np.abs(slow.double().cpu().numpy() - want).max() / sc))
for shift in (0.0, -2.5, 40.0):
    case((58, 39, 64), (39, 8, 22), 2, 5, [4, 4, 1], 1, 8.0, shift, 125)
    case((67, 58, 72), (43, 31, 53), 1, 4, [1], 1, 30.0, shift, 158)
