"""Diagnose the two tile-parity failures of r2a: (order 1, dct1, extrapolate 0, pull) and (order 5, replicate, count)."""
import os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, 'torch-interpol_b200'), os.path.join(ROOT, 'tests')]
import oracle
import interpol_b200 as ib
from interpol_b200 import pushpull as pp
from test_gpu_tile_parity import _case, VSHAPE, SHAPE

def rel(a, b):
    return np.abs(a - b).max() / np.abs(b).max()

def run(order, bound, ex, op):
    vol, img, grid = _case(order, 900 + 10 * order + bound, amp=6.0)
    grid = (grid * 1.2 - 3.0).contiguous()
    b, o = [bound], [order]
    v64, i64, g64 = vol.double().numpy(), img.double().numpy(), grid.double().numpy()
    if op == 'pull':
        f = lambda: pp.grid_pull(vol.cuda(), grid.cuda(), b, o, ex)
        want = oracle.grid_pull(v64, g64, b, o, ex)
        want32 = oracle.grid_pull(vol.numpy(), grid.numpy(), b, o, ex)
    else:
        f = lambda: pp.grid_count(grid.cuda(), list(VSHAPE), b, o, ex)
        want = oracle.grid_count(g64, VSHAPE, b, o, ex)
        want32 = oracle.grid_count(grid.numpy(), VSHAPE, b, o, ex)
    got = f().double().cpu().numpy(); k = ib.last_kernel()
    pp.flags = 1
    gen = f().double().cpu().numpy(); kg = ib.last_kernel()
    pp.flags = 0
    print('order %d bound %d ex %d %s: tile(%s) vs f64 oracle %.3g | generic(%s) vs oracle %.3g | oracle f32 vs f64 %.3g | tile vs generic %.3g'
          % (order, bound, ex, op, k, rel(got, want), kg, rel(gen, want), rel(want32.astype(np.float64), want), rel(got, gen)))
    d = np.abs(got - want)
    idx = np.unravel_index(np.argsort(d.ravel())[-5:], d.shape)
    for t in zip(*idx):
        line = '   at %s got %.6f want %.6f generic %.6f' % (t, got[t], want[t], gen[t])
        if op == 'pull':
            line += ' coord %s' % (g64[(t[0],) + t[2:]],)
        print(line)
    print('   n bad (> 1e-5 max):', int((d > 1e-5 * np.abs(want).max()).sum()), 'of', d.size)

run(1, 2, 0, 'pull')
run(1, 2, 1, 'pull')
run(1, 3, 0, 'pull')
run(5, 1, 1, 'count')
run(3, 1, 1, 'count')
