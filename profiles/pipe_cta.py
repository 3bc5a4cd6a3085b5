#!/usr/bin/env python
"""Per-CTA timeline of the persistent pull kernel (library built with -DIB200_PIPE_TIMERS): start / end of consumer
warp 0 of every CTA on the global timer, its loop ticks and the tiles it processed.
    python profiles/pipe_cta.py [--order N] [--bound B]"""
import ctypes, os, sys
os.environ['IB200_PIPE_DEBUG'] = '1'
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'torch-interpol_b200')); sys.path.insert(0, ROOT)
import numpy as np
import torch
from bench import make_workload
import interpol_b200 as ib
from interpol_b200 import pushpull as pp, _lib


def arg(name, default):
    return int(sys.argv[sys.argv.index(name) + 1]) if name in sys.argv else default


order, bound = arg('--order', 3), arg('--bound', 3)
vol, grid = make_workload(256, 'cuda')
for _ in range(3):
    pp.grid_pull(vol, grid, [bound], [order], 1)
torch.cuda.synchronize()
a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
a.record(); pp.grid_pull(vol, grid, [bound], [order], 1); pp.grid_pull(vol, grid, [bound], [order], 1); b.record(); torch.cuda.synchronize()
buf = (ctypes.c_longlong * 1024)()
_lib.lib().ib200_debug_pipe_cta(buf)
d = np.array(buf[:148 * 4], dtype=np.int64).reshape(148, 4)
t0 = d[:, 0].min()
start, end, ticks, tiles = (d[:, 0] - t0) / 1e3, (d[:, 1] - t0) / 1e3, d[:, 2], d[:, 3]
print('order', order, 'bound', bound, 'two launches %.1f us' % (a.elapsed_time(b) * 1e3), ib.last_kernel())
print('consumer start  us: min %.1f  max %.1f' % (start.min(), start.max()))
print('consumer end    us: min %.1f  mean %.1f  max %.1f' % (end.min(), end.mean(), end.max()))
print('loop duration   us: min %.1f  mean %.1f  max %.1f' % ((end - start).min(), (end - start).mean(), (end - start).max()))
print('loop ticks        : min %d  mean %d  max %d   -> %.3f GHz' % (ticks.min(), ticks.mean(), ticks.max(), (ticks / ((end - start) * 1e3)).mean()))
print('tiles per CTA     : min %d  mean %.1f  max %d  total %d' % (tiles.min(), tiles.mean(), tiles.max(), tiles.sum()))
print('ticks per tile    : min %.0f  mean %.0f  max %.0f' % ((ticks / tiles).min(), (ticks / tiles).mean(), (ticks / tiles).max()))
