#!/usr/bin/env python
"""Phase timers of the pull pipe kernel (CTA 0: producer warp and consumer warp 0), in SM clock ticks.
    python profiles/pipe_debug.py [--order N] [--channels C] [--bound B]
The timers are compiled in only when the library is built with IB200_TUNE=1 (python torch-interpol_b200/build.py --force)."""
import ctypes, os, sys
os.environ['IB200_PIPE_DEBUG'] = '1'
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'torch-interpol_b200')); sys.path.insert(0, ROOT)
import torch
from bench import make_workload
import interpol_b200 as ib
from interpol_b200 import pushpull as pp, _lib


def arg(name, default):
    return int(sys.argv[sys.argv.index(name) + 1]) if name in sys.argv else default


order, C, bound = arg('--order', 3), arg('--channels', 1), arg('--bound', 3)
vol, grid = make_workload(256, 'cuda')
vol = vol.expand(1, C, 256, 256, 256).contiguous()
for _ in range(3):
    pp.grid_pull(vol, grid, [bound], [order], 1)
torch.cuda.synchronize()
a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
a.record(); pp.grid_pull(vol, grid, [bound], [order], 1); b.record(); torch.cuda.synchronize()
buf = (ctypes.c_longlong * 16)()
_lib.lib().ib200_debug_pipe_counters(buf)
names = ['P wait kfull', 'P plan', 'P request grid', 'P wait bempty', 'P issue box', 'C wait gfull', 'C wait bfull',
         'C geom+fixup', 'C rows (taps)', 'C merge+release', 'C items', 'C whole loop']
items = max(buf[10], 1)
print('order', order, 'C', C, 'bound', bound, 'kernel %.1f us' % (a.elapsed_time(b) * 1e3), ib.last_kernel(), 'items', items)
for i, nm in enumerate(names):
    print('%-16s total %10d ticks   per item %8.0f' % (nm, buf[i], buf[i] / items))
# ---- push ----
img = vol
pp.flags = 8
for _ in range(3):
    pp.grid_push(img, grid, [256] * 3, [bound], [order], 1)
torch.cuda.synchronize()
a.record(); pp.grid_push(img, grid, [256] * 3, [bound], [order], 1); b.record(); torch.cuda.synchronize()
_lib.lib().ib200_debug_push_counters(buf)
names = ['P wait bdone', 'P flush issue', 'P flush read', 'P wait kfull', 'C wait bfull', 'C zero', 'C cells', 'C scale', 'C atomics',
         'C wait rows', 'C fold', 'C convert', 'C items', 'C whole loop']
items = max(buf[12], 1)
print('PUSH order', order, 'C', C, 'bound', bound, 'kernel %.1f us' % (a.elapsed_time(b) * 1e3), ib.last_kernel(), 'items', items)
for i, nm in enumerate(names):
    print('%-16s total %10d ticks   per item %8.0f' % (nm, buf[i], buf[i] / items))
