#!/usr/bin/env python
"""Per-op device timings (CUDA events, median of N after warm-up) used while
tuning; prints one line per op with Mvox/s and the fraction of the measured HBM
roofline (algorithmic bytes of SURVEY 8(d))."""
import argparse
import json
import os
import sys
import statistics

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'torch-interpol_b200'))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from bench import make_workload, measured_peak  # noqa: E402
import interpol_b200 as ib  # noqa: E402
from interpol_b200 import pushpull as pp  # noqa: E402


def timeit(fn, n=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return statistics.median(ts), min(ts)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--size', type=int, default=256)
    ap.add_argument('--order', type=int, default=3)
    ap.add_argument('--bound', type=int, default=3)
    ap.add_argument('--channels', type=int, default=1)
    ap.add_argument('--dtype', default='f32')
    ap.add_argument('--ops', default='pull,pull_generic,push,push_generic,count,grad,coeff')
    ap.add_argument('--incoherent', action='store_true')
    ap.add_argument('--flags', type=int, default=0, help='IB200_FLAG_* bits for the non-generic ops (4: no pipe / box kernels)')
    ap.add_argument('--force-pipe', action='store_true', help='take the persistent kernels wherever they apply (IB200_FLAG_FORCE_PIPE)')
    a = ap.parse_args()
    peak, _ = measured_peak()
    dt = {'f32': torch.float32, 'f16': torch.float16, 'f64': torch.float64, 'bf16': torch.bfloat16}[a.dtype]
    s = {'f32': 4, 'f16': 2, 'f64': 8, 'bf16': 2}[a.dtype]
    n = a.size
    vol, grid = make_workload(n, 'cuda')
    if a.incoherent:
        grid = grid + torch.randn_like(grid) * 20
    vol = vol.expand(1, a.channels, n, n, n).contiguous().to(dt)
    grid = grid.to(dt)
    C, N = a.channels, n ** 3
    b, o = [a.bound], [a.order]
    res = {}
    for op in a.ops.split(','):
        pp.flags = 1 if op.endswith('_generic') else (8 if a.force_pipe else a.flags)
        base = op.replace('_generic', '')
        if base == 'pull':
            fn = lambda: pp.grid_pull(vol, grid, b, o, 1); by = N * (3 * s + 2 * C * s)
        elif base == 'push':
            fn = lambda: pp.grid_push(vol, grid, [n] * 3, b, o, 1); by = N * (3 * s + 2 * C * s)
        elif base == 'count':
            fn = lambda: pp.grid_count(grid, [n] * 3, b, o, 1); by = N * (3 * s + s)
        elif base == 'grad':
            fn = lambda: pp.grid_grad(vol, grid, b, o, 1); by = N * (3 * s + C * s + 3 * C * s)
        elif base == 'bwd':          # GridPull.backward: push(grad) + fused sum_c grad(vol) * grad
            vr, gr = vol.clone().requires_grad_(), grid.clone().requires_grad_()
            gout = torch.randn_like(vol)
            fn = lambda: pp.grid_pull_backward(gout, vr, gr, b, o, 1); by = N * (3 * s + 2 * C * s) + N * (3 * s + 2 * C * s + 3 * s)
        elif base == 'bwd_grid':     # grid branch alone
            gout = torch.randn_like(vol)
            fn = lambda: pp.grid_pull_grad_grid(gout, vol, grid, b, o, 1); by = N * (3 * s + 2 * C * s + 3 * s)
        elif base == 'coeff':
            fn = lambda: ib.spline_coeff_nd(vol, interpolation=a.order, bound=a.bound, dim=3); by = 2 * C * N * s
        else:
            continue
        med, best = timeit(fn)
        units = C * N if base == 'coeff' else N
        res[op] = dict(ms=med, best_ms=best, mvox_s=units / med / 1e3, frac=by / (med * 1e-3) / 1e9 / peak,
                       kernel=ib.last_kernel())
        print('%-14s %9.3f ms (best %8.3f)  %10.0f Mvox/s  %5.1f %% of %d GB/s   [%s]' % (
            op, med, best, res[op]['mvox_s'], 100 * res[op]['frac'], peak, res[op]['kernel']))
    pp.flags = 0
    print(json.dumps(dict(config=vars(a), results=res)))


if __name__ == '__main__':
    main()
