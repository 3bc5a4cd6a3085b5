#!/usr/bin/env python
"""Same-device baseline (optional): the UNMODIFIED reference (baseline/_ref, TorchScript + ATen ops) on CUDA
tensors of the same B200, next to this engine, on the headline workload (256^3 fp32 cubic dct2 pull + push) and on
cfg 2 (128^3).  Device time by CUDA events, median of 5 after 2 warm-ups (the first call compiles TorchScript)."""
import os
import statistics
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'torch-interpol_b200')); sys.path.insert(0, ROOT)
import torch  # noqa: E402
import bench  # noqa: E402
import interpol_b200 as ib  # noqa: E402


def timeit(fn, n=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
        a.record(); out = fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
        del out
    return statistics.median(ts)


def main():
    ref = bench.load_reference()
    if ref is None:
        print('baseline/_ref not installed: nothing to compare with')
        return
    kw = dict(interpolation=3, bound='dct2', extrapolate=True)
    for n in (128, 256):
        vol, grid = bench.make_workload(n, 'cuda')
        shape = [n] * 3
        rows = []
        for name, mod in (('reference (TorchScript, cuda)', ref), ('interpol_b200', ib)):
            t_pull = timeit(lambda: mod.grid_pull(vol, grid, **kw))
            pulled = mod.grid_pull(vol, grid, **kw)
            t_push = timeit(lambda: mod.grid_push(pulled, grid, shape=shape, **kw))
            rows.append((name, t_pull, t_push, pulled))
            torch.cuda.empty_cache()
        err = (rows[0][3] - rows[1][3]).abs().max().item() / rows[0][3].abs().max().item()
        for name, tp, ts, _ in rows:
            print('%d^3 cubic dct2  %-30s pull %9.3f ms  push %9.3f ms  (%.0f Mvox/s pull+push)'
                  % (n, name, tp, ts, 2 * n ** 3 / ((tp + ts) * 1e-3) / 1e6))
        print('%d^3: speed-up pull %.0fx, push %.0fx; max-norm relative difference of the pulled volumes %.2e; peak memory of the '
              'reference call %.1f GB' % (n, rows[0][1] / rows[1][1], rows[0][2] / rows[1][2], err, torch.cuda.max_memory_allocated() / 1e9))


if __name__ == '__main__':
    main()
