"""Golden vectors for `resize`: outputs of the UNMODIFIED reference (interpol.resize, CPU, float64) on seeded inputs.  Run in the build container, where /root/reference is importable:

    PYTHONPATH=/root/reference python tests/golden/make_golden_resize.py

writes tests/golden/resize.npz (inputs are regenerated from the seeds by the tests)."""
import os
import sys
import warnings

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))

CASES = []
for shape, factor in (((13,), [2.3]), ((9, 11), [2, 0.6]), ((7, 8, 6), [1.5, 2, 0.5])):
    for anchor in ('c', 'e', 'f', 'l'):
        for order in (0, 1, 2, 3, 5):
            for bound in ('nearest', 'dct2', 'dft', 'zero'):
                CASES.append(dict(shape=shape, factor=factor, anchor=anchor, order=order, bound=bound, prefilter=True, extrapolate=True))
CASES.append(dict(shape=(9, 11), factor=[1.7, 1.3], anchor='e', order=3, bound='dct1', prefilter=False, extrapolate=False))
CASES.append(dict(shape=(7, 8, 6), factor=[2, 2, 2], anchor='c', order=[1, 3, 2], bound=['dct2', 'zero', 'dft'], prefilter=True, extrapolate=True))
CASES.append(dict(shape=(7, 8, 6), factor=[0.7, 1.9, 1.2], anchor=['e', 'c', 'f'], order=4, bound='replicate', prefilter=True, extrapolate=2))


# restrict (the adjoint): a subset of the same option space
RCASES = []
for shape, factor in (((23,), [2.3]), ((12, 15), [2, 1.5]), ((10, 9, 8), [1.5, 2, 1.25])):
    for anchor in ('c', 'e', 'f', 'l'):
        for order, bound in ((0, 'nearest'), (1, 'zero'), (2, 'dct2'), (3, 'dft'), (5, 'dct1')):
            RCASES.append(dict(shape=shape, factor=factor, anchor=anchor, order=order, bound=bound, reduce_sum=(order % 2 == 0)))


def make_input(i, shape):
    g = torch.Generator().manual_seed(5000 + i)
    return torch.randn([1, 2, *shape], generator=g, dtype=torch.float64)


if __name__ == '__main__':
    warnings.filterwarnings('ignore')
    sys.path.insert(0, os.environ.get('IB200_REFERENCE', '/root/reference'))
    import interpol  # the unmodified reference: only needed to (re)generate the fixtures
    out = {}
    for i, c in enumerate(CASES):
        x = make_input(i, c['shape'])
        y = interpol.resize(x, factor=c['factor'], anchor=c['anchor'], interpolation=c['order'], bound=c['bound'],
                            prefilter=c['prefilter'], extrapolate=c['extrapolate'])
        out['case%d' % i] = y.numpy()
    for i, c in enumerate(RCASES):
        x = make_input(10000 + i, c['shape'])
        y = interpol.restrict(x, factor=c['factor'], anchor=c['anchor'], interpolation=c['order'], bound=c['bound'],
                              reduce_sum=c['reduce_sum'])
        out['rcase%d' % i] = y.numpy()
    np.savez_compressed(os.path.join(HERE, 'resize.npz'), **out)
    print('wrote', len(out), 'arrays for', len(CASES), '+', len(RCASES), 'cases')
