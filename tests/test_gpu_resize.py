"""GPU tests of the separable `resize` (one ib200_resample_axis pass per axis): against golden outputs
of the unmodified reference (tests/golden/resize.npz, made by tests/golden/make_golden_resize.py), and
against this package's own dense-grid path (grid_pull on the tensor-product grid, itself checked
against the oracle) over anchors x orders x bounds x dims x dtypes."""
import os
import sys

import numpy as np
import pytest
import torch

from conftest import rel_err

pytestmark = pytest.mark.gpu

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden'))


def test_resize_vs_reference_golden():
    import interpol_b200 as ib
    import make_golden_resize as mg
    gold = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'resize.npz'))
    worst = {torch.float64: 0.0, torch.float32: 0.0}
    for i, c in enumerate(mg.CASES):
        x = mg.make_input(i, c['shape'])
        want = gold['case%d' % i]
        for dt, tol in ((torch.float64, 1e-10), (torch.float32, 1e-5)):
            got = ib.resize(x.to(dt).cuda(), factor=c['factor'], anchor=c['anchor'], interpolation=c['order'],
                            bound=c['bound'], prefilter=c['prefilter'], extrapolate=c['extrapolate'])
            assert ib.last_kernel().startswith('resample_axis'), ib.last_kernel()
            assert got.dtype == dt and tuple(got.shape) == want.shape, (c, got.shape, want.shape)
            scale = np.abs(want).max()
            err = 0.0 if scale == 0 else np.abs(got.double().cpu().numpy() - want).max() / scale
            worst[dt] = max(worst[dt], err)
            assert err <= tol, (i, c, dt, err)
    print('worst errors', worst)


@pytest.mark.parametrize('dtype,tol', [(torch.float32, 2e-6), (torch.float64, 1e-12), (torch.float16, 2e-3), (torch.bfloat16, 2e-2)],
                         ids=['f32', 'f64', 'f16', 'bf16'])
@pytest.mark.parametrize('shape,outshape', [((300,), (77,)), ((40, 52), (64, 33)), ((24, 20, 28), (36, 40, 16)), ((20, 24, 32), (20, 48, 32))])
def test_separable_matches_dense_grid(shape, outshape, dtype, tol):
    import interpol_b200 as ib
    import importlib
    rz = importlib.import_module('interpol_b200.resize')      # (the package attribute `resize` is the function)
    gen = torch.Generator().manual_seed(len(shape) * 100 + outshape[0])
    x = torch.randn([2, 3, *shape], generator=gen).to(dtype).cuda()
    for anchor in ('c', 'e', 'f', 'l'):
        for order, bound, ex in ((0, 'nearest', True), (1, 'zero', False), (2, 'dct1', True), (3, 'dct2', True), (3, 'dst2', 2),
                                 (4, 'dft', True), (5, 'dst1', False), (7, 'replicate', True)):
            kw = dict(shape=list(outshape), anchor=anchor, interpolation=order, bound=bound, extrapolate=ex, prefilter=False)
            if anchor in 'fl':
                kw = dict(factor=[o / i for o, i in zip(outshape, shape)], anchor=anchor, interpolation=order, bound=bound,
                          extrapolate=ex, prefilter=False)
            a = ib.resize(x, **kw)
            assert ib.last_kernel().startswith('resample_axis'), ib.last_kernel()
            rz.SEPARABLE = False
            try:
                b = ib.resize(x, **kw)
            finally:
                rz.SEPARABLE = True
            assert not ib.last_kernel().startswith('resample_axis')
            assert a.shape == b.shape and a.dtype == b.dtype
            if order == 0 and dtype in (torch.float16, torch.bfloat16):
                continue        # 16-bit coordinates: nearest picks may legitimately differ at exact halves
            scale = b.double().abs().max().item()
            err = 0.0 if scale == 0 else (a.double() - b.double()).abs().max().item() / scale
            assert err <= tol, (shape, outshape, anchor, order, bound, ex, err)


def test_resize_cpu_tensor_and_fallbacks():
    import interpol_b200 as ib
    x = torch.randn(1, 1, 12, 10)
    y = ib.resize(x, factor=[2, 2], interpolation=3, bound='dct2')
    assert not y.is_cuda and tuple(y.shape) == (1, 1, 24, 20)
    xg = torch.randn(1, 1, 12, 10, device='cuda', requires_grad=True)
    ib.resize(xg, factor=[2, 2], interpolation=1, prefilter=False).sum().backward()     # autograd: dense-grid path
    assert xg.grad is not None and tuple(xg.grad.shape) == (1, 1, 12, 10)
