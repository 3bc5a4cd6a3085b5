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


def test_restrict_vs_reference_golden():
    import interpol_b200 as ib
    import make_golden_resize as mg
    gold = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'resize.npz'))
    for i, c in enumerate(mg.RCASES):
        x = mg.make_input(10000 + i, c['shape'])
        want = gold['rcase%d' % i]
        for dt, tol in ((torch.float64, 1e-10), (torch.float32, 1e-5)):
            if dt == torch.float32 and c['order'] == 0:
                continue        # nearest at coordinates that are exact halves in float64 (e.g. i * 2/3 - 1/6): the
                                # float32 lattice lands on either side; float64 is compared above
            got = ib.restrict(x.to(dt).cuda(), factor=c['factor'], anchor=c['anchor'], interpolation=c['order'],
                              bound=c['bound'], reduce_sum=c['reduce_sum'])
            assert ib.last_kernel().startswith('resample_adjoint'), ib.last_kernel()
            assert got.dtype == dt and tuple(got.shape) == want.shape, (c, got.shape, want.shape)
            scale = np.abs(want).max()
            err = 0.0 if scale == 0 else np.abs(got.double().cpu().numpy() - want).max() / scale
            assert err <= tol, (i, c, dt, err)


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


@pytest.mark.parametrize('dtype,tol', [(torch.float32, 4e-6), (torch.float64, 1e-12)], ids=['f32', 'f64'])
@pytest.mark.parametrize('shape,outshape', [((300,), (77,)), ((52, 40), (20, 33)), ((24, 20, 28), (12, 13, 16))])
def test_separable_restrict_matches_dense_grid(shape, outshape, dtype, tol):
    import importlib
    import interpol_b200 as ib
    rs = importlib.import_module('interpol_b200.restrict')
    gen = torch.Generator().manual_seed(len(shape) * 100 + outshape[0])
    x = torch.randn([2, 3, *shape], generator=gen).to(dtype).cuda()
    for anchor in ('c', 'e', 'f', 'l'):
        for order, bound, ex in ((0, 'nearest', True), (1, 'zero', False), (2, 'dct1', True), (3, 'dct2', True), (3, 'dst2', 2),
                                 (4, 'dft', True), (5, 'dst1', False), (7, 'replicate', True)):
            for reduce_sum in (False, True):
                kw = dict(shape=list(outshape), anchor=anchor, interpolation=order, bound=bound, extrapolate=ex, reduce_sum=reduce_sum)
                if anchor in 'fl':
                    kw = dict(factor=[i / o for o, i in zip(outshape, shape)], anchor=anchor, interpolation=order, bound=bound,
                              extrapolate=ex, reduce_sum=reduce_sum)
                a = ib.restrict(x, **kw)
                assert ib.last_kernel().startswith('resample_adjoint'), ib.last_kernel()
                rs.SEPARABLE = False
                try:
                    b = ib.restrict(x, **kw)
                finally:
                    rs.SEPARABLE = True
                assert not ib.last_kernel().startswith('resample_adjoint')
                assert a.shape == b.shape and a.dtype == b.dtype
                scale = b.double().abs().max().item()
                err = 0.0 if scale == 0 else (a.double() - b.double()).abs().max().item() / scale
                assert err <= tol, (shape, outshape, anchor, order, bound, ex, reduce_sum, err)


def test_separable_resize_restrict_autograd():
    """resize's backward is the adjoint pass and vice versa: <resize(x), y> == <x, d/dx <resize(x), y>>,
    gradcheck in float64, and agreement with the dense-grid path's gradients."""
    import importlib
    import interpol_b200 as ib
    rz = importlib.import_module('interpol_b200.resize')
    gen = torch.Generator().manual_seed(3)
    x = torch.randn([1, 2, 9, 7], generator=gen, dtype=torch.float64).cuda().requires_grad_()
    kw = dict(factor=[1.7, 2.2], anchor='e', interpolation=3, bound='dct2', prefilter=True)
    assert torch.autograd.gradcheck(lambda t: ib.resize(t, **kw), (x,), rtol=1e-6, atol=1e-8, nondet_tol=1e-12)
    assert torch.autograd.gradcheck(lambda t: ib.restrict(t, factor=[1.5, 2], anchor='c', interpolation=2, bound='dft'), (x,),
                                    rtol=1e-6, atol=1e-8, nondet_tol=1e-12)
    y = ib.resize(x, **kw)
    assert ib.last_kernel().startswith('resample_axis')
    w = torch.randn(y.shape, generator=gen, dtype=torch.float64).cuda()
    (g_sep,) = torch.autograd.grad((y * w).sum(), x)
    rz.SEPARABLE = False
    try:
        (g_dense,) = torch.autograd.grad((ib.resize(x, **kw) * w).sum(), x)
    finally:
        rz.SEPARABLE = True
    assert (g_sep - g_dense).abs().max().item() <= 1e-10 * g_dense.abs().max().item()
